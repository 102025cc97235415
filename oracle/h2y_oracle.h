/*
 * h2y_oracle.h -- CPU restatement of the hdr2yuv / yuv2tiff per-pixel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / reported CPU baseline.
 * The product path (hdr2yuv_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  The restatement is checked sample-for-sample against
 * the reference's own sources compiled unmodified (oracle/_ref/libh2yref.so,
 * built by oracle/build.py from /root/reference) in tests/test_oracle_cpu.py,
 * and against golden vectors generated from that compiled reference
 * (tests/golden/, generator tests/golden/make_golden.py).  The reference ships
 * no golden vectors of its own (test.sh:89-94 compares against ref/ files that
 * are not in the repository).  The only arithmetic outside the reference is
 * glibc libm pow/log/fmax (convert.cpp:24,36,48,61,71-73,83-85), unpinned by the
 * reference; the stated tolerance there is <= 1 output code, deviations counted.
 *
 * Build: gcc -O2 -ffp-contract=off (x86-64 SSE2 scalar, no FMA, as make.sh:4-16).
 * Plane order everywhere is index 0/1/2 = G/B/R = Y/Cb/Cr (tiff.cpp:309-311).
 */
#ifndef H2Y_ORACLE_H
#define H2Y_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* hdr.h:14-17 */
enum { ORC_CHROMA_400 = 0, ORC_CHROMA_420 = 1, ORC_CHROMA_422 = 2, ORC_CHROMA_444 = 3 };
/* hdr.h:296-297 */
enum { ORC_PIC_U16 = 1, ORC_PIC_F32 = 2 };
/* hdr.h:104-134 (only the values the hot path tests) */
enum {
    ORC_TRANSFER_BT709 = 1, ORC_TRANSFER_BT601 = 6, ORC_TRANSFER_LINEAR = 8,
    ORC_TRANSFER_BT2020_10 = 14, ORC_TRANSFER_BT2020_12 = 15, ORC_TRANSFER_PQ = 16,
    ORC_TRANSFER_RHO_GAMMA = 18
};
/* hdr.h:168-191 */
enum {
    ORC_MATRIX_GBR = 0, ORC_MATRIX_BT709 = 1, ORC_MATRIX_BT2020NC = 9, ORC_MATRIX_BT2020C = 10,
    ORC_MATRIX_YDZDX = 11, ORC_MATRIX_YDZDX_Y500 = 12, ORC_MATRIX_YDZDX_Y100 = 13,
    ORC_MATRIX_YUVPRIME1 = 14, ORC_MATRIX_YUVPRIME2 = 15
};

/* hdr.h:345-356 */
typedef struct {
    unsigned long minCV, maxCV;
    unsigned short minVR, maxVR, minVRC, maxVRC, Half;
} orc_clip_t;

/* the pic_t fields the hot path reads (hdr.h:359-392) */
typedef struct {
    int width, height;
    int chroma_format_idc;
    int transfer_characteristics, colour_primaries, matrix_coeffs;
    int bit_depth, video_full_range_flag;
    int pic_buffer_type;
    int est_floor[3], est_ceiling[3];        /* pic_stats_t.estimated_* (hdr.h:315-316) */
    uint16_t *buf[3];
    float *fbuf[3];
} orc_pic_t;

void orc_set_clip(int bit_depth, int full_range, orc_clip_t *clip);          /* common.cpp:300-327 */
void orc_plane_dims(int w, int h, int chroma, int pw[3], int ph[3]);           /* common.cpp:191-198 */
void orc_pic_stats(orc_pic_t *pic, float fmin_out[3], float fmax_out[3]);      /* common.cpp:66-168 */

float orc_pq_eotf(float V);                 /* PQ10000_f   convert.cpp:43-51 */
float orc_pq_oetf(float L);                 /* PQ10000_r   convert.cpp:56-63 */
float orc_bt1886_eotf(float V);             /* bt1886_f with g=2.4f,Lw=1,Lb=0  convert.cpp:67-75,1051-1057 */
float orc_bt1886_oetf(float L);             /* bt1886_r   convert.cpp:79-87,1096-1102 */
float orc_rho_gamma_eotf(float V);          /* RHO_GAMMA_f convert.cpp:12-26 */
float orc_rho_gamma_oetf(float L);          /* RHO_GAMMA_r convert.cpp:29-39 */

/* convert.cpp:879-1315 (U16-output branch 1146-1220; F32-output branch 1222-1304) */
int orc_matrix_convert(orc_pic_t *out, const orc_pic_t *in);

/* convert.cpp:261-383 / 91-172; row-major planes */
void orc_subsample_fir(uint16_t *dst, const uint16_t *src, int w, int h,
                       unsigned long minCV, unsigned long maxCV);
void orc_subsample_fir_h(uint16_t *dst422, const uint16_t *src, int w, int h,
                         unsigned long minCV, unsigned long maxCV);            /* stage 1 only: 290-321 */
void orc_subsample_box(uint16_t *dst, const uint16_t *src, int w, int h);
/* convert.cpp:513-874 without the Y'u''v'' branch; resampler 0=box else FIR */
int orc_convert(orc_pic_t *out, const orc_pic_t *in, int resampler);
/* tiff.cpp:457-550: shift + range clamp in place on the logical plane sizes */
int orc_write_yuv_clamp(orc_pic_t *pic, int src_bit_depth);

/* whole forward chain as main() sequences it (hdr2yuv.cpp:791-930).
 * in: 4:4:4 planar G,B,R (buf or fbuf).  dst_yuv: Y then Cb then Cr, the .yuv frame layout. */
typedef struct {
    int dst_bit_depth, dst_full_range, dst_transfer, dst_primaries, dst_matrix, dst_chroma;
    int resampler;
} orc_dst_t;
int orc_forward_frame(orc_pic_t *in, const orc_dst_t *dst, uint16_t *dst_yuv);

/* read_tiff de-interleave + on-read clip (tiff.cpp:265-315) and read_exr widening (exr.cpp:209-235) */
void orc_load_rgb16(const uint16_t *rgb, int npix, int nchan, int full_range, uint16_t *planes[3]);
void orc_load_half(const uint16_t *half_px, int npix, int nchan, float *planes[3]);
float orc_half_to_float(uint16_t h);

/* yuv2tiff.cpp:575-692 on row-major planes (src is (w/2)x(h/2), dst is w x h) */
void orc_upsample_420to444(const uint16_t *src, uint16_t *dst, int w, int h, int fir,
                           unsigned short minCV, unsigned short maxCV);

enum { ORC_INV_YDZDX = 0, ORC_INV_709 = 1, ORC_INV_2020 = 2, ORC_INV_Y100 = 3, ORC_INV_Y500 = 4 };
typedef struct {
    int width, height;
    int bit_depth;       /* 10, 12 (default), 14: B10/B14 keywords yuv2tiff.cpp:137-157 */
    int matrix;          /* ORC_INV_* : 709/2020/Y100/Y500 keywords, default Y'DzDx */
    int fir;             /* 0 when BOX given (yuv2tiff.cpp:117) */
    int full_range;      /* FULL keyword */
    int alpha;           /* ALPHA keyword: 4 samples per pixel, A=65535 */
    int ybar;            /* -X keyword: Y'DzDx rebuilt around the 2x2 mean of Y' (yuv2tiff.cpp:162, 365-399) */
} orc_inv_params;
/* one loop iteration of yuv2tiff main (yuv2tiff.cpp:278-552): yuv = Y,Cb,Cr 4:2:0 planes;
 * rgb = interleaved R,G,B(,A) 16-bit rows.  Returns the invalidPixels count. */
long orc_yuv2tiff_frame(const orc_inv_params *p, const uint16_t *yuv, uint16_t *rgb);

/* hdr2yuv's 4:4:4 inverse (convert.cpp:1320-1867) on U16 planes Y,Cb,Cr -> G,B,R; returns invalidPixels or -2 */
long orc_matrix_inverse(uint16_t *out_planes[3], const uint16_t *in_planes[3], int w, int h, int matrix_coeffs,
                        int in_bit_depth, int in_full_range, int out_bit_depth);
/* write_tiff's compute (tiff.cpp:605-628): planes G,B,R -> interleaved R,G,B, << (pic depth - src depth) */
void orc_write_tiff_rows(uint16_t *rgb, const uint16_t *planes[3], long npix, int pic_bit_depth, int src_bit_depth);

#ifdef __cplusplus
}
#endif
#endif
