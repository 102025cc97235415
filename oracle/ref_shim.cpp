/*
 * ref_shim.cpp -- extern "C" doorways into the UNMODIFIED reference sources.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled by oracle/build.py together with the reference's own
 * convert.cpp / common.cpp / tiff.cpp (read in place from /root/reference, never copied)
 * into oracle/_ref/libh2yref.so.  It sequences the reference functions exactly as the
 * reference main() does (hdr2yuv.cpp:791-930) on caller-supplied planes, so that tests can
 * pin the C restatement (h2y_oracle.c) and the CUDA path against the real thing, and so that
 * bench.py can time the reference's CPU path (cpu_baseline.kind = "reference").
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fcntl.h>
#include <unistd.h>

#include "hdr.h"

/* not declared in hdr.h */
void Subsample444to420_FIR(unsigned short *dst_plane, unsigned short *src_plane, short width, short height,
                           unsigned long minCV, unsigned long maxCV);
void Subsample444to420_box(unsigned short *dst_plane, unsigned short *src_plane, short width, short height,
                           unsigned long minCV, unsigned long maxCV);
void Subsample420to444(unsigned short **src, unsigned short **dst, short width, short height, short algorithmn,
                       unsigned short minCV, unsigned short maxCV);
float PQ10000_f(float V);
float PQ10000_r(float L);
float bt1886_f(float V, float gamma, float Lw, float Lb);
float bt1886_r(float L, float gamma, float Lw, float Lb);
float RHO_GAMMA_f(float V);
float RHO_GAMMA_r(float L);

/* symbols hdr2yuv's other translation units would provide; never reached from here */
void read_exr(pic_t *, char *) { abort(); }
int write_exr_file(char *, int, int, int, pic_t *) { abort(); }
/* dpx_read / dpx_write_float come from the reference's own dpx.cpp (built with oracle/stub/h2y_half_stub.h) */

namespace {
/* the reference printf()s on every call; park stdout on /dev/null while it runs */
struct Quiet {
    int saved;
    Quiet()
    {
        fflush(stdout);
        saved = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        dup2(nul, 1);
        close(nul);
    }
    ~Quiet()
    {
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};
double now_s()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
}  // namespace

extern "C" {

/* cfg: [0]w [1]h [2]src_is_f32 [3]src_bit_depth [4]src_full_range [5]src_transfer [6]src_primaries
 *      [7]src_matrix [8]dst_bit_depth [9]dst_full_range [10]dst_transfer [11]dst_primaries
 *      [12]dst_matrix [13]dst_chroma [14]resampler
 * planes: G,B,R (u16 or f32).  dst_yuv: Y,Cb,Cr packed like a .yuv frame.
 * tmp444 (optional): the three planes matrix_convert produced.  stats (optional): floor[3], ceiling[3].
 * seconds (optional): time spent in pic_stats + matrix_convert + convert + write_yuv. */
int ref_forward_frame(const int *cfg, const void *const planes[3], unsigned short *dst_yuv,
                      unsigned short *tmp444, int *stats, double *seconds)
{
    Quiet q;
    const int w = cfg[0], h = cfg[1];
    const size_t n = (size_t)w * h;
    static hdr_t hd;
    memset(&hd, 0, sizeof(hd));
    hd.user_args.chroma_resampler_type = cfg[14];
    hd.user_args.verbose_level = 0;
    pic_t *in = &hd.in_pic, *out = &hd.out_pic;
    pic_t tmp;
    memset(&tmp, 0, sizeof(tmp));

    const int in_type = cfg[2] ? PIC_TYPE_F32 : PIC_TYPE_U16;
    init_pic(in, w, h, CHROMA_444, cfg[3], cfg[4], cfg[6], cfg[5], cfg[7], 0, in_type, 0, 0, "in_pic");
    for (int c = 0; c < 3; c++) {
        if (cfg[2]) memcpy(in->fbuf[c], planes[c], n * sizeof(float));
        else memcpy(in->buf[c], planes[c], n * sizeof(unsigned short));
    }
    /* what parse_options leaves in out_pic for a .yuv destination */
    out->width = w; out->height = h;
    out->bit_depth = cfg[8]; out->video_full_range_flag = cfg[9];
    out->transfer_characteristics = cfg[10]; out->colour_primaries = cfg[11];
    out->matrix_coeffs = cfg[12]; out->chroma_format_idc = cfg[13];
    out->pic_buffer_type = PIC_TYPE_U16;

    int rc = 0;
    double t0 = now_s();
    pic_stats(in, &in->stats, 1);
    const int tmp_depth = out->pic_buffer_type == in->pic_buffer_type ? in->bit_depth : out->bit_depth;
    init_pic(&tmp, w, h, in->chroma_format_idc, tmp_depth, out->video_full_range_flag, out->colour_primaries,
             out->transfer_characteristics, out->matrix_coeffs, 0, out->pic_buffer_type, 0, 0, "tmp_pic");
    rc = matrix_convert(&tmp, &hd, in);
    if (stats)
        for (int c = 0; c < 3; c++) {
            stats[c] = in->stats.estimated_floor[c];
            stats[3 + c] = in->stats.estimated_ceiling[c];
        }
    if (tmp444)
        for (int c = 0; c < 3; c++) memcpy(tmp444 + c * n, tmp.buf[c], n * 2);

    init_pic(out, w, h, out->chroma_format_idc, out->bit_depth, out->video_full_range_flag, out->colour_primaries,
             out->transfer_characteristics, out->matrix_coeffs, 0, PIC_TYPE_U16, 0, 0, "out_pic");
    if (rc == 0) {
        if (out->chroma_format_idc != in->chroma_format_idc) rc = convert(out, &hd, &tmp);
        else for (int c = 0; c < 3; c++) memcpy(out->buf[c], tmp.buf[c], n * 2);
    }
    if (rc == 0) {
        if (tmp.bit_depth - out->bit_depth < 0) rc = 10;          /* write_yuv would exit(0) */
        else {
            char devnull[] = "/dev/null";
            write_yuv(devnull, &hd, out, tmp.bit_depth);
        }
    }
    double t1 = now_s();
    if (seconds) *seconds = t1 - t0;
    if (rc == 0) {
        unsigned short *o = dst_yuv;
        for (int c = 0; c < 3; c++) {
            size_t pn = (size_t)out->plane[c].width * out->plane[c].height;
            memcpy(o, out->buf[c], pn * 2);
            o += pn;
        }
    }
    deinit_pic(in);
    deinit_pic(&tmp);
    deinit_pic(out);
    return rc;
}

/* matrix_convert into an F32 picture (convert.cpp:1117-1122, 1222-1304): what hdr2yuv.cpp:797-823 runs for an .exr / .dpx
 * destination.  cfg as ref_forward_frame (cfg[13], cfg[14] unused); out444: three float planes.  The caller chooses
 * combinations whose tmp depth is defined (an F32 source with an F32 destination makes set_pic_clip shift by 32). */
int ref_matrix_convert_f32out(const int *cfg, const void *const planes[3], float *out444)
{
    Quiet q;
    const int w = cfg[0], h = cfg[1];
    const size_t n = (size_t)w * h;
    static hdr_t hd;
    memset(&hd, 0, sizeof(hd));
    pic_t *in = &hd.in_pic;
    pic_t tmp;
    memset(&tmp, 0, sizeof(tmp));
    const int in_type = cfg[2] ? PIC_TYPE_F32 : PIC_TYPE_U16;
    init_pic(in, w, h, CHROMA_444, cfg[3], cfg[4], cfg[6], cfg[5], cfg[7], 0, in_type, 0, 0, "in_pic");
    for (int c = 0; c < 3; c++) {
        if (cfg[2]) memcpy(in->fbuf[c], planes[c], n * sizeof(float));
        else memcpy(in->buf[c], planes[c], n * sizeof(unsigned short));
    }
    pic_stats(in, &in->stats, 1);
    const int tmp_depth = in_type == PIC_TYPE_F32 ? in->bit_depth : cfg[8];          /* hdr2yuv.cpp:805-808 */
    init_pic(&tmp, w, h, CHROMA_444, tmp_depth, cfg[9], cfg[11], cfg[10], cfg[12], 0, PIC_TYPE_F32, 0, 0, "tmp_pic");
    const int rc = matrix_convert(&tmp, &hd, in);
    if (rc == 0)
        for (int c = 0; c < 3; c++) memcpy(out444 + c * n, tmp.fbuf[c], n * sizeof(float));
    deinit_pic(in);
    deinit_pic(&tmp);
    return rc;
}

void ref_subsample_fir(unsigned short *dst, const unsigned short *src, int w, int h, unsigned long lo, unsigned long hi)
{
    Quiet q;
    Subsample444to420_FIR(dst, const_cast<unsigned short *>(src), (short)w, (short)h, lo, hi);
}

void ref_subsample_box(unsigned short *dst, const unsigned short *src, int w, int h)
{
    Quiet q;
    Subsample444to420_box(dst, const_cast<unsigned short *>(src), (short)w, (short)h, 0, 65535);
}

/* row-major in/out around the reference's column-major pointer arrays */
void ref_upsample_420to444(const unsigned short *src, unsigned short *dst, int w, int h, int fir,
                           unsigned short lo, unsigned short hi)
{
    const int wh = w / 2, hh = h / 2;
    unsigned short **s = (unsigned short **)malloc(wh * sizeof(*s)), **d = (unsigned short **)malloc(w * sizeof(*d));
    for (int x = 0; x < wh; x++) {
        s[x] = (unsigned short *)malloc(hh * 2);
        for (int y = 0; y < hh; y++) s[x][y] = src[(size_t)y * wh + x];
    }
    for (int x = 0; x < w; x++) d[x] = (unsigned short *)malloc(h * 2);
    Subsample420to444(s, d, (short)w, (short)h, (short)fir, lo, hi);
    for (int x = 0; x < w; x++) {
        for (int y = 0; y < h; y++) dst[(size_t)y * w + x] = d[x][y];
        free(d[x]);
    }
    for (int x = 0; x < wh; x++) free(s[x]);
    free(s);
    free(d);
}

/* which: 0 PQ_f 1 PQ_r 2 bt1886_f 3 bt1886_r 4 rho_f 5 rho_r */
void ref_transfer(int which, const float *in, float *out, long n)
{
    for (long i = 0; i < n; i++) {
        float x = in[i];
        switch (which) {
        case 0: out[i] = PQ10000_f(x); break;
        case 1: out[i] = PQ10000_r(x); break;
        case 2: out[i] = bt1886_f(x, 2.4f, 1.0f, 0.0f); break;
        case 3: out[i] = bt1886_r(x, 2.4f, 1.0f, 0.0f); break;
        case 4: out[i] = RHO_GAMMA_f(x); break;
        default: out[i] = RHO_GAMMA_r(x); break;
        }
    }
}

void ref_pic_stats(int is_f32, int bit_depth, int w, int h, const void *const planes[3], int *floor_ceil)
{
    Quiet q;
    pic_t p;
    memset(&p, 0, sizeof(p));
    init_pic(&p, w, h, CHROMA_444, bit_depth, 1, 0, 0, 0, 0, is_f32 ? PIC_TYPE_F32 : PIC_TYPE_U16, 0, 0, "stats");
    for (int c = 0; c < 3; c++) {
        if (is_f32) memcpy(p.fbuf[c], planes[c], (size_t)w * h * 4);
        else memcpy(p.buf[c], planes[c], (size_t)w * h * 2);
    }
    pic_stats(&p, &p.stats, 1);
    for (int c = 0; c < 3; c++) {
        floor_ceil[c] = p.stats.estimated_floor[c];
        floor_ceil[3 + c] = p.stats.estimated_ceiling[c];
    }
    deinit_pic(&p);
}

/* the reference's matrix_inverse (convert.cpp:1320) on U16 4:4:4 planes Y,Cb,Cr -> G,B,R.
 * cfg: [0]w [1]h [2]matrix_coeffs [3]in_bit_depth [4]in_full_range [5]out_bit_depth.  Returns matrix_inverse's rc. */
int ref_matrix_inverse(const int *cfg, const void *const in_planes[3], unsigned short *const out_planes[3])
{
    Quiet q;
    const int w = cfg[0], h = cfg[1];
    const size_t n = (size_t)w * h;
    static hdr_t hd;
    memset(&hd, 0, sizeof(hd));
    pic_t in, out;
    memset(&in, 0, sizeof(in));
    memset(&out, 0, sizeof(out));
    init_pic(&in, w, h, CHROMA_444, cfg[3], cfg[4], 0, 0, cfg[2], 0, PIC_TYPE_U16, 0, 0, "in");
    init_pic(&out, w, h, CHROMA_444, cfg[5], cfg[4], 0, 0, 0, 0, PIC_TYPE_U16, 0, 0, "out");
    for (int c = 0; c < 3; c++) memcpy(in.buf[c], in_planes[c], n * 2);
    const int rc = matrix_inverse(&out, &hd, &in);
    for (int c = 0; c < 3; c++) memcpy(out_planes[c], out.buf[c], n * 2);
    deinit_pic(&in);
    deinit_pic(&out);
    return rc;
}
}
