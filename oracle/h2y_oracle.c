/*
 * h2y_oracle.c -- CPU restatement of the hdr2yuv / yuv2tiff per-pixel hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see h2y_oracle.h).  Plain C, row-major planes, written from
 * the reference's arithmetic, not its text; every function cites the file:line it follows.
 * Parity PINNED against the compiled reference (oracle/_ref) and tests/golden/.
 *
 * Numerics: x86-64 SSE2 scalar, round-to-nearest-even, no FMA contraction
 * (gcc -O2 -ffp-contract=off), glibc libm.  Where the reference's C++ picks a float
 * overload of pow/log (both arguments float) this file calls powf/logf explicitly.
 */
#include "h2y_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- C++ -> integer conversions as the x86-64 build of the reference performs them ------- */
/* float -> unsigned int compiles to a 64-bit cvttss2si whose low half is kept. */
static unsigned int f2u(float x) { return (unsigned int)(long long)x; }
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---- common.cpp:300-327 ------------------------------------------------------------------ */
void orc_set_clip(int bit_depth, int full_range, orc_clip_t *c)
{
    c->minCV = 0;
    c->maxCV = (1 << bit_depth) - 1;
    c->Half = (unsigned short)(1 << (bit_depth - 1));
    if (full_range == 0) {
        unsigned short D = (unsigned short)(1 << (bit_depth - 8));
        c->minVR = (unsigned short)(16 * D);
        c->maxVR = (unsigned short)(219 * D + c->minVR);
        c->minVRC = c->minVR;
        c->maxVRC = (unsigned short)(224 * D + c->minVRC);
    } else {
        c->minVR = 0;
        c->maxVR = (unsigned short)c->maxCV;
        c->minVRC = 0;
        c->maxVRC = (unsigned short)c->maxCV;
    }
}

/* ---- common.cpp:191-198 ------------------------------------------------------------------ */
void orc_plane_dims(int w, int h, int chroma, int pw[3], int ph[3])
{
    int ws = chroma == ORC_CHROMA_444 ? 0 : 1;
    int hs = chroma == ORC_CHROMA_420 ? 1 : 0;
    pw[0] = w;
    ph[0] = h;
    pw[1] = pw[2] = w >> ws;
    ph[1] = ph[2] = h >> hs;
}

/* ---- common.cpp:66-168: per-plane extrema -> estimated floor / ceiling ------------------- */
void orc_pic_stats(orc_pic_t *pic, float fmin_out[3], float fmax_out[3])
{
    int pw[3], ph[3];
    orc_plane_dims(pic->width, pic->height, pic->chroma_format_idc, pw, ph);
    for (int c = 0; c < 3; c++) {
        long n = (long)pw[c] * ph[c];
        if (pic->pic_buffer_type == ORC_PIC_U16) {
            unsigned short lo = 65535, hi = 0;
            const uint16_t *p = pic->buf[c];
            for (long i = 0; i < n; i++) {
                if (p[i] < lo) lo = p[i];
                if (p[i] > hi) hi = p[i];
            }
            /* snap cascade, common.cpp:94-106 */
            int D = 1 << (pic->bit_depth - 8);
            int ymax = 219 * D + 16 * D, cmax = 224 * D + 16 * D;
            int ceil_est = hi;
            if (ceil_est < ymax && ceil_est > (ymax * 3) / 4) ceil_est = ymax;
            if (ceil_est < cmax && ceil_est > (cmax * 3) / 4) ceil_est = cmax;
            pic->est_floor[c] = lo;
            pic->est_ceiling[c] = ceil_est;
            if (fmin_out) fmin_out[c] = (float)lo;
            if (fmax_out) fmax_out[c] = (float)hi;
        } else {
            /* note the FLT_MIN (smallest positive) seed of the maximum, common.cpp:119 */
            float lo = FLT_MAX, hi = FLT_MIN;
            const float *p = pic->fbuf[c];
            for (long i = 0; i < n; i++) {
                float s = p[i];
                lo = s < lo ? s : lo;
                hi = s > hi ? s : hi;
            }
            pic->est_floor[c] = (int)lo;      /* truncation, common.cpp:135-136 */
            pic->est_ceiling[c] = (int)hi;
            if (fmin_out) fmin_out[c] = lo;
            if (fmax_out) fmax_out[c] = hi;
        }
    }
}

/* ---- transfer functions, convert.cpp:12-87 (float in/out, double inside) ----------------- */
float orc_pq_eotf(float V)
{
    double vp = pow((double)V, 1.0 / 78.84375);
    double num = fmax(vp - 0.8359375, 0.0);
    double den = 18.8515625 - 18.6875 * vp;
    return (float)pow(num / den, 1.0 / 0.1593017578);
}

float orc_pq_oetf(float L)
{
    double lp = pow((double)L, 0.1593017578);
    double ratio = (0.8359375 + 18.8515625 * lp) / (1 + 18.6875 * lp);
    return (float)pow(ratio, 78.84375);
}

/* a and b of BT.1886 for the fixed call arguments gamma=2.4f, Lw=1, Lb=0 (convert.cpp:1051-1057) */
static void bt1886_ab(float *a, float *b, double *g)
{
    const float gamma = 2.4f, Lw = 1.0f, Lb = 0.0f;
    double ig = 1. / (double)gamma;
    *a = (float)pow(pow((double)Lw, ig) - pow((double)Lb, ig), (double)gamma);
    *b = (float)(pow((double)Lb, ig) / (pow((double)Lw, ig) - pow((double)Lb, ig)));
    *g = (double)gamma;
}

float orc_bt1886_eotf(float V)
{
    float a, b;
    double g;
    bt1886_ab(&a, &b, &g);
    float vb = V + b;
    return (float)((double)a * pow(fmax((double)vb, 0.), g));
}

float orc_bt1886_oetf(float L)
{
    float a, b;
    double g;
    bt1886_ab(&a, &b, &g);
    float la = L / a;
    return (float)(pow(fmax((double)la, 0.), 1. / g) - (double)b);
}

float orc_rho_gamma_eotf(float V)
{
    const float rho = 25.0f, gamma = 2.4f;
    float rv = powf(rho, V);                       /* both arguments float: C++ float overload */
    double base = ((double)rv - 1.0) / ((double)rho - 1.0);
    return (float)pow(base, (double)gamma);
}

float orc_rho_gamma_oetf(float L)
{
    const float rho = 25.0f, gamma = 2.4f;
    double lg = pow((double)L, 1.0 / (double)gamma);
    double num = log(1.0 + ((double)rho - 1.0) * lg);
    return (float)(num / (double)logf(rho));       /* log(float) -> float overload */
}

static int is_gamma_family(int t)
{
    return t == ORC_TRANSFER_BT709 || t == ORC_TRANSFER_BT2020_10 || t == ORC_TRANSFER_BT2020_12 ||
           t == ORC_TRANSFER_BT601;
}

/* steps "to linear" then "to destination transfer" of convert.cpp:1021-1109 */
static float change_transfer(float x, int src_t, int dst_t)
{
    int cur = src_t;
    if (cur != ORC_TRANSFER_LINEAR) {
        if (cur == ORC_TRANSFER_PQ) { x = orc_pq_eotf(x); cur = ORC_TRANSFER_LINEAR; }
        else if (cur == ORC_TRANSFER_RHO_GAMMA) { x = orc_rho_gamma_eotf(x); cur = ORC_TRANSFER_LINEAR; }
        else if (is_gamma_family(cur)) { x = orc_bt1886_eotf(x); cur = ORC_TRANSFER_LINEAR; }
        /* unsupported source transfer: value left untouched (convert.cpp:1058-1062) */
    }
    if (cur == ORC_TRANSFER_LINEAR && dst_t != ORC_TRANSFER_LINEAR) {
        if (dst_t == ORC_TRANSFER_PQ) x = orc_pq_oetf(x);
        else if (dst_t == ORC_TRANSFER_RHO_GAMMA) x = orc_rho_gamma_oetf(x);
        else if (is_gamma_family(dst_t)) x = orc_bt1886_oetf(x);
    }
    return x;
}

/* ---- convert.cpp:879-1315 ---------------------------------------------------------------- */
int orc_matrix_convert(orc_pic_t *out, const orc_pic_t *in)
{
    if (in->chroma_format_idc != ORC_CHROMA_444 || out->chroma_format_idc != ORC_CHROMA_444)
        return 1;                                                  /* convert.cpp:886-890 */

    orc_clip_t clip;
    orc_set_clip(out->bit_depth, out->video_full_range_flag, &clip);

    float P = 0, Q = 0, RR = 0, S = 0;                             /* convert.cpp:911-925 */
    if (out->matrix_coeffs == ORC_MATRIX_YDZDX_Y100) { P = -0.5; Q = 0.491722; RR = 0.5; S = -0.49495; }
    else if (out->matrix_coeffs == ORC_MATRIX_YDZDX_Y500) { P = -0.5; Q = 0.493393; RR = 0.5; S = -0.49602; }

    const int change = in->transfer_characteristics != out->transfer_characteristics;
    float range[3] = {0, 0, 0}, offset[3] = {0, 0, 0};
    if (change)
        for (int c = 0; c < 3; c++) {                              /* convert.cpp:936-940 */
            range[c] = (float)(in->est_ceiling[c] - in->est_floor[c]);
            offset[c] = (float)in->est_floor[c];
        }

    const int passthrough = out->matrix_coeffs == in->matrix_coeffs &&
                            out->colour_primaries == in->colour_primaries;
    const int m = out->matrix_coeffs;
    if (!passthrough && m != ORC_MATRIX_YDZDX && m != ORC_MATRIX_BT2020NC && m != ORC_MATRIX_BT709 &&
        m != ORC_MATRIX_YDZDX_Y100 && m != ORC_MATRIX_YDZDX_Y500 && m != ORC_MATRIX_YUVPRIME2)
        return 2;                                                  /* reference exit(0)s, 1195-1198 */
    if (out->pic_buffer_type != ORC_PIC_U16 && out->pic_buffer_type != ORC_PIC_F32)
        return 3;

    const long n = (long)in->width * in->height;
    float tmpF = 0.0f;
    for (long i = 0; i < n; i++) {
        float v[3];                                                /* G, B, R */
        for (int c = 0; c < 3; c++)
            v[c] = in->pic_buffer_type == ORC_PIC_F32 ? in->fbuf[c][i] : (float)in->buf[c][i];

        if (change) {
            for (int c = 0; c < 3; c++) {
                v[c] = (v[c] - offset[c]) / range[c];              /* convert.cpp:1017-1019 */
                v[c] = change_transfer(v[c], in->transfer_characteristics, out->transfer_characteristics);
            }
            if (out->pic_buffer_type == ORC_PIC_F32) {             /* convert.cpp:1116-1122 */
                for (int c = 0; c < 3; c++) v[c] = v[c] * range[c] + offset[c];
            } else if (out->video_full_range_flag) {               /* 1127-1132 */
                for (int c = 0; c < 3; c++) v[c] = v[c] * clip.maxCV;
            } else if (m == ORC_MATRIX_GBR) {                      /* 1133-1138 */
                for (int c = 0; c < 3; c++) v[c] = v[c] * clip.maxVR + clip.minVR;
            } else {                                               /* 1139-1144 */
                v[0] = v[0] * clip.maxVR + clip.minVR;
                v[1] = v[1] * clip.maxVRC + clip.minVRC;
                v[2] = v[2] * clip.maxVRC + clip.minVRC;
            }
        }
        const float G = v[0], B = v[1], R = v[2];

        if (out->pic_buffer_type == ORC_PIC_U16) {
            unsigned int Y;
            long Cb, Cr;
            if (passthrough) {                                     /* 1159-1166 */
                Y = f2u(G); Cb = f2u(B); Cr = f2u(R);
            } else {
                if (m == ORC_MATRIX_YDZDX) {                       /* 1170-1174 */
                    Y = f2u(G);
                    Cb = (int)(-G / 2.0 + B / 2.0 + 0.5);
                    Cr = (int)(-G / 2.0 + R / 2.0 + 0.5);
                } else if (m == ORC_MATRIX_BT2020NC) {             /* 1176-1180 */
                    tmpF = (float)((0.2627 * R + 0.6780 * G + 0.0593 * B) + 0.5);
                    Y = f2u(tmpF);
                    Cb = (int)((B - tmpF) / 1.8814 + 0.5);
                    Cr = (int)((R - tmpF) / 1.4746 + 0.5);
                } else if (m == ORC_MATRIX_BT709) {                /* 1181-1185 */
                    tmpF = (float)((0.2126 * R + 0.7152 * G + 0.0722 * B) + 0.5);
                    Y = f2u(tmpF);
                    Cb = (int)((B - tmpF) / 1.8556 + 0.5);
                    Cr = (int)((R - tmpF) / 1.5748 + 0.5);
                } else if (m == ORC_MATRIX_YDZDX_Y100 || m == ORC_MATRIX_YDZDX_Y500) { /* 1186-1190 */
                    Y = f2u(G);
                    Cb = (int)(P * G + Q * B + 0.5);
                    Cr = (int)(RR * R + S * G + 0.5);
                } else {                                           /* YUVPRIME2 1191-1194 */
                    Y = f2u(G); Cb = f2u(B); Cr = f2u(R);
                }
                Cb = Cb + clip.Half - 1;                           /* 1200-1201 */
                Cr = Cr + clip.Half - 1;
            }
            /* clamp; chroma compares go through unsigned long (1207-1213) */
            if (Y > clip.maxCV) Y = (unsigned int)clip.maxCV;
            if ((unsigned long)Cb > clip.maxCV) Cb = (long)clip.maxCV;
            if ((unsigned long)Cr > clip.maxCV) Cr = (long)clip.maxCV;
            out->buf[0][i] = (uint16_t)Y;
            out->buf[1][i] = (uint16_t)Cb;
            out->buf[2][i] = (uint16_t)Cr;
        } else {
            float Y, Cb, Cr;
            if (passthrough) { Y = G; Cb = B; Cr = R; }            /* 1239-1245 */
            else {
                if (m == ORC_MATRIX_YDZDX) {
                    Y = G;
                    Cb = (float)(-G / 2.0 + B / 2.0 + 0.5);
                    Cr = (float)(-G / 2.0 + R / 2.0 + 0.5);
                } else if (m == ORC_MATRIX_BT2020NC) {
                    tmpF = (float)((0.2627 * R + 0.6780 * G + 0.0593 * B) + 0.5);
                    Y = tmpF;
                    Cb = (float)((B - tmpF) / 1.8814 + 0.5);
                    Cr = (float)((R - tmpF) / 1.4746 + 0.5);
                } else if (m == ORC_MATRIX_BT709) {
                    tmpF = (float)((0.2126 * R + 0.7152 * G + 0.0722 * B) + 0.5);
                    Y = tmpF;
                    Cb = (float)((B - tmpF) / 1.8556 + 0.5);
                    Cr = (float)((R - tmpF) / 1.5748 + 0.5);
                } else if (m == ORC_MATRIX_YDZDX_Y100 || m == ORC_MATRIX_YDZDX_Y500) {
                    Y = G;
                    Cb = (float)(P * G + Q * B + 0.5);
                    Cr = (float)(RR * R + S * G + 0.5);
                } else { Y = G; Cb = B; Cr = R; }
                Cb = Cb + clip.Half - 1;                           /* float + int, float - int: 1280-1281 */
                Cr = Cr + clip.Half - 1;
            }
            const float hi = (float)clip.maxCV, lo = (float)clip.minCV;
            if (Y > hi) Y = hi;
            if (Y < lo) Y = lo;
            if (Cb > hi) Cb = hi;
            if (Cb < lo) Cb = lo;
            if (Cr > hi) Cr = hi;
            if (Cr < lo) Cr = lo;
            out->fbuf[0][i] = Y;
            out->fbuf[1][i] = Cb;
            out->fbuf[2][i] = Cr;
        }
    }
    return 0;
}

/* ---- FIR 4:4:4 -> 4:2:2 -> 4:2:0, convert.cpp:261-383 ------------------------------------ */
static uint16_t fir_finish(float t, unsigned long minCV, unsigned long maxCV)
{
    if (t > maxCV) t = maxCV;      /* compare/assign through float conversion of the limit */
    if (t < minCV) t = minCV;
    return (uint16_t)t;
}

void orc_subsample_fir_h(uint16_t *d, const uint16_t *src, int w, int h,
                         unsigned long minCV, unsigned long maxCV)
{
    const float k21 = 21.0f / 512.0f, k52 = 52.0f / 512.0f, k159 = 159.0f / 512.0f, k256 = 256.0f / 512.0f;
    const int wh = w >> 1;
    for (int y = 0; y < h; y++) {
        const uint16_t *s = src + (long)y * w;
        for (int x = 0; x < w; x += 2) {
            int l5 = clampi(x - 5, 0, w - 1), l3 = clampi(x - 3, 0, w - 1), l1 = clampi(x - 1, 0, w - 1);
            int r1 = clampi(x + 1, 0, w - 1), r3 = clampi(x + 3, 0, w - 1), r5 = clampi(x + 5, 0, w - 1);
            float t = k21 * ((float)s[l5] + (float)s[r5]) - k52 * ((float)s[l3] + (float)s[r3]) +
                      k159 * ((float)s[l1] + (float)s[r1]) + k256 * (float)s[x];
            t = (float)((double)t + 0.5);
            d[(long)y * wh + (x >> 1)] = fir_finish(t, minCV, maxCV);
        }
    }
}

void orc_subsample_fir(uint16_t *dst, const uint16_t *src, int w, int h,
                       unsigned long minCV, unsigned long maxCV)
{
    const int wh = w >> 1;
    uint16_t *mid = (uint16_t *)malloc((size_t)h * wh * sizeof(uint16_t));
    orc_subsample_fir_h(mid, src, w, h, minCV, maxCV);

    const float k228 = 228.0f / 512.0f, k70 = 70.0f / 512.0f, k37 = 37.0f / 512.0f, k21 = 21.0f / 512.0f,
                k11 = 11.0f / 512.0f, k5 = 5.0f / 512.0f;
    for (int y = 0; y < h; y += 2) {
        const uint16_t *row[12];                      /* rows y-5 .. y+6, replicated at the borders */
        for (int k = 0; k < 12; k++) row[k] = mid + (long)clampi(y - 5 + k, 0, h - 1) * wh;
        for (int x = 0; x < wh; x++) {
            float t = k228 * ((float)row[5][x] + (float)row[6][x]) + k70 * ((float)row[4][x] + (float)row[7][x]) -
                      k37 * ((float)row[3][x] + (float)row[8][x]) - k21 * ((float)row[2][x] + (float)row[9][x]) +
                      k11 * ((float)row[1][x] + (float)row[10][x]) + k5 * ((float)row[0][x] + (float)row[11][x]);
            t = (float)((double)t + 0.5);
            dst[(long)(y >> 1) * wh + x] = fir_finish(t, minCV, maxCV);
        }
    }
    free(mid);
}

/* ---- box, convert.cpp:91-172: truncating mean of each 2x2, walked in 4x4 blocks ----------- */
void orc_subsample_box(uint16_t *dst, const uint16_t *src, int w, int h)
{
    const int wh = w / 2;
    for (int y = 0; y < h; y += 4)
        for (int x = 0; x < w; x += 4)
            for (int by = 0; by < 2; by++)
                for (int bx = 0; bx < 2; bx++) {
                    const uint16_t *p = src + (long)(y + 2 * by) * w + x + 2 * bx;
                    unsigned long sum = (unsigned long)p[0] + p[1] + p[w] + p[w + 1];
                    dst[(long)(y / 2 + by) * wh + x / 2 + bx] = (uint16_t)(sum / 4);
                }
}

/* ---- convert.cpp:513-874 (no Y'u''v'' branch) -------------------------------------------- */
int orc_convert(orc_pic_t *out, const orc_pic_t *in, int resampler)
{
    if (in->pic_buffer_type != ORC_PIC_U16 || out->pic_buffer_type != ORC_PIC_U16) return 1; /* 523-528 */
    orc_clip_t clip;
    orc_set_clip(in->bit_depth, in->video_full_range_flag, &clip);      /* clip of the INPUT pic, 520 */
    const int w = in->width, h = in->height;
    if (out->matrix_coeffs == ORC_MATRIX_YUVPRIME2 && out->chroma_format_idc == ORC_CHROMA_420) {
        /* Y'u''v'' (convert.cpp:533-801).  The 16-bit source planes are Y' (rho-gamma coded), Z, X.  Linear Y from
         * the EOTF, four planes subsampled, then u' = 4X/(X+15Y+3Z), v' = 9Y/(X+15Y+3Z) per 4:2:0 sample.  The
         * reference overwrites u'',v'' with u',v' ("HACK", 732-734), so the Y' plane and the 0.25 floor never enter. */
        if (resampler != 0 && resampler != 1) return 5;            /* neither branch runs: planes stay unwritten */
        const size_t n = (size_t)w * h;
        uint16_t *linY = (uint16_t *)malloc(n * 2), *sY = (uint16_t *)malloc(n * 2), *sZ = (uint16_t *)malloc(n * 2),
                 *sX = (uint16_t *)malloc(n * 2);
        for (size_t i = 0; i < n; i++) {                           /* 561-596 */
            float gamma_Y = ((float)in->buf[0][i]) / 65535.0;      /* float / double -> double -> float */
            float Y_f = orc_rho_gamma_eotf(gamma_Y);
            linY[i] = (unsigned short)(Y_f * 65535.0);
        }
        if (resampler == 1) {                                      /* 600-631 */
            orc_subsample_fir(sY, linY, w, h, clip.minCV, clip.maxCV);
            orc_subsample_fir(sZ, in->buf[1], w, h, clip.minCV, clip.maxCV);
            orc_subsample_fir(sX, in->buf[2], w, h, clip.minCV, clip.maxCV);
        } else {
            orc_subsample_box(sY, linY, w, h);
            orc_subsample_box(sZ, in->buf[1], w, h);
            orc_subsample_box(sX, in->buf[2], w, h);
        }
        const long nc = (long)(w / 2) * (h / 2);
        for (long i = 0; i < nc; i++) {                            /* 662-753 */
            double X = ((double)sX[i]) / 65535.0, Z = ((double)sZ[i]) / 65535.0, Y = ((double)sY[i]) / 65535.0;
            double sum = (X + 15.0 * Y + 3.0 * Z);
            double u_prime = 0.0, v_prime = 0.0;
            if (sum > 0.0) { u_prime = 4.0 * X / sum; v_prime = 9.0 * Y / sum; }
            u_prime = u_prime < 0.0 ? 0.0 : u_prime;
            v_prime = v_prime < 0.0 ? 0.0 : v_prime;
            u_prime = u_prime > 1.0 ? 1.0 : u_prime;
            v_prime = v_prime > 1.0 ? 1.0 : v_prime;
            out->buf[1][i] = (unsigned short)(u_prime * 65535.0);
            out->buf[2][i] = (unsigned short)(v_prime * 65535.0);
        }
        free(linY); free(sY); free(sZ); free(sX);
    } else if (out->chroma_format_idc == ORC_CHROMA_420) {
        for (int c = 1; c < 3; c++) {
            if (resampler == 0) orc_subsample_box(out->buf[c], in->buf[c], w, h);
            else orc_subsample_fir(out->buf[c], in->buf[c], w, h, clip.minCV, clip.maxCV);
        }
    } else if (out->chroma_format_idc == ORC_CHROMA_444) {
        for (int c = 1; c < 3; c++) memcpy(out->buf[c], in->buf[c], (size_t)w * h * 2);
    } else if (out->chroma_format_idc == ORC_CHROMA_422) {
        /* The reference has no 4:2:2 branch; defined here as FIR stage 1 only. */
        if (resampler == 0) return 4;
        for (int c = 1; c < 3; c++) orc_subsample_fir_h(out->buf[c], in->buf[c], w, h, clip.minCV, clip.maxCV);
    }
    memcpy(out->buf[0], in->buf[0], (size_t)w * h * 2);                 /* 857-859 */
    return 0;
}

/* ---- tiff.cpp:457-550 -------------------------------------------------------------------- */
int orc_write_yuv_clamp(orc_pic_t *pic, int src_bit_depth)
{
    const int shift = src_bit_depth - pic->bit_depth;
    if (shift < 0) return 1;                                            /* tiff.cpp:396-401 */
    orc_clip_t clip;
    orc_set_clip(pic->bit_depth, pic->video_full_range_flag, &clip);
    int pw[3], ph[3];
    orc_plane_dims(pic->width, pic->height, pic->chroma_format_idc, pw, ph);
    for (int c = 0; c < 3; c++) {
        const unsigned short lo = c == 0 ? clip.minVR : clip.minVRC;
        const unsigned short hi = c == 0 ? clip.maxVR : clip.maxVRC;
        long n = (long)pw[c] * ph[c];
        for (long i = 0; i < n; i++) {
            unsigned short v = (unsigned short)(pic->buf[c][i] >> shift);
            if (pic->video_full_range_flag == 0) {
                v = v < lo ? lo : v;
                v = v > hi ? hi : v;
            } else {
                v = v > clip.maxCV ? (unsigned short)clip.maxCV : v;
            }
            pic->buf[c][i] = v;
        }
    }
    return 0;
}

/* ---- whole chain, hdr2yuv.cpp:791-930 ---------------------------------------------------- */
int orc_forward_frame(orc_pic_t *in, const orc_dst_t *d, uint16_t *dst_yuv)
{
    const int w = in->width, h = in->height;
    const size_t n = (size_t)w * h;
    orc_pic_stats(in, NULL, NULL);                                      /* 797 */

    orc_pic_t tmp = *in, out = *in;
    /* destination of a .yuv is U16 (hdr2yuv.cpp:421-424); tmp depth rule 803-808 */
    tmp.bit_depth = in->pic_buffer_type == ORC_PIC_U16 ? in->bit_depth : d->dst_bit_depth;
    tmp.pic_buffer_type = ORC_PIC_U16;
    tmp.video_full_range_flag = d->dst_full_range;
    tmp.colour_primaries = d->dst_primaries;
    tmp.transfer_characteristics = d->dst_transfer;
    tmp.matrix_coeffs = d->dst_matrix;
    out = tmp;
    out.bit_depth = d->dst_bit_depth;
    out.chroma_format_idc = d->dst_chroma;
    int rc = 0;
    for (int c = 0; c < 3; c++) {
        tmp.buf[c] = (uint16_t *)malloc(n * 2);
        out.buf[c] = (uint16_t *)malloc(n * 2);
        tmp.fbuf[c] = out.fbuf[c] = NULL;
    }
    rc = orc_matrix_convert(&tmp, in);                                  /* 821 */
    if (rc == 0) {
        if (out.chroma_format_idc != in->chroma_format_idc) rc = orc_convert(&out, &tmp, d->resampler); /* 868-897 */
        else for (int c = 0; c < 3; c++) memcpy(out.buf[c], tmp.buf[c], n * 2);                 /* 913-921 */
    }
    if (rc == 0) rc = orc_write_yuv_clamp(&out, tmp.bit_depth) ? 10 : 0;  /* 928 */
    if (rc == 0) {
        int pw[3], ph[3];
        orc_plane_dims(w, h, out.chroma_format_idc, pw, ph);
        uint16_t *o = dst_yuv;
        for (int c = 0; c < 3; c++) {
            memcpy(o, out.buf[c], (size_t)pw[c] * ph[c] * 2);
            o += (size_t)pw[c] * ph[c];
        }
    }
    for (int c = 0; c < 3; c++) { free(tmp.buf[c]); free(out.buf[c]); }
    return rc;
}

/* ---- input staging: tiff.cpp:265-315, exr.cpp:209-235 ------------------------------------ */
void orc_load_rgb16(const uint16_t *rgb, int npix, int nchan, int full_range, uint16_t *planes[3])
{
    orc_clip_t clip;
    orc_set_clip(16, full_range, &clip);
    for (int i = 0; i < npix; i++) {
        unsigned int v[3] = { rgb[(long)i * nchan + 1], rgb[(long)i * nchan + 2], rgb[(long)i * nchan] }; /* G,B,R */
        for (int c = 0; c < 3; c++) {
            if (full_range == 0) {
                if (v[c] < clip.minVR) v[c] = clip.minVR;
                if (v[c] > clip.maxVR) v[c] = clip.maxVR;
            }
            planes[c][i] = (uint16_t)v[c];
        }
    }
}

float orc_half_to_float(uint16_t h)
{
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 1023u, bits;
    if (e == 0) {
        if (m == 0) bits = sign;
        else {
            int sh = 0;
            while (!(m & 1024u)) { m <<= 1; sh++; }
            bits = sign | ((uint32_t)(113 - sh) << 23) | ((m & 1023u) << 13);
        }
    } else if (e == 31) bits = sign | 0x7f800000u | (m << 13);
    else bits = sign | ((e + 112u) << 23) | (m << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

void orc_load_half(const uint16_t *px, int npix, int nchan, float *planes[3])
{
    for (int i = 0; i < npix; i++) {                 /* Rgba order r,g,b,a -> planes G,B,R */
        planes[0][i] = orc_half_to_float(px[(long)i * nchan + 1]);
        planes[1][i] = orc_half_to_float(px[(long)i * nchan + 2]);
        planes[2][i] = orc_half_to_float(px[(long)i * nchan]);
    }
}

/* ---- yuv2tiff.cpp:575-692 on row-major planes --------------------------------------------- */
static uint16_t up_finish(float t, unsigned short minCV, unsigned short maxCV)
{
    if (t > maxCV) t = maxCV;
    if (t < minCV) t = minCV;
    return (uint16_t)t;
}

void orc_upsample_420to444(const uint16_t *src, uint16_t *dst, int w, int h, int fir,
                           unsigned short minCV, unsigned short maxCV)
{
    const int wh = w / 2, hh = h / 2;
    if (!fir) {                                                         /* 577-588 */
        for (int y = 0; y < hh; y++)
            for (int x = 0; x < wh; x++) {
                uint16_t v = src[(long)y * wh + x];
                dst[(long)(2 * y) * w + 2 * x] = dst[(long)(2 * y) * w + 2 * x + 1] = v;
                dst[(long)(2 * y + 1) * w + 2 * x] = dst[(long)(2 * y + 1) * w + 2 * x + 1] = v;
            }
        return;
    }
    uint16_t *mid = (uint16_t *)malloc((size_t)wh * h * 2);             /* 4:2:2, wh x h */
    const float a3 = 3.0f / 256.0f, a16 = 16.0f / 256.0f, a67 = 67.0f / 256.0f, a227 = 227.0f / 256.0f,
                a32 = 32.0f / 256.0f, a7 = 7.0f / 256.0f;
    for (int j = 0; j < hh; j++) {                                      /* vertical 2-phase, 615-650 */
        const uint16_t *m3 = src + (long)clampi(j - 3, 0, hh - 1) * wh, *m2 = src + (long)clampi(j - 2, 0, hh - 1) * wh,
                       *m1 = src + (long)clampi(j - 1, 0, hh - 1) * wh, *c0 = src + (long)j * wh,
                       *p1 = src + (long)clampi(j + 1, 0, hh - 1) * wh, *p2 = src + (long)clampi(j + 2, 0, hh - 1) * wh,
                       *p3 = src + (long)clampi(j + 3, 0, hh - 1) * wh;
        for (int x = 0; x < wh; x++) {
            float t = a3 * (float)m3[x] - a16 * (float)m2[x] + a67 * (float)m1[x] + a227 * (float)c0[x] -
                      a32 * (float)p1[x] + a7 * (float)p2[x];
            t = (float)((double)t + 0.5);
            mid[(long)(2 * j) * wh + x] = up_finish(t, minCV, maxCV);
            t = a3 * (float)p3[x] - a16 * (float)p2[x] + a67 * (float)p1[x] + a227 * (float)c0[x] -
                a32 * (float)m1[x] + a7 * (float)m2[x];
            t = (float)((double)t + 0.5);
            mid[(long)(2 * j + 1) * wh + x] = up_finish(t, minCV, maxCV);
        }
    }
    const float b21 = 21.0f / 256.0f, b52 = 52.0f / 256.0f, b159 = 159.0f / 256.0f;
    for (int y = 0; y < h; y++) {                                       /* horizontal, 661-687 */
        const uint16_t *r = mid + (long)y * wh;
        for (int i = 0; i < wh; i++) {
            int l2 = clampi(i - 2, 0, wh - 1), l1 = clampi(i - 1, 0, wh - 1), r1 = clampi(i + 1, 0, wh - 1),
                r2 = clampi(i + 2, 0, wh - 1), r3 = clampi(i + 3, 0, wh - 1);
            dst[(long)y * w + 2 * i] = r[i];
            float t = b21 * ((float)r[l2] + (float)r[r3]) - b52 * ((float)r[l1] + (float)r[r2]) +
                      b159 * ((float)r[i] + (float)r[r1]);
            t = (float)((double)t + 0.5);
            dst[(long)y * w + 2 * i + 1] = up_finish(t, minCV, maxCV);
        }
    }
    free(mid);
}

/* ---- yuv2tiff.cpp:278-552: one frame ------------------------------------------------------ */
long orc_yuv2tiff_frame(const orc_inv_params *p, const uint16_t *yuv, uint16_t *rgb)
{
    const int w = p->width, h = p->height, wh = w / 2, hh = h / 2;
    const size_t n = (size_t)w * h, nc = (size_t)wh * hh;
    /* keyword table yuv2tiff.cpp:89-100, 137-157 */
    const int SR = 16 - p->bit_depth;
    const unsigned short Half = (unsigned short)(1 << (p->bit_depth - 1));
    const unsigned short Full = (unsigned short)(1 << p->bit_depth);
    const unsigned short minCV = 0, maxCV = (unsigned short)(Full - 1);
    const unsigned short D = (unsigned short)(1 << (p->bit_depth - 10));
    unsigned short minVR = 0, maxVR = 0, minVRC = 0, maxVRC = 0;
    if (!p->full_range) {                                               /* 178-186 */
        minVR = (unsigned short)(64 * D);
        maxVR = (unsigned short)(876 * D + minVR);
        minVRC = minVR;
        maxVRC = (unsigned short)(896 * D + minVRC);
    }
    float T = 0, U = 0, V = 0, W = 0;                                   /* 207-219 */
    if (p->matrix == ORC_INV_Y100) { T = 0.98989899; U = 2.0; V = 1.016835017; W = 2.03367; }
    else if (p->matrix == ORC_INV_Y500) { T = 0.99203764; U = 2.0; V = 1.013391241; W = 2.026782; }

    uint16_t *Yp = (uint16_t *)malloc(n * 2), *cb = (uint16_t *)malloc(nc * 2), *cr = (uint16_t *)malloc(nc * 2);
    uint16_t *cb4 = (uint16_t *)malloc(n * 2), *cr4 = (uint16_t *)malloc(n * 2);
    for (size_t i = 0; i < n; i++) {                                    /* 283-294 */
        uint16_t v = yuv[i];
        if (!p->full_range) { v = v < minVR ? minVR : v; v = v > maxVR ? maxVR : v; }
        Yp[i] = v;
    }
    for (size_t i = 0; i < nc; i++) {                                   /* 297-320 */
        uint16_t a = yuv[n + i], b = yuv[n + nc + i];
        if (!p->full_range) {
            a = a < minVRC ? minVRC : a; a = a > maxVRC ? maxVRC : a;
            b = b < minVRC ? minVRC : b; b = b > maxVRC ? maxVRC : b;
        }
        cb[i] = a; cr[i] = b;
    }
    orc_upsample_420to444(cb, cb4, w, h, p->fir, minCV, maxCV);         /* 341-342 */
    orc_upsample_420to444(cr, cr4, w, h, p->fir, minCV, maxCV);

    const int nch = p->alpha ? 4 : 3;
    long invalid = 0;
    float tmpF;
    for (size_t i = 0; i < n; i++) {
        int Yav = Yp[i];
        const int Ysave = Yav;
        int Rp, Bp;
        const float fcb = (float)cb4[i], fcr = (float)cr4[i];
        const double top = Full - 1.0;
        if (p->matrix == ORC_INV_YDZDX && p->ybar) {                    /* -X: 365-399 */
            /* mean of the 2x2 luma block this pixel belongs to, clamped to [1, Full-1] */
            const size_t x = i % (size_t)w, y = i / (size_t)w, bx = x & ~(size_t)1, by = y & ~(size_t)1;
            int Ybar = (int)Yp[by * w + bx] + (int)Yp[by * w + bx + 1] + (int)Yp[(by + 1) * w + bx] + (int)Yp[(by + 1) * w + bx + 1];
            Ybar = Ybar / 4;
            if (Ybar < 1) Ybar = 1;
            if (Ybar > (Full - 1)) Ybar = Full - 1;
            float RED, BLUE;
            RED = (2.0 * (float)(cr4[i]) - (float)(Full - 1.0) + (float)Ybar) * ((float)Yav) / ((float)Ybar);
            BLUE = (2.0 * (float)(cb4[i]) - (float)(Full - 1.0) + (float)Ybar) * ((float)Yav) / ((float)Ybar);
            if (RED > (Full - 1.5)) RED = Full - 1.0;
            if (BLUE > (Full - 1.5)) BLUE = Full - 1.0;
            Rp = RED;
            Bp = BLUE;
        } else if (p->matrix == ORC_INV_YDZDX) {                        /* 399-402 */
            Rp = 2 * (int)cr4[i] - (int)top + Yav;
            Bp = 2 * (int)cb4[i] - (int)top + Yav;
        } else if (p->matrix == ORC_INV_2020 || p->matrix == ORC_INV_709) {   /* 403-423 */
            const int is2020 = p->matrix == ORC_INV_2020;
            const double kb = is2020 ? 1.8814 : 1.8556, kr = is2020 ? 1.4746 : 1.5748;
            const double wb = is2020 ? 0.0593 : 0.07222, wr = is2020 ? 0.2627 : 0.2126, wg = is2020 ? 0.6780 : 0.7152;
            tmpF = (float)((fcb - (Half - 0.5)) * kb + Yav);
            if (tmpF > top) tmpF = (float)top;
            Bp = (int)tmpF;
            tmpF = (float)((fcr - (Half - 0.5)) * kr + Yav);
            if (tmpF > top) tmpF = (float)top;
            Rp = (int)tmpF;
            tmpF = (float)(((float)Yav - wb * (float)Bp - wr * (float)Rp) / wg + 0.5);
            if (tmpF > top) tmpF = (float)top;
            Yav = (int)tmpF;
        } else {                                                        /* Y100 / Y500, 424-430 */
            tmpF = (float)((fcb - (Half - 0.5)) * W + ((float)Yav) * V);
            if (tmpF > top) tmpF = (float)top;
            Bp = (int)tmpF;
            tmpF = (float)((fcr - (Half - 0.5)) * U + ((float)Yav) * T);
            if (tmpF > top) tmpF = (float)top;
            Rp = (int)tmpF;
        }
        if (Yav < 0) { Yav = 0; invalid++; }                            /* 478-483 */
        if (Rp < 0) { Rp = 0; if (Ysave != 0) invalid++; }              /* 485-497 */
        if (Bp < 0) { Bp = 0; if (Ysave != 0) invalid++; }              /* 499-511 */
        if (!p->full_range) {                                           /* 515-524 */
            Rp = clampi(Rp, minVR, maxVR);
            Yav = clampi(Yav, minVR, maxVR);
            Bp = clampi(Bp, minVR, maxVR);
        }
        uint16_t *o = rgb + i * nch;                                    /* 535-544 */
        o[0] = (uint16_t)(((unsigned short)Rp) << SR);
        o[1] = (uint16_t)(((unsigned short)Yav) << SR);
        o[2] = (uint16_t)(((unsigned short)Bp) << SR);
        if (p->alpha) o[3] = 65535;
    }
    free(Yp); free(cb); free(cr); free(cb4); free(cr4);
    return invalid;
}

/* ---- convert.cpp:1320-1867 matrix_inverse (U16 pictures, 4:4:4) + tiff.cpp:559-652 write_tiff ------ */
/* Returns the invalidPixels count, or -2 for "Can't determine color difference to use?" (convert.cpp:1735-1738).
 * Quirks kept: Half / Full are hard-coded 2048 / 4096 (1337-1338); D2020 tests MATRIX_BT2020c = 10 (1325);
 * DXYZ is cleared by comparing matrix_coeffs with the BOOLEANS D709 / D2020 / Y100 / Y500 (1387), so only
 * matrix_coeffs == 1 takes the 709 equations, matrix_coeffs == 0 is the error, everything else is Y'DzDx;
 * Bp / Rp stay float (not truncated) inside the green equation; FULLRANGE is hard-coded 0 (1343) so the
 * clamp always uses in_pic->clip; the final shift is by |in.bit_depth - out.bit_depth| (1796-1809). */
long orc_matrix_inverse(uint16_t *out_planes[3], const uint16_t *in_planes[3], int w, int h, int matrix_coeffs,
                        int in_bit_depth, int in_full_range, int out_bit_depth)
{
    const short D709 = matrix_coeffs == ORC_MATRIX_BT709;
    const short D2020 = matrix_coeffs == ORC_MATRIX_BT2020C;
    const short Y500 = 0, Y100 = 0;
    short DXYZ = 1;
    const unsigned short Half = 2048, Full = 4096;
    if (matrix_coeffs == D709 || matrix_coeffs == D2020 || matrix_coeffs == Y100 || matrix_coeffs == Y500) DXYZ = 0;
    orc_clip_t clip;
    orc_set_clip(in_bit_depth, in_full_range, &clip);
    long invalid = 0;
    const long n = (long)w * h;
    for (long i = 0; i < n; i++) {
        float tmpF, Yav = (float)in_planes[0][i], Cb = (float)in_planes[1][i], Cr = (float)in_planes[2][i], Rp, Bp;
        if (DXYZ) {                                                     /* 1707-1710 */
            Rp = (float)(2.0 * Cr - (Full - 1.0) + Yav);
            Bp = (float)(2.0 * Cb - (Full - 1.0) + Yav);
        } else if (D2020) {                                             /* 1711-1721 (unreachable: see above) */
            tmpF = (float)(((float)(Cb) - (Half - 0.5)) * 1.8814 + Yav);
            if (tmpF > (Full - 1.0)) tmpF = (float)(Full - 1.0);
            Bp = tmpF;
            tmpF = (float)(((float)(Cr) - (Half - 0.5)) * 1.4746 + Yav);
            if (tmpF > (Full - 1.0)) tmpF = (float)(Full - 1.0);
            Rp = tmpF;
            tmpF = (float)(((float)Yav - 0.0593 * (float)Bp - 0.2627 * (float)Rp) / 0.6780 + 0.5);
            if (tmpF > (Full - 1.0)) tmpF = (float)(Full - 1.0);
            Yav = tmpF;
        } else if (D709) {                                              /* 1722-1733 */
            tmpF = (float)(((float)(Cb) - (Half - 0.5)) * 1.8556 + Yav);
            if (tmpF > (Full - 1.0)) tmpF = (float)(Full - 1.0);
            Bp = tmpF;
            tmpF = (float)(((float)(Cr) - (Half - 0.5)) * 1.5748 + Yav);
            if (tmpF > (Full - 1.0)) tmpF = (float)(Full - 1.0);
            Rp = tmpF;
            tmpF = (float)(((float)Yav - 0.07222 * (float)Bp - 0.2126 * (float)Rp) / 0.7152 + 0.5);
            if (tmpF > (Full - 1.0)) tmpF = (float)(Full - 1.0);
            Yav = tmpF;
        } else {
            return -2;
        }
        int G = (int)Yav, B = (int)Bp, R = (int)Rp;                     /* 1758-1779 */
        if (G < 0) { G = 0; invalid++; }
        if (R < 0) { invalid++; R = 0; }
        if (B < 0) { invalid++; B = 0; }
        R = (R < clip.minVR) ? clip.minVR : R;                          /* 1783-1793, FULLRANGE == 0 */
        G = (G < clip.minVR) ? clip.minVR : G;
        B = (B < clip.minVR) ? clip.minVR : B;
        R = (R > clip.maxVR) ? clip.maxVR : R;
        G = (G > clip.maxVR) ? clip.maxVR : G;
        B = (B > clip.maxVR) ? clip.maxVR : B;
        if (in_bit_depth > out_bit_depth) {                             /* 1796-1809 */
            const int shift = in_bit_depth - out_bit_depth;
            R >>= shift; G >>= shift; B >>= shift;
        } else {
            const int shift = out_bit_depth - in_bit_depth;
            R <<= shift; G <<= shift; B <<= shift;
        }
        out_planes[0][i] = (uint16_t)G;                                 /* 1816-1818 */
        out_planes[1][i] = (uint16_t)B;
        out_planes[2][i] = (uint16_t)R;
    }
    return invalid;
}

/* tiff.cpp:605-628: planes G,B,R -> interleaved R,G,B rows, each sample << (pic.bit_depth - src_bit_depth) */
void orc_write_tiff_rows(uint16_t *rgb, const uint16_t *planes[3], long npix, int pic_bit_depth, int src_bit_depth)
{
    const short SR = (short)(pic_bit_depth - src_bit_depth);
    for (long i = 0; i < npix; i++) {
        int G = planes[0][i], B = planes[1][i], R = planes[2][i];
        R = R << SR; G = G << SR; B = B << SR;
        rgb[3 * i + 0] = (unsigned short)R;
        rgb[3 * i + 1] = (unsigned short)G;
        rgb[3 * i + 2] = (unsigned short)B;
    }
}
