/*
 * Minimal stand-in for libtiff's <tiffio.h>, enough to COMPILE the reference's tiff.cpp /
 * yuv2tiff.cpp for the oracle (libtiff headers are not in this image).
 * TEST INFRASTRUCTURE ONLY.  A "TIFF" opened for writing is a flat file that receives the
 * raw strips back to back (no header, no IFD).  A file opened for reading must be a classic
 * little-endian TIFF with uncompressed strips (what this repo's own writer and the tests
 * produce): the stub parses the first IFD and serves the eight tags get_tiff_info() and
 * read_tiff() ask for (tiff.cpp:34-47, 173) with libtiff's value types, and raw strips.
 */
#ifndef H2Y_STUB_TIFFIO_H
#define H2Y_STUB_TIFFIO_H
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef uint32_t uint32;
typedef uint16_t uint16;
typedef uint32_t tstrip_t;
typedef void *tdata_t;
typedef long tsize_t;
typedef struct {
    FILE *fp;
    int reading;
    uint32_t width, length, rows_per_strip, nstrips;
    uint16_t bits, spp, planar, minv, maxv;
    int has_minv, has_maxv;
    uint32_t *offsets, *counts;
} TIFF;

#define TIFFTAG_IMAGEWIDTH 256
#define TIFFTAG_IMAGELENGTH 257
#define TIFFTAG_BITSPERSAMPLE 258
#define TIFFTAG_PHOTOMETRIC 262
#define TIFFTAG_SAMPLESPERPIXEL 277
#define TIFFTAG_ROWSPERSTRIP 278
#define TIFFTAG_STRIPBYTECOUNTS 279
#define TIFFTAG_MINSAMPLEVALUE 280
#define TIFFTAG_MAXSAMPLEVALUE 281
#define TIFFTAG_PLANARCONFIG 284
#define PLANARCONFIG_CONTIG 1
#define PHOTOMETRIC_RGB 2

static inline uint32_t h2y_stub_rd(FILE *fp, int bytes)
{
    unsigned char b[4] = {0, 0, 0, 0};
    if (fread(b, 1, (size_t)bytes, fp) != (size_t)bytes) return 0;
    return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
}
/* entry value(s) `index` .. of an IFD entry: type 3 = SHORT, 4 = LONG */
static inline uint32_t *h2y_stub_values(FILE *fp, int type, uint32_t count, long entry_pos)
{
    uint32_t *v = (uint32_t *)calloc(count ? count : 1, sizeof(uint32_t));
    const int sz = type == 3 ? 2 : 4;
    long back = ftell(fp);
    if ((long)count * sz <= 4) fseek(fp, entry_pos + 8, SEEK_SET);
    else { fseek(fp, entry_pos + 8, SEEK_SET); uint32_t off = h2y_stub_rd(fp, 4); fseek(fp, (long)off, SEEK_SET); }
    for (uint32_t i = 0; i < count; i++) v[i] = h2y_stub_rd(fp, sz);
    fseek(fp, back, SEEK_SET);
    return v;
}
static inline TIFF *TIFFOpen(const char *name, const char *mode)
{
    TIFF *t = (TIFF *)calloc(1, sizeof(TIFF));
    if (mode[0] == 'w') {
        t->fp = fopen(name, "wb");
        if (!t->fp) { free(t); return NULL; }
        return t;
    }
    t->fp = fopen(name, "rb");
    if (!t->fp) { free(t); return NULL; }
    t->reading = 1;
    t->bits = 1; t->spp = 1; t->planar = 1; t->rows_per_strip = 0xffffffffu;
    if (h2y_stub_rd(t->fp, 2) != 0x4949 || h2y_stub_rd(t->fp, 2) != 42) { fclose(t->fp); free(t); return NULL; }
    fseek(t->fp, (long)h2y_stub_rd(t->fp, 4), SEEK_SET);
    const uint32_t n = h2y_stub_rd(t->fp, 2);
    for (uint32_t e = 0; e < n; e++) {
        const long pos = ftell(t->fp);
        const uint32_t tag = h2y_stub_rd(t->fp, 2), type = h2y_stub_rd(t->fp, 2), count = h2y_stub_rd(t->fp, 4);
        uint32_t *v = h2y_stub_values(t->fp, (int)type, count, pos);
        switch (tag) {
        case TIFFTAG_IMAGEWIDTH: t->width = v[0]; break;
        case TIFFTAG_IMAGELENGTH: t->length = v[0]; break;
        case TIFFTAG_BITSPERSAMPLE: t->bits = (uint16_t)v[0]; break;
        case TIFFTAG_SAMPLESPERPIXEL: t->spp = (uint16_t)v[0]; break;
        case TIFFTAG_ROWSPERSTRIP: t->rows_per_strip = v[0]; break;
        case TIFFTAG_PLANARCONFIG: t->planar = (uint16_t)v[0]; break;
        case TIFFTAG_MINSAMPLEVALUE: t->minv = (uint16_t)v[0]; t->has_minv = 1; break;
        case TIFFTAG_MAXSAMPLEVALUE: t->maxv = (uint16_t)v[0]; t->has_maxv = 1; break;
        case 273: t->offsets = v; t->nstrips = count; v = NULL; break;
        case TIFFTAG_STRIPBYTECOUNTS: t->counts = v; v = NULL; break;
        default: break;
        }
        free(v);
        fseek(t->fp, pos + 12, SEEK_SET);
    }
    if (!t->offsets || !t->counts || !t->width || !t->length) { fclose(t->fp); free(t->offsets); free(t->counts); free(t); return NULL; }
    if (t->rows_per_strip > t->length) t->rows_per_strip = t->length;
    return t;
}
static inline void TIFFClose(TIFF *t) { if (t) { fclose(t->fp); free(t->offsets); free(t->counts); free(t); } }
static inline int TIFFSetField(TIFF *, uint32, ...) { return 1; }
/* libtiff's value types: uint32 for the geometry tags, uint16 for the sample tags, a pointer to the count array */
static inline int TIFFGetField(TIFF *t, uint32 tag, ...)
{
    if (!t || !t->reading) return 0;
    va_list ap;
    va_start(ap, tag);
    void *p = va_arg(ap, void *);
    va_end(ap);
    switch (tag) {
    case TIFFTAG_IMAGEWIDTH: *(uint32_t *)p = t->width; return 1;
    case TIFFTAG_IMAGELENGTH: *(uint32_t *)p = t->length; return 1;
    case TIFFTAG_ROWSPERSTRIP: *(uint32_t *)p = t->rows_per_strip; return 1;
    case TIFFTAG_BITSPERSAMPLE: *(uint16_t *)p = t->bits; return 1;
    case TIFFTAG_SAMPLESPERPIXEL: *(uint16_t *)p = t->spp; return 1;
    case TIFFTAG_PLANARCONFIG: *(uint16_t *)p = t->planar; return 1;
    case TIFFTAG_MINSAMPLEVALUE: if (!t->has_minv) return 0; *(uint16_t *)p = t->minv; return 1;
    case TIFFTAG_MAXSAMPLEVALUE: if (!t->has_maxv) return 0; *(uint16_t *)p = t->maxv; return 1;
    case TIFFTAG_STRIPBYTECOUNTS: *(uint32_t **)p = t->counts; return 1;
    default: return 0;
    }
}
static inline tsize_t TIFFStripSize(TIFF *t)
{
    return t && t->reading ? (tsize_t)t->rows_per_strip * t->width * t->spp * (t->bits / 8) : 0;
}
static inline tstrip_t TIFFNumberOfStrips(TIFF *t) { return t && t->reading ? t->nstrips : 0; }
static inline tsize_t TIFFReadRawStrip(TIFF *t, tstrip_t strip, tdata_t buf, tsize_t size)
{
    if (!t || !t->reading || strip >= t->nstrips) return -1;
    if ((tsize_t)t->counts[strip] < size) size = (tsize_t)t->counts[strip];
    fseek(t->fp, (long)t->offsets[strip], SEEK_SET);
    return (tsize_t)fread(buf, 1, (size_t)size, t->fp);
}
static inline tsize_t TIFFWriteRawStrip(TIFF *t, tstrip_t, tdata_t data, tsize_t n)
{
    return (tsize_t)fwrite(data, 1, (size_t)n, t->fp);
}
static inline tdata_t _TIFFmalloc(tsize_t n) { return malloc((size_t)n); }
static inline void _TIFFfree(tdata_t p) { free(p); }
#endif
