/*
 * Minimal stand-in for libtiff's <tiffio.h>, enough to COMPILE the reference's tiff.cpp /
 * yuv2tiff.cpp for the oracle (libtiff headers are not in this image).
 * TEST INFRASTRUCTURE ONLY.  A "TIFF" opened for writing is a flat file that receives the
 * raw strips back to back (no header, no IFD); files cannot be opened for reading.
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
typedef struct { FILE *fp; } TIFF;

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

static inline TIFF *TIFFOpen(const char *name, const char *mode)
{
    if (mode[0] != 'w') return NULL;
    FILE *fp = fopen(name, "wb");
    if (!fp) return NULL;
    TIFF *t = (TIFF *)malloc(sizeof(TIFF));
    t->fp = fp;
    return t;
}
static inline void TIFFClose(TIFF *t) { if (t) { fclose(t->fp); free(t); } }
static inline int TIFFSetField(TIFF *, uint32, ...) { return 1; }
static inline int TIFFGetField(TIFF *, uint32, ...) { return 0; }
static inline tsize_t TIFFStripSize(TIFF *) { return 0; }
static inline tstrip_t TIFFNumberOfStrips(TIFF *) { return 0; }
static inline tsize_t TIFFReadRawStrip(TIFF *, tstrip_t, tdata_t, tsize_t) { return -1; }
static inline tsize_t TIFFWriteRawStrip(TIFF *t, tstrip_t, tdata_t data, tsize_t n)
{
    return (tsize_t)fwrite(data, 1, (size_t)n, t->fp);
}
static inline tdata_t _TIFFmalloc(tsize_t n) { return malloc((size_t)n); }
static inline void _TIFFfree(tdata_t p) { free(p); }
#endif
