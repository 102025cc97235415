// h2y_io.cpp -- see h2y_io.h.  Host-side file decoding only; no pixel arithmetic of the hot path lives here.
#include "h2y_io.h"

#include <sys/mman.h>
#include <sys/stat.h>

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>

namespace h2yio {

namespace {

// A file opened for reading is mapped: the decoders address it in place (one scanline chunk at a time) instead of
// paying two system calls and a copy per chunk, which is what bounded the sequence host before (a 4K EXR has 2160 chunks).
struct File {
    FILE *f = nullptr;
    const uint8_t *map = nullptr;
    size_t map_size = 0;
    explicit File(const std::string &p, const char *mode)
    {
        f = fopen(p.c_str(), mode);
        if (f && mode[0] == 'r') {
            struct stat st;
            if (fstat(fileno(f), &st) == 0 && st.st_size > 0) {
                void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fileno(f), 0);
                if (m != MAP_FAILED) { map = (const uint8_t *)m; map_size = (size_t)st.st_size; madvise(m, map_size, MADV_SEQUENTIAL); }
            }
        }
    }
    ~File()
    {
        if (map) munmap(const_cast<uint8_t *>(map), map_size);
        if (f) fclose(f);
    }
    // n bytes at off, in place; nullptr when the range is not inside the file (or the file could not be mapped)
    const uint8_t *ptr_at(size_t n, uint64_t off) const
    {
        if (!map || off > map_size || n > map_size - off) return nullptr;
        return map + off;
    }
    bool read_at(void *dst, size_t n, uint64_t off)
    {
        if (map) {
            const uint8_t *p = ptr_at(n, off);
            if (!p) return false;
            memcpy(dst, p, n);
            return true;
        }
        if (!f || fseeko(f, (off_t)off, SEEK_SET) != 0) return false;
        return fread(dst, 1, n, f) == n;
    }
};

bool fail(std::string *err, const std::string &msg)
{
    if (err) *err = msg;
    return false;
}

inline uint16_t bswap16(uint16_t v) { return (uint16_t)((v >> 8) | (v << 8)); }
inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }

// ---- TIFF --------------------------------------------------------------------------------------------
struct TiffDir {
    bool big_endian = false;
    uint32_t width = 0, height = 0, bits = 0, compression = 1, photometric = 2, spp = 1, rows_per_strip = 0xffffffffu,
             planar = 1;
    std::vector<uint32_t> offsets, counts;
};

struct TiffReader {
    File file;
    TiffDir d;
    explicit TiffReader(const std::string &p) : file(p, "rb") {}
    uint16_t u16(const uint8_t *p) const { uint16_t v; memcpy(&v, p, 2); return d.big_endian ? bswap16(v) : v; }
    uint32_t u32(const uint8_t *p) const { uint32_t v; memcpy(&v, p, 4); return d.big_endian ? bswap32(v) : v; }

    bool values(uint16_t type, uint32_t count, const uint8_t *field, std::vector<uint32_t> *out)
    {
        const size_t esz = type == 3 ? 2 : (type == 4 ? 4 : (type == 1 ? 1 : 0));
        if (!esz || count > (1u << 24)) return false;              // a strip table never needs more than one entry per row
        std::vector<uint8_t> buf(esz * count);
        if (esz * count <= 4) memcpy(buf.data(), field, esz * count);
        else if (!file.read_at(buf.data(), buf.size(), u32(field))) return false;
        out->resize(count);
        for (uint32_t i = 0; i < count; i++)
            (*out)[i] = type == 3 ? u16(&buf[2 * i]) : (type == 4 ? u32(&buf[4 * i]) : buf[i]);
        return true;
    }

    bool parse(std::string *err)
    {
        if (!file.f) return fail(err, "unable to open file");
        uint8_t hdr[8];
        if (!file.read_at(hdr, 8, 0)) return fail(err, "short TIFF header");
        if (hdr[0] == 'I' && hdr[1] == 'I') d.big_endian = false;
        else if (hdr[0] == 'M' && hdr[1] == 'M') d.big_endian = true;
        else return fail(err, "not a TIFF file");
        if (u16(hdr + 2) != 42) return fail(err, "not a classic TIFF (BigTIFF is not supported)");
        const uint32_t ifd = u32(hdr + 4);
        uint8_t nb[2];
        if (!file.read_at(nb, 2, ifd)) return fail(err, "bad IFD offset");
        const int n = u16(nb);
        std::vector<uint8_t> ent((size_t)n * 12);
        if (!file.read_at(ent.data(), ent.size(), ifd + 2)) return fail(err, "short IFD");
        for (int i = 0; i < n; i++) {
            const uint8_t *e = &ent[(size_t)i * 12];
            const uint16_t tag = u16(e), type = u16(e + 2);
            const uint32_t count = u32(e + 4);
            std::vector<uint32_t> v;
            if (tag != 273 && tag != 279 && count > 8) continue;
            if (!values(type, count, e + 8, &v) || v.empty()) continue;
            switch (tag) {
            case 256: d.width = v[0]; break;
            case 257: d.height = v[0]; break;
            case 258: d.bits = v[0]; break;
            case 259: d.compression = v[0]; break;
            case 262: d.photometric = v[0]; break;
            case 273: d.offsets = v; break;
            case 277: d.spp = v[0]; break;
            case 278: d.rows_per_strip = v[0]; break;
            case 279: d.counts = v; break;
            case 284: d.planar = v[0]; break;
            default: break;
            }
        }
        if (!d.width || !d.height) return fail(err, "TIFF without dimensions");
        if (d.width > 65535u || d.height > 65535u) return fail(err, "TIFF dimensions out of range");
        if (d.compression != 1) return fail(err, "compressed TIFF is not supported (baseline raw strips only)");
        if (d.bits != 16) return fail(err, "TIFF BitsPerSample must be 16");
        if (d.spp != 3 && d.spp != 4) return fail(err, "TIFF SamplesPerPixel must be 3 or 4");
        if (d.planar != 1) return fail(err, "planar TIFF is not supported");
        if (d.offsets.empty()) return fail(err, "TIFF without strips");
        if (d.rows_per_strip > d.height) d.rows_per_strip = d.height;
        return true;
    }
};

// ---- EXR ---------------------------------------------------------------------------------------------
struct ExrChannel { std::string name; int type = 1; int xs = 1, ys = 1; };
struct ExrHeader {
    std::vector<ExrChannel> ch;      // file order (alphabetical)
    int compression = 0;
    int xmin = 0, ymin = 0, xmax = -1, ymax = -1;
    int line_order = 0;
    uint64_t table_offset = 0;
    int width() const { return xmax - xmin + 1; }
    int height() const { return ymax - ymin + 1; }
    int lines_per_block() const { return compression == 3 ? 16 : 1; }
};

bool exr_parse(File &f, ExrHeader *h, std::string *err)
{
    if (!f.f) return fail(err, "unable to open file");
    fseeko(f.f, 0, SEEK_END);
    const uint64_t size = (uint64_t)ftello(f.f);
    const size_t hmax = (size_t)std::min<uint64_t>(size, 1u << 20);
    std::vector<uint8_t> b(hmax);
    if (!f.read_at(b.data(), hmax, 0) || hmax < 8) return fail(err, "short EXR file");
    uint32_t magic, ver;
    memcpy(&magic, &b[0], 4);
    memcpy(&ver, &b[4], 4);
    if (magic != 20000630u) return fail(err, "not an OpenEXR file");
    if (ver & 0x200) return fail(err, "tiled EXR is not supported");
    if (ver & 0x1800) return fail(err, "deep / multi-part EXR is not supported");
    size_t p = 8;
    auto cstr = [&](std::string *s) {
        size_t e = p;
        while (e < hmax && b[e]) e++;
        if (e >= hmax) return false;
        s->assign((const char *)&b[p], e - p);
        p = e + 1;
        return true;
    };
    for (;;) {
        std::string name, type;
        if (p >= hmax) return fail(err, "EXR header too large");
        if (b[p] == 0) { p++; break; }
        if (!cstr(&name) || !cstr(&type) || p + 4 > hmax) return fail(err, "bad EXR attribute");
        int32_t asz;
        memcpy(&asz, &b[p], 4);
        p += 4;
        if (asz < 0 || p + (size_t)asz > hmax) return fail(err, "bad EXR attribute size");
        const uint8_t *v = &b[p];
        if (name == "channels") {
            size_t q = 0;
            while (q < (size_t)asz && v[q]) {
                ExrChannel c;
                while (q < (size_t)asz && v[q]) c.name.push_back((char)v[q++]);
                q++;
                if (q + 16 > (size_t)asz) return fail(err, "bad EXR channel list");
                int32_t t, xs, ys;
                memcpy(&t, v + q, 4); memcpy(&xs, v + q + 8, 4); memcpy(&ys, v + q + 12, 4);
                c.type = t; c.xs = xs; c.ys = ys;
                q += 16;
                h->ch.push_back(c);
            }
        } else if (name == "compression" && asz >= 1) h->compression = v[0];
        else if (name == "dataWindow" && asz >= 16) {
            int32_t w[4];
            memcpy(w, v, 16);
            h->xmin = w[0]; h->ymin = w[1]; h->xmax = w[2]; h->ymax = w[3];
        } else if (name == "lineOrder" && asz >= 1) h->line_order = v[0];
        p += (size_t)asz;
    }
    h->table_offset = p;
    if (h->xmax < h->xmin || h->ymax < h->ymin) return fail(err, "EXR without a data window");
    if ((int64_t)h->xmax - h->xmin >= 65535 || (int64_t)h->ymax - h->ymin >= 65535) return fail(err, "EXR data window out of range");
    if (h->compression != 0 && h->compression != 2 && h->compression != 3)
        return fail(err, "EXR compression must be NONE, ZIPS or ZIP");
    for (const auto &c : h->ch) {
        if (c.xs != 1 || c.ys != 1) return fail(err, "subsampled EXR channels are not supported");
        if (c.type != 1 && c.type != 2) return fail(err, "EXR channels must be HALF or FLOAT");
    }
    return true;
}

// ZIP post-processing of OpenEXR: undo the byte-delta predictor, then re-interleave the two byte halves
void exr_unfilter(std::vector<uint8_t> &raw, std::vector<uint8_t> &out)
{
    const size_t n = raw.size();
    for (size_t i = 1; i < n; i++) raw[i] = (uint8_t)(raw[i - 1] + raw[i] - 128);
    out.resize(n);
    const size_t half = (n + 1) / 2;
    for (size_t i = 0, a = 0, b2 = half; i < n;) {
        out[i++] = raw[a++];
        if (i < n) out[i++] = raw[b2++];
    }
}

void exr_filter(const std::vector<uint8_t> &in, std::vector<uint8_t> &out)
{
    const size_t n = in.size();
    out.resize(n);
    const size_t half = (n + 1) / 2;
    for (size_t i = 0, a = 0, b2 = half; i < n;) {
        out[a++] = in[i++];
        if (i < n) out[b2++] = in[i++];
    }
    uint8_t prev = out.empty() ? 0 : out[0];
    for (size_t i = 1; i < n; i++) {
        const uint8_t cur = out[i];
        out[i] = (uint8_t)(cur - prev + 128);
        prev = cur;
    }
}

}   // namespace

uint16_t float_to_half(float f)
{
    // round to nearest even, as Imath's half(float)
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u | (man >> 13) : 0));
    if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        const int shift = 14 - exp;
        uint32_t h = man >> shift;
        const uint32_t rem = man & ((1u << shift) - 1), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (h & 1))) h++;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((uint32_t)exp << 10) | (man >> 13);
    const uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) h++;
    return (uint16_t)(sign | h);
}

// ---- TIFF API ----------------------------------------------------------------------------------------
bool tiff_probe(const std::string &path, ImageInfo *info, std::string *err)
{
    TiffReader r(path);
    if (!r.parse(err)) return false;
    info->width = (int)r.d.width; info->height = (int)r.d.height; info->channels = (int)r.d.spp; info->bits = 16;
    info->is_half = false;
    return true;
}

bool tiff_read(const std::string &path, uint16_t *dst, int crop_w, int crop_h, ImageInfo *info, std::string *err)
{
    TiffReader r(path);
    if (!r.parse(err)) return false;
    const TiffDir &d = r.d;
    const int w = (int)d.width, h = (int)d.height, ch = (int)d.spp;
    const int ow = crop_w > 0 && crop_w < w ? crop_w : w, oh = crop_h > 0 && crop_h < h ? crop_h : h;
    const int x0 = (w - ow) / 2, y0 = (h - oh) / 2;                     // centred cut-out (tiff.cpp:191-220)
    const size_t row_bytes = (size_t)w * ch * 2;
    std::vector<uint8_t> strip;
    const uint32_t rps = d.rows_per_strip;
    // every output row must come from a strip that is really there: a short strip table, RowsPerStrip = 0 or byte counts
    // smaller than the rows they claim would otherwise leave rows of the (uninitialised, pinned) destination as they were
    if (rps == 0) return fail(err, "TIFF RowsPerStrip is 0");
    if ((uint64_t)d.offsets.size() * rps < (uint64_t)h) return fail(err, "TIFF strips do not cover the image");
    if (row_bytes * std::min<uint64_t>(rps, (uint64_t)h) > ((uint64_t)1 << 31)) return fail(err, "TIFF strip too large");
    int rows_done = 0;
    for (size_t s = 0; s < d.offsets.size(); s++) {
        const int first_row = (int)(s * rps), nrows = std::min<int>((int)rps, h - first_row);
        if (nrows <= 0) break;
        if (first_row + nrows <= y0 || first_row >= y0 + oh) continue;
        if (s < d.counts.size() && (uint64_t)d.counts[s] < (uint64_t)row_bytes * nrows) return fail(err, "TIFF strip byte count too small");
        strip.resize(row_bytes * nrows);
        if (!r.file.read_at(strip.data(), strip.size(), d.offsets[s])) return fail(err, "short TIFF strip");
        for (int rr = 0; rr < nrows; rr++) {
            const int y = first_row + rr - y0;
            if (y < 0 || y >= oh) continue;
            const uint16_t *src = reinterpret_cast<const uint16_t *>(&strip[row_bytes * rr]) + (size_t)x0 * ch;
            uint16_t *o = dst + (size_t)y * ow * ch;
            if (d.big_endian) for (int i = 0; i < ow * ch; i++) o[i] = bswap16(src[i]);
            else memcpy(o, src, (size_t)ow * ch * 2);
            rows_done++;
        }
    }
    if (rows_done != oh) return fail(err, "TIFF strips do not cover the image");
    info->width = ow; info->height = oh; info->channels = ch; info->bits = 16; info->is_half = false;
    return true;
}

bool tiff_write_rgb16(const std::string &path, const uint16_t *src, int width, int height, int channels, std::string *err)
{
    File f(path, "wb");
    if (!f.f) return fail(err, "unable to create " + path);
    const uint32_t row_bytes = (uint32_t)width * channels * 2;
    const uint32_t data_off = 8;
    const uint32_t ntags = 10;
    // layout: header | pixel rows | bits array | strip offsets | strip counts | IFD
    const uint32_t bits_off = data_off + row_bytes * (uint32_t)height;
    const uint32_t offs_off = bits_off + 2 * (uint32_t)channels + (channels & 1 ? 2 : 0);
    const uint32_t cnts_off = offs_off + 4 * (uint32_t)height;
    const uint32_t ifd_off = cnts_off + 4 * (uint32_t)height;
    uint8_t hdr[8] = {'I', 'I', 42, 0, 0, 0, 0, 0};
    memcpy(hdr + 4, &ifd_off, 4);
    fwrite(hdr, 1, 8, f.f);
    fwrite(src, 1, (size_t)row_bytes * height, f.f);
    std::vector<uint16_t> bits(channels + (channels & 1), 16);
    fwrite(bits.data(), 2, bits.size(), f.f);
    std::vector<uint32_t> offs(height), cnts(height, row_bytes);
    for (int y = 0; y < height; y++) offs[y] = data_off + row_bytes * (uint32_t)y;
    fwrite(offs.data(), 4, height, f.f);
    fwrite(cnts.data(), 4, height, f.f);
    struct Ent { uint16_t tag, type; uint32_t count, value; };
    const Ent ents[ntags] = {
        {256, 4, 1, (uint32_t)width}, {257, 4, 1, (uint32_t)height}, {258, 3, (uint32_t)channels, bits_off},
        {259, 3, 1, 1}, {262, 3, 1, 2}, {273, 4, (uint32_t)height, offs_off}, {277, 3, 1, (uint32_t)channels},
        {278, 4, 1, 1}, {279, 4, (uint32_t)height, cnts_off}, {284, 3, 1, 1}};
    uint16_t n = ntags;
    fwrite(&n, 2, 1, f.f);
    for (const Ent &e : ents) {
        uint8_t b[12];
        memcpy(b, &e.tag, 2); memcpy(b + 2, &e.type, 2); memcpy(b + 4, &e.count, 4);
        uint32_t v = e.value;
        if (e.type == 3 && e.count == 1) v &= 0xffffu;
        memcpy(b + 8, &v, 4);
        fwrite(b, 1, 12, f.f);
    }
    uint32_t next = 0;
    fwrite(&next, 4, 1, f.f);
    return ferror(f.f) ? fail(err, "write error on " + path) : true;
}

// ---- EXR API -----------------------------------------------------------------------------------------
bool exr_probe(const std::string &path, ImageInfo *info, std::string *err)
{
    File f(path, "rb");
    ExrHeader h;
    if (!exr_parse(f, &h, err)) return false;
    bool has_a = false;
    for (const auto &c : h.ch) has_a |= c.name == "A";
    info->width = h.width(); info->height = h.height(); info->channels = has_a ? 4 : 3; info->bits = 16; info->is_half = true;
    return true;
}

bool exr_read_half(const std::string &path, uint16_t *dst, int out_channels, ImageInfo *info, std::string *err)
{
    File f(path, "rb");
    ExrHeader h;
    if (!exr_parse(f, &h, err)) return false;
    const int w = h.width(), ht = h.height();
    // where each file channel goes in the interleaved output (r,g,b,a = 0,1,2,3); others are skipped
    std::vector<int> dest(h.ch.size(), -1);
    bool have[4] = {false, false, false, false};
    size_t line_bytes = 0;
    for (size_t i = 0; i < h.ch.size(); i++) {
        const std::string &n = h.ch[i].name;
        const int k = n == "R" ? 0 : (n == "G" ? 1 : (n == "B" ? 2 : (n == "A" ? 3 : -1)));
        if (k >= 0 && k < out_channels) { dest[i] = k; have[k] = true; }
        line_bytes += (size_t)w * (h.ch[i].type == 1 ? 2 : 4);
    }
    if (!have[0] || !have[1] || !have[2]) return fail(err, "EXR file needs R, G and B channels");
    if (out_channels == 4 && !have[3])
        for (size_t i = 0; i < (size_t)w * ht; i++) dst[i * 4 + 3] = 0x3C00;     // RgbaInputFile fills A with 1
    const int lpb = h.lines_per_block();
    const int nblocks = (ht + lpb - 1) / lpb;
    std::vector<uint64_t> table(nblocks);
    if (!f.read_at(table.data(), (size_t)nblocks * 8, h.table_offset)) return fail(err, "short EXR offset table");
    std::vector<uint8_t> packed, raw, lines;
    const bool bgr_half = out_channels == 3 && h.ch.size() == 3 && h.ch[0].name == "B" && h.ch[1].name == "G" && h.ch[2].name == "R" &&
                          h.ch[0].type == 1 && h.ch[1].type == 1 && h.ch[2].type == 1;
    for (int b = 0; b < nblocks; b++) {
        int32_t hdr2[2];
        if (!f.read_at(hdr2, 8, table[b])) return fail(err, "bad EXR chunk offset");
        const int y0 = hdr2[0] - h.ymin, psize = hdr2[1];
        const int nl = std::min(lpb, ht - y0);
        if (y0 < 0 || nl <= 0 || psize < 0) return fail(err, "bad EXR chunk");
        const size_t want = line_bytes * nl;
        // a chunk is stored raw when compression does not make it smaller, so it is never larger than the lines it holds
        if ((size_t)psize > want || (h.compression == 0 && (size_t)psize != want)) return fail(err, "bad EXR chunk size");
        const uint8_t *data = f.ptr_at((size_t)psize, table[b] + 8);                // in place when the file is mapped
        if (!data) {
            packed.resize((size_t)psize);
            if (!f.read_at(packed.data(), packed.size(), table[b] + 8)) return fail(err, "short EXR chunk");
            data = packed.data();
        }
        if (h.compression == 0 || (size_t)psize == want) { }                         // stored raw when not smaller
        else {
            const uint8_t *pk = data;
            raw.resize(want);
            uLongf got = (uLongf)want;
            if (uncompress(raw.data(), &got, pk, (uLong)psize) != Z_OK || got != want)
                return fail(err, "EXR zlib stream is corrupt");
            exr_unfilter(raw, lines);
            data = lines.data();
        }
        for (int l = 0; l < nl; l++) {
            const uint8_t *p = data + line_bytes * l;
            uint16_t *o = dst + (size_t)(y0 + l) * w * out_channels;
            if (bgr_half) {
                // the common file: half B, G, R (alphabetical channel order) -> interleaved r, g, b in one pass
                const uint16_t *pb = reinterpret_cast<const uint16_t *>(p), *pg = pb + w, *pr = pg + w;
                int x = 0;
                if ((reinterpret_cast<uintptr_t>(o) & 7) == 0)
                    for (; x + 4 <= w; x += 4) {       // four pixels = three 64-bit stores
                        uint64_t r4, g4, b4;
                        memcpy(&r4, pr + x, 8); memcpy(&g4, pg + x, 8); memcpy(&b4, pb + x, 8);
                        const uint64_t r0 = r4 & 0xffff, r1 = (r4 >> 16) & 0xffff, r2 = (r4 >> 32) & 0xffff, r3 = r4 >> 48;
                        const uint64_t g0 = g4 & 0xffff, g1 = (g4 >> 16) & 0xffff, g2 = (g4 >> 32) & 0xffff, g3 = g4 >> 48;
                        const uint64_t b0 = b4 & 0xffff, b1 = (b4 >> 16) & 0xffff, b2 = (b4 >> 32) & 0xffff, b3 = b4 >> 48;
                        uint64_t *o64 = reinterpret_cast<uint64_t *>(o + 3 * x);
                        o64[0] = r0 | (g0 << 16) | (b0 << 32) | (r1 << 48);
                        o64[1] = g1 | (b1 << 16) | (r2 << 32) | (g2 << 48);
                        o64[2] = b2 | (r3 << 16) | (g3 << 32) | (b3 << 48);
                    }
                for (; x < w; x++) { o[3 * x] = pr[x]; o[3 * x + 1] = pg[x]; o[3 * x + 2] = pb[x]; }
                continue;
            }
            for (size_t c = 0; c < h.ch.size(); c++) {
                const bool is_half = h.ch[c].type == 1;
                if (dest[c] >= 0) {
                    uint16_t *oc = o + dest[c];
                    if (is_half) {
                        const uint16_t *s = reinterpret_cast<const uint16_t *>(p);
                        for (int x = 0; x < w; x++) oc[(size_t)x * out_channels] = s[x];
                    } else {
                        for (int x = 0; x < w; x++) {
                            float v;
                            memcpy(&v, p + 4 * (size_t)x, 4);
                            oc[(size_t)x * out_channels] = float_to_half(v);
                        }
                    }
                }
                p += (size_t)w * (is_half ? 2 : 4);
            }
        }
    }
    info->width = w; info->height = ht; info->channels = out_channels; info->bits = 16; info->is_half = true;
    return true;
}

bool exr_write_half(const std::string &path, const uint16_t *src, int width, int height, int channels, int compression,
                    std::string *err)
{
    if (compression != 0 && compression != 2 && compression != 3) return fail(err, "compression must be 0, 2 or 3");
    File f(path, "wb");
    if (!f.f) return fail(err, "unable to create " + path);
    std::vector<uint8_t> hb;
    auto put = [&](const void *p, size_t n) { hb.insert(hb.end(), (const uint8_t *)p, (const uint8_t *)p + n); };
    auto puts0 = [&](const char *s) { put(s, strlen(s) + 1); };
    auto attr = [&](const char *name, const char *type, const void *v, int32_t n) { puts0(name); puts0(type); put(&n, 4); put(v, (size_t)n); };
    const uint32_t magic = 20000630u, ver = 2;
    put(&magic, 4); put(&ver, 4);
    const char *names = channels == 4 ? "ABGR" : "BGR";        // alphabetical, as the format requires
    std::vector<uint8_t> cl;
    for (const char *c = names; *c; c++) {
        cl.push_back((uint8_t)*c); cl.push_back(0);
        const int32_t t = 1, z = 0, one = 1;
        cl.insert(cl.end(), (const uint8_t *)&t, (const uint8_t *)&t + 4);
        cl.insert(cl.end(), (const uint8_t *)&z, (const uint8_t *)&z + 4);
        cl.insert(cl.end(), (const uint8_t *)&one, (const uint8_t *)&one + 4);
        cl.insert(cl.end(), (const uint8_t *)&one, (const uint8_t *)&one + 4);
    }
    cl.push_back(0);
    attr("channels", "chlist", cl.data(), (int32_t)cl.size());
    const uint8_t comp = (uint8_t)compression, lo = 0;
    attr("compression", "compression", &comp, 1);
    const int32_t win[4] = {0, 0, width - 1, height - 1};
    attr("dataWindow", "box2i", win, 16);
    attr("displayWindow", "box2i", win, 16);
    attr("lineOrder", "lineOrder", &lo, 1);
    const float par = 1.0f, swc[2] = {0.0f, 0.0f}, sww = 1.0f;
    attr("pixelAspectRatio", "float", &par, 4);
    attr("screenWindowCenter", "v2f", swc, 8);
    attr("screenWindowWidth", "float", &sww, 4);
    hb.push_back(0);
    const int lpb = compression == 3 ? 16 : 1, nblocks = (height + lpb - 1) / lpb;
    const size_t line_bytes = (size_t)width * channels * 2;
    std::vector<uint64_t> table(nblocks);
    std::vector<std::vector<uint8_t>> chunks(nblocks);
    uint64_t off = hb.size() + (uint64_t)nblocks * 8;
    const int order3[3] = {2, 1, 0}, order4[4] = {3, 2, 1, 0};           // file channel -> r,g,b,a index
    for (int b = 0; b < nblocks; b++) {
        const int y0 = b * lpb, nl = std::min(lpb, height - y0);
        std::vector<uint8_t> plain(line_bytes * nl);
        for (int l = 0; l < nl; l++)
            for (int c = 0; c < channels; c++) {
                const int k = channels == 4 ? order4[c] : order3[c];
                uint16_t *o = reinterpret_cast<uint16_t *>(&plain[line_bytes * l + (size_t)c * width * 2]);
                const uint16_t *s = src + (size_t)(y0 + l) * width * channels + k;
                for (int x = 0; x < width; x++) o[x] = s[(size_t)x * channels];
            }
        std::vector<uint8_t> &ck = chunks[b];
        std::vector<uint8_t> body;
        if (compression == 0) body = plain;
        else {
            std::vector<uint8_t> filt;
            exr_filter(plain, filt);
            uLongf cap = compressBound((uLong)filt.size());
            body.resize(cap);
            if (compress2(body.data(), &cap, filt.data(), (uLong)filt.size(), 6) != Z_OK) return fail(err, "zlib compress failed");
            body.resize(cap);
            if (body.size() >= plain.size()) body = plain;
        }
        const int32_t y = y0, sz = (int32_t)body.size();
        ck.resize(8 + body.size());
        memcpy(&ck[0], &y, 4); memcpy(&ck[4], &sz, 4); memcpy(&ck[8], body.data(), body.size());
        table[b] = off;
        off += ck.size();
    }
    fwrite(hb.data(), 1, hb.size(), f.f);
    fwrite(table.data(), 8, nblocks, f.f);
    for (auto &ck : chunks) fwrite(ck.data(), 1, ck.size(), f.f);
    return ferror(f.f) ? fail(err, "write error on " + path) : true;
}

// ---- DPX ---------------------------------------------------------------------------------------------
namespace {
uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
uint32_t le32(const uint8_t *p) { return ((uint32_t)p[3] << 24) | ((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0]; }
}   // namespace

bool dpx_probe(const std::string &path, ImageInfo *info, std::string *err)
{
    File f(path, "rb");
    if (!f.f) return fail(err, "Cannot open dpx input file " + path);
    uint8_t h[2048];
    if (!f.read_at(h, sizeof(h), 0)) return fail(err, path + ": shorter than a dpx header");
    const bool be = !memcmp(h, "SDPX", 4), le = !memcmp(h, "XPDS", 4);
    if (!be && !le) return fail(err, path + ": bad magic number in dpx header");
    auto u32 = [&](int off) { return be ? be32(h + off) : le32(h + off); };
    info->big_endian = be;
    info->data_offset = u32(4);
    info->width = (int)(int16_t)u32(772);         // dpx_read narrows both to short (dpx.cpp:296-304)
    info->height = (int)(int16_t)u32(776);
    info->bits = h[803];
    info->channels = 3;
    info->is_half = false;
    if (info->bits != 10 && info->bits != 16 && info->bits != 32) {     // dpx.cpp:316-341
        char msg[512];
        if (info->bits == 12) snprintf(msg, sizeof(msg), " dpx packing of file %s is 12-bit, which is not (yet) supported (although it should be)", path.c_str());
        else snprintf(msg, sizeof(msg), " dpx packing of file %s is %d-bits, which is not supported", path.c_str(), info->bits);
        return fail(err, msg);
    }
    if (info->width < 1 || info->height < 1) return fail(err, path + ": bad dpx picture size");
    return true;
}

bool dpx_read_words(const std::string &path, uint32_t *dst, ImageInfo *info, std::string *err)
{
    if (!dpx_probe(path, info, err)) return false;
    File f(path, "rb");
    if (!f.f) return fail(err, "Cannot open dpx input file " + path);
    const size_t bpp = info->bits == 10 ? 4 : (info->bits == 16 ? 6 : 12);
    if (!f.read_at(dst, (size_t)info->width * info->height * bpp, info->data_offset)) return fail(err, "short read from " + path);
    return true;
}

bool dpx_write_float(const std::string &path, const float *g, const float *b, const float *r, int width, int height, std::string *err)
{
    File f(path, "wb");
    if (!f.f) return fail(err, " Cannot open dpx output file " + path);
    // The header dpx_write_float builds (its INTEL_LE switch is off, so every field is in the machine's order: a
    // little-endian file whose magic reads "XPDS").  Undefined fields are 0xff, then the defined ones are stored.
    std::vector<uint8_t> h(2048, 0);
    static const int undefined[][2] = {{784, 800}, {812, 816}, {852, 892}, {924, 964}, {996, 1036}, {1068, 1108}, {1140, 1180},
                                       {1212, 1252}, {1284, 1324}, {1408, 1432}, {1620, 1644}, {1712, 1732}, {1920, 1972}, {20, 36}};
    for (const auto &u : undefined) memset(h.data() + u[0], 0xff, (size_t)(u[1] - u[0]));
    h[1931] = 0;                                         // byte alignment: defined as 0
    auto put32 = [&](int off, uint32_t v) { memcpy(h.data() + off, &v, 4); };
    auto put16 = [&](int off, uint16_t v) { memcpy(h.data() + off, &v, 2); };
    const uint32_t size = (uint32_t)width * (uint32_t)height * 12u;
    put32(0, 0x53445058u);                               // magic
    put32(4, 2048);                                      // image offset
    memcpy(h.data() + 8, "v2.0", 5);                     // version (sprintf writes the terminator too)
    put32(16, size + 2048);                              // total file size
    put32(660, 0xffffffffu);                             // encryption key
    put16(768, 0); put16(770, 1);                        // orientation, number of elements
    put32(772, (uint32_t)(int)(short)width); put32(776, (uint32_t)(int)(short)height);
    put32(780, 1);                                       // data signed
    h[800] = 50; h[801] = 2; h[802] = 4; h[803] = 32;    // RGB, linear, unspecified colorimetry, 32 bits per element
    put16(804, 0); put16(806, 0);                        // no packing, no run-length encoding
    put32(808, 2048); put32(812, 0); put32(816, 0);      // offset to the pixels, no end-of-line / end-of-image padding
    put32(1424, 0xffffffffu); put32(1428, 0xffffffffu);
    if (fwrite(h.data(), 1, h.size(), f.f) != h.size()) return fail(err, "write error on " + path);
    std::vector<float> row((size_t)width * 3);
    for (int y = 0; y < height; y++) {
        const size_t o = (size_t)y * width;
        for (int x = 0; x < width; x++) { row[3 * (size_t)x] = r[o + x]; row[3 * (size_t)x + 1] = g[o + x]; row[3 * (size_t)x + 2] = b[o + x]; }
        if (fwrite(row.data(), 4, row.size(), f.f) != row.size()) return fail(err, "write error on " + path);
    }
    return true;
}

bool dpx_write_raw(const std::string &path, const void *rgb, int bits, int width, int height, bool big_endian, std::string *err)
{
    if (bits != 16 && bits != 32) return fail(err, "dpx_write_raw: bits must be 16 or 32");
    File f(path, "wb");
    if (!f.f) return fail(err, "unable to create " + path);
    std::vector<uint8_t> h(2048, 0);
    auto put = [&](int off, uint32_t v) {
        for (int i = 0; i < 4; i++) h[off + i] = (uint8_t)(v >> (big_endian ? 24 - 8 * i : 8 * i));
    };
    const size_t bytes = (size_t)width * height * 3 * (bits / 8);
    memcpy(h.data(), big_endian ? "SDPX" : "XPDS", 4);
    put(4, 2048);
    memcpy(h.data() + 8, "V1.0", 4);
    put(16, 2048 + (uint32_t)bytes);
    h[770] = big_endian ? 0 : 1; h[771] = big_endian ? 1 : 0;
    put(772, (uint32_t)width);
    put(776, (uint32_t)height);
    h[800] = 50;
    h[803] = (uint8_t)bits;
    fwrite(h.data(), 1, h.size(), f.f);
    const uint8_t *src = static_cast<const uint8_t *>(rgb);
    const int es = bits / 8;
    std::vector<uint8_t> out(bytes);
    for (size_t i = 0; i < bytes / es; i++)              // samples arrive in the machine's (little-endian) order
        for (int k = 0; k < es; k++) out[i * es + k] = src[i * es + (big_endian ? es - 1 - k : k)];
    fwrite(out.data(), 1, out.size(), f.f);
    return ferror(f.f) ? fail(err, "write error on " + path) : true;
}

bool dpx_write_10bit(const std::string &path, const uint16_t *rgb10, int width, int height, bool big_endian, std::string *err)
{
    File f(path, "wb");
    if (!f.f) return fail(err, "unable to create " + path);
    std::vector<uint8_t> h(2048, 0);
    auto put = [&](int off, uint32_t v) {
        for (int i = 0; i < 4; i++) h[off + i] = (uint8_t)(v >> (big_endian ? 24 - 8 * i : 8 * i));
    };
    memcpy(h.data(), big_endian ? "SDPX" : "XPDS", 4);
    put(4, 2048);                                   // offset to image data
    memcpy(h.data() + 8, "V1.0", 4);
    put(16, 2048 + (uint32_t)width * height * 4);   // file size
    h[768] = 0; h[769] = 0;                         // orientation
    h[770] = big_endian ? 0 : 1; h[771] = big_endian ? 1 : 0;   // number of image elements = 1
    put(772, (uint32_t)width);
    put(776, (uint32_t)height);
    h[800] = 50;                                    // descriptor: RGB
    h[803] = 10;                                    // bit size
    fwrite(h.data(), 1, h.size(), f.f);
    std::vector<uint8_t> row((size_t)width * 4);
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < width; x++) {
            const uint16_t *px = rgb10 + ((size_t)y * width + x) * 3;
            const uint32_t w = ((uint32_t)(px[0] & 1023) << 22) | ((uint32_t)(px[1] & 1023) << 12) | ((uint32_t)(px[2] & 1023) << 2);
            for (int i = 0; i < 4; i++) row[(size_t)x * 4 + i] = (uint8_t)(w >> (big_endian ? 24 - 8 * i : 8 * i));
        }
        fwrite(row.data(), 1, row.size(), f.f);
    }
    return ferror(f.f) ? fail(err, "write error on " + path) : true;
}

bool exr_write_rgba_from_float(const std::string &path, const float *g, const float *b, const float *r, int width, int height,
                               int compression, std::string *err)
{
    std::vector<uint16_t> px((size_t)width * height * 4);
    for (size_t i = 0; i < (size_t)width * height; i++) {
        px[4 * i] = float_to_half(r[i]); px[4 * i + 1] = float_to_half(g[i]); px[4 * i + 2] = float_to_half(b[i]);
        px[4 * i + 3] = 0;                                // write_exr_file leaves alpha at Array2D's zero (exr.cpp:107-124)
    }
    return exr_write_half(path, px.data(), width, height, 4, compression, err);
}

// ---- raw ---------------------------------------------------------------------------------------------
uint64_t file_size(const std::string &path)
{
    File f(path, "rb");
    if (!f.f) return 0;
    fseeko(f.f, 0, SEEK_END);
    return (uint64_t)ftello(f.f);
}

bool file_read_at(const std::string &path, void *dst, size_t bytes, uint64_t offset, std::string *err)
{
    File f(path, "rb");
    if (!f.f) return fail(err, "unable to open file " + path);
    if (!f.read_at(dst, bytes, offset)) return fail(err, "short read from " + path);
    return true;
}

bool rgb_planar_read(const std::string &path, uint16_t *dst, int width, int height, long frame, std::string *err)
{
    const size_t n = (size_t)width * height;
    std::vector<uint16_t> planes(n * 3);
    if (!file_read_at(path, planes.data(), n * 6, (uint64_t)frame * n * 6, err)) return false;
    const uint16_t *r = planes.data(), *g = r + n, *b = g + n;          // file order R,G,B (hdr2yuv.cpp:633-638)
    for (size_t i = 0; i < n; i++) { dst[3 * i] = r[i]; dst[3 * i + 1] = g[i]; dst[3 * i + 2] = b[i]; }
    return true;
}

std::string sequence_name(const std::string &first, int index)
{
    if (first.find('%') != std::string::npos) {
        char buf[4096];
        snprintf(buf, sizeof(buf), first.c_str(), index);
        return buf;
    }
    if (index == 0) return first;
    const size_t slash = first.find_last_of('/');
    const size_t base = slash == std::string::npos ? 0 : slash + 1;
    size_t dot = first.find_last_of('.');
    if (dot == std::string::npos || dot < base) dot = first.size();
    size_t e = dot;
    while (e > base && !(first[e - 1] >= '0' && first[e - 1] <= '9')) e--;
    size_t s = e;
    while (s > base && first[s - 1] >= '0' && first[s - 1] <= '9') s--;
    if (s == e) return first;                                       // no digits: a single file
    const std::string digits = first.substr(s, e - s);
    char buf[64];
    snprintf(buf, sizeof(buf), "%0*lld", (int)digits.size(), atoll(digits.c_str()) + index);
    return first.substr(0, s) + buf + first.substr(e);
}

}   // namespace h2yio
