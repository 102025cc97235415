// h2y_iotool -- test helper for the native readers/writers (no GPU):
//   h2y_iotool write-exr  <out.exr>  <w> <h> <channels> <compression 0|2|3> <raw half samples file>
//   h2y_iotool write-tiff <out.tiff> <w> <h> <channels> <raw u16 samples file>
//   h2y_iotool read-exr   <in.exr>  <out raw> [channels]
//   h2y_iotool read-tiff  <in.tiff> <out raw> [crop_w crop_h]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "h2y_io.h"

static bool slurp(const char *p, std::vector<uint16_t> *v, size_t n)
{
    v->resize(n);
    FILE *f = fopen(p, "rb");
    if (!f) return false;
    const bool ok = fread(v->data(), 2, n, f) == n;
    fclose(f);
    return ok;
}

static bool dump(const char *p, const std::vector<uint16_t> &v)
{
    FILE *f = fopen(p, "wb");
    if (!f) return false;
    const bool ok = fwrite(v.data(), 2, v.size(), f) == v.size();
    fclose(f);
    return ok;
}

int main(int argc, char **argv)
{
    std::string err;
    if (argc >= 8 && !strcmp(argv[1], "write-exr")) {
        const int w = atoi(argv[3]), h = atoi(argv[4]), c = atoi(argv[5]), comp = atoi(argv[6]);
        std::vector<uint16_t> v;
        if (!slurp(argv[7], &v, (size_t)w * h * c)) { printf("cannot read %s\n", argv[7]); return 1; }
        if (!h2yio::exr_write_half(argv[2], v.data(), w, h, c, comp, &err)) { printf("%s\n", err.c_str()); return 1; }
        return 0;
    }
    if (argc >= 7 && !strcmp(argv[1], "write-tiff")) {
        const int w = atoi(argv[3]), h = atoi(argv[4]), c = atoi(argv[5]);
        std::vector<uint16_t> v;
        if (!slurp(argv[6], &v, (size_t)w * h * c)) { printf("cannot read %s\n", argv[6]); return 1; }
        if (!h2yio::tiff_write_rgb16(argv[2], v.data(), w, h, c, &err)) { printf("%s\n", err.c_str()); return 1; }
        return 0;
    }
    if (argc >= 4 && !strcmp(argv[1], "read-exr")) {
        h2yio::ImageInfo info;
        if (!h2yio::exr_probe(argv[2], &info, &err)) { printf("%s\n", err.c_str()); return 1; }
        const int c = argc >= 5 ? atoi(argv[4]) : 3;
        std::vector<uint16_t> v((size_t)info.width * info.height * c);
        if (!h2yio::exr_read_half(argv[2], v.data(), c, &info, &err)) { printf("%s\n", err.c_str()); return 1; }
        printf("%d %d %d\n", info.width, info.height, c);
        return dump(argv[3], v) ? 0 : 1;
    }
    if (argc >= 4 && !strcmp(argv[1], "read-tiff")) {
        h2yio::ImageInfo info;
        if (!h2yio::tiff_probe(argv[2], &info, &err)) { printf("%s\n", err.c_str()); return 1; }
        const int cw = argc >= 6 ? atoi(argv[4]) : 0, chh = argc >= 6 ? atoi(argv[5]) : 0;
        const int ow = cw > 0 && cw < info.width ? cw : info.width, oh = chh > 0 && chh < info.height ? chh : info.height;
        std::vector<uint16_t> v((size_t)ow * oh * info.channels);
        if (!h2yio::tiff_read(argv[2], v.data(), cw, chh, &info, &err)) { printf("%s\n", err.c_str()); return 1; }
        printf("%d %d %d\n", info.width, info.height, info.channels);
        return dump(argv[3], v) ? 0 : 1;
    }
    if (argc >= 7 && !strcmp(argv[1], "write-dpx")) {            // write-dpx out.dpx W H BE(0|1) rgb10.bin (u16 R,G,B codes)
        const int w = atoi(argv[3]), h = atoi(argv[4]), be = atoi(argv[5]);
        std::vector<uint16_t> v;
        if (!slurp(argv[6], &v, (size_t)w * h * 3)) { printf("cannot read %s\n", argv[6]); return 1; }
        if (!h2yio::dpx_write_10bit(argv[2], v.data(), w, h, be != 0, &err)) { printf("%s\n", err.c_str()); return 1; }
        return 0;
    }
    if (argc >= 8 && !strcmp(argv[1], "write-dpx-raw")) {        // write-dpx-raw out.dpx W H BE(0|1) BITS(16|32) rgb.bin (u16 or float R,G,B)
        const int w = atoi(argv[3]), h = atoi(argv[4]), be = atoi(argv[5]), bits = atoi(argv[6]);
        std::vector<uint16_t> v;
        if (!slurp(argv[7], &v, (size_t)w * h * 3 * (bits / 16))) { printf("cannot read %s\n", argv[7]); return 1; }
        if (!h2yio::dpx_write_raw(argv[2], v.data(), bits, w, h, be != 0, &err)) { printf("%s\n", err.c_str()); return 1; }
        return 0;
    }
    if (argc >= 6 && !strcmp(argv[1], "write-dpx-float")) {      // write-dpx-float out.dpx W H planes.bin (float G, B, R planes)
        const int w = atoi(argv[3]), h = atoi(argv[4]);
        std::vector<uint16_t> v;
        if (!slurp(argv[5], &v, (size_t)w * h * 6)) { printf("cannot read %s\n", argv[5]); return 1; }
        const float *p = reinterpret_cast<const float *>(v.data());
        if (!h2yio::dpx_write_float(argv[2], p, p + (size_t)w * h, p + 2 * (size_t)w * h, w, h, &err)) { printf("%s\n", err.c_str()); return 1; }
        return 0;
    }
    if (argc >= 4 && !strcmp(argv[1], "read-dpx")) {             // read-dpx in.dpx words.bin: the stored words, as read
        h2yio::ImageInfo info;
        if (!h2yio::dpx_probe(argv[2], &info, &err)) { printf("%s\n", err.c_str()); return 1; }
        std::vector<uint16_t> v((size_t)info.width * info.height * (info.bits == 10 ? 2 : (info.bits == 16 ? 3 : 6)));
        if (!h2yio::dpx_read_words(argv[2], reinterpret_cast<uint32_t *>(v.data()), &info, &err)) { printf("%s\n", err.c_str()); return 1; }
        printf("%d %d %d %d\n", info.width, info.height, info.bits, info.big_endian ? 1 : 0);
        return dump(argv[3], v) ? 0 : 1;
    }
    printf("usage: see the header of h2y_iotool.cpp\n");
    return 2;
}
