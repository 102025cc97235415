// yuv2tiff -- command-line host of the inverse path with the reference's keyword surface
// (yuv2tiff.cpp:106-198): yuv2tiff <in.yuv> [B10|B14] [709|2020|Y100|Y500] [HD1920|HD960] [BOX] [FULL] [ALPHA]
// [-f frames] [-I] [-X] [YUVPRIME2].  Frames of planar 4:2:0 u16 go to the GPU through h2y_inverse_host; each frame is written
// as tifXYZ/XpYpZp%05d.tif (16-bit RGB, one strip per row), as the reference does (yuv2tiff.cpp:322-342).
//
// Beyond the reference: -d N shards the frames over N GPUs, -o DIR changes the output directory.  The
// reference's while(yuvIn) loop emits one extra garbage frame at end of file (yuv2tiff.cpp:278, 562-565);
// this host stops at the last complete frame.
#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "h2y_io.h"
#include "hdr2yuv_b200.h"

int main(int argc, char *argv[])
{
    if (argc < 2 || !strcmp(argv[1], "-h")) {
        printf("\n ARGS:\n 709 (use Rec709 Color Dif)\n 2020 (use Rec2020 Color Dif)\n HD1920 (Format is 1920x1080 images)\n"
               " HD960 (Format is 960x540 images)\n B10 / B14 (bit depth, default 12)\n Y100 / Y500\n BOX FULL ALPHA\n -f N (frames)\n"
               " -d N (GPUs)  -o DIR (output directory, default tifXYZ)\n"
               " (no args get Y'DzDx color difference and 3840x2160 cutout)\n\n\n");
        return 0;
    }
    h2y_inverse_params ip;
    memset(&ip, 0, sizeof(ip));
    ip.width = 3840; ip.height = 2160; ip.bit_depth = 12; ip.matrix = H2Y_INV_YDzDx; ip.fir = 1;
    int frames = -999, devices = 1, ipixf = 0;
    std::string outdir = "tifXYZ";
    for (int arg = 2; arg < argc; arg++) {
        const char *k = argv[arg];
        if (!strcmp(k, "BOX")) ip.fir = 0;
        else if (!strcmp(k, "FULL")) ip.full_range = 1;
        else if (!strcmp(k, "ALPHA")) ip.alpha = 1;
        else if (!strcmp(k, "709")) ip.matrix = H2Y_INV_709;
        else if (!strcmp(k, "2020")) ip.matrix = H2Y_INV_2020;
        else if (!strcmp(k, "Y100")) ip.matrix = H2Y_INV_Y100;
        else if (!strcmp(k, "Y500")) ip.matrix = H2Y_INV_Y500;
        else if (!strcmp(k, "HD1920")) { ip.width = 1920; ip.height = 1080; }
        else if (!strcmp(k, "HD960")) { ip.width = 960; ip.height = 540; }
        else if (!strcmp(k, "B10")) { ip.bit_depth = 10; printf("\nprocessing line data 10 bits\n"); }
        else if (!strcmp(k, "B14")) { ip.bit_depth = 14; printf("\nprocessing line data 14 bits\n"); }
        else if (!strcmp(k, "-I")) ipixf = 1;
        else if (!strcmp(k, "-f") && arg + 1 < argc) frames = atoi(argv[++arg]);
        else if (!strcmp(k, "-d") && arg + 1 < argc) devices = atoi(argv[++arg]);
        else if (!strcmp(k, "-o") && arg + 1 < argc) outdir = argv[++arg];
        else if (!strcmp(k, "-X")) ip.ybar = 1;                  // Ybar reconstruction inside Y'DzDx (yuv2tiff.cpp:162, 365-399)
        else if (!strcmp(k, "YUVPRIME2"))                        // sets a flag no reachable branch tests: DXYZ stays 1 (86, 159, 389)
            printf("Processing for Y'u''v'' (Y=rho-gamma, u'',v'' = linear)\n");
    }
    if (ip.matrix == H2Y_INV_709) printf("Processing for Rec709\n");
    if (ip.matrix == H2Y_INV_2020) printf("Processing for Rec2020\n");
    if (ip.width == 1920) printf("Processing for HD1920x1080\n");
    if (ip.width == 960) printf("Processing for HD960x540\n");
    if (!ip.full_range) printf("Processing for Video Range\n");
    const int nch = ip.alpha ? 4 : 3;
    printf("Stripsize (bytes): %d, %d (pixels), numStrips %d \n", ip.width * nch * 2, ip.width * nch * 2 / 8, ip.height);

    const size_t in_bytes = h2y_yuv_frame_bytes(ip.width, ip.height, H2Y_CHROMA_420), out_bytes = h2y_rgb_frame_bytes(&ip);
    const uint64_t have = h2yio::file_size(argv[1]) / in_bytes;
    if (have == 0) { printf("ERROR: %s holds no complete %dx%d 4:2:0 frame\n", argv[1], ip.width, ip.height); return 1; }
    const int nframes = frames > 0 ? (int)std::min<uint64_t>((uint64_t)frames, have) : (int)have;
    mkdir(outdir.c_str(), 0755);

    const int ndev = std::max(1, std::min(devices, nframes));
    std::vector<int> rc(ndev, 0);
    std::vector<uint32_t> invalid(nframes, 0);
    std::vector<std::thread> workers;
    for (int d = 0; d < ndev; d++) {
        workers.emplace_back([&, d]() {
            int lo = 0, hi = 0;
            h2y_frame_range(d, ndev, nframes, &lo, &hi);
            h2y_ctx *ctx = nullptr;
            h2y_status st = h2y_ctx_create(d, &ctx);
            if (st != H2Y_OK) { printf("ERROR: device %d: %s\n", d, h2y_status_string(st)); rc[d] = 1; return; }
            const int nb = std::min(hi - lo, (int)std::max<size_t>(1, (256u << 20) / out_bytes));
            uint8_t *hin = (uint8_t *)h2y_host_alloc(in_bytes * nb), *hout = (uint8_t *)h2y_host_alloc(out_bytes * nb);
            if (!hin || !hout) { printf("ERROR: pinned host allocation failed\n"); rc[d] = 1; return; }
            for (int f0 = lo; f0 < hi && rc[d] == 0; f0 += nb) {
                const int n = std::min(nb, hi - f0);
                std::string err;
                if (!h2yio::file_read_at(argv[1], hin, in_bytes * n, (uint64_t)f0 * in_bytes, &err)) { printf("ERROR: %s\n", err.c_str()); rc[d] = 1; break; }
                st = h2y_inverse_host(ctx, &ip, hin, in_bytes, hout, out_bytes, n, &invalid[f0]);
                if (st != H2Y_OK) { printf("%s (h2y_status %d)\n", h2y_status_string(st), (int)st); rc[d] = 2; break; }
                std::vector<std::thread> wr;            // TIFF writes of a batch in parallel
                std::vector<char> ok(n, 1);
                for (int i = 0; i < n; i++)
                    wr.emplace_back([&, i]() {
                        char name[64];
                        snprintf(name, sizeof(name), "/XpYpZp%05d.tif", f0 + i);
                        std::string e2;
                        ok[i] = h2yio::tiff_write_rgb16(outdir + name, reinterpret_cast<const uint16_t *>(hout + (size_t)i * out_bytes),
                                                        ip.width, ip.height, nch, &e2);
                    });
                for (auto &t : wr) t.join();
                for (int i = 0; i < n; i++) if (!ok[i]) { printf("ERROR: unable to write frame %d\n", f0 + i); rc[d] = 1; }
            }
            h2y_host_free(hin); h2y_host_free(hout);
            h2y_ctx_destroy(ctx);
        });
    }
    for (auto &t : workers) t.join();
    for (int d = 0; d < ndev; d++) if (rc[d]) return rc[d];
    for (int f = 0; f < nframes; f++) {
        printf("Writing %s/XpYpZp%05d.tif\n", outdir.c_str(), f);
        if (ipixf || invalid[f]) printf("Invalid Pixels:  %u\n", invalid[f]);
    }
    printf("%d frame(s) converted\n", nframes);
    return 0;
}
