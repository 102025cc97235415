// hdr2yuv -- command-line host of the B200 conversion path with the reference's option surface
// (hdr2yuv.cpp:56-580).  Reads .tiff / .exr / .rgb / .yuv(4:4:4) sources, converts on the GPU through the
// C-ABI (include/hdr2yuv_b200.h) and appends planar u16 frames to the .yuv destination exactly where the
// reference's append-mode writer would put them (tiff.cpp:440).
//
// Beyond the reference: --n_frames / --src_start_frame address real frame sequences (the reference parses
// n_frames and ignores it, hdr2yuv.cpp:157-160), and --devices N shards the frame range over N GPUs
// (contiguous ranges, one context and one pinned pipeline per GPU, no collective).  Frame i always lands at
// byte offset i * frame_bytes of the output, so the file is identical for every N.
#include <fcntl.h>
#include <strings.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <chrono>
#include <thread>
#include <vector>

#include "h2y_io.h"
#include "hdr2yuv_b200.h"

namespace {

enum FileType { FT_UNDEFINED = 0, FT_YUV, FT_TIFF, FT_EXR, FT_Y4M, FT_DPX, FT_RGB };
struct TypeInfo { int idx; const char *name; int supported; };
// hdr.h:48-68; the .exr/.dpx/.rgb destinations are outside the accelerated path; a .tiff destination is
// served for .yuv sources (matrix_inverse + write_tiff, hdr2yuv.cpp:818-819, 930-933)
const TypeInfo kInputTypes[] = {{FT_UNDEFINED, "UNDEFINED", 0}, {FT_YUV, "yuv", 1}, {FT_TIFF, "tiff", 1}, {FT_EXR, "exr", 1},
                                {FT_Y4M, "y4m", 0}, {FT_DPX, "dpx", 1}, {FT_RGB, "rgb", 1}};
const TypeInfo kOutputTypes[] = {{FT_UNDEFINED, "UNDEFINED", 0}, {FT_YUV, "yuv", 1}, {FT_TIFF, "tiff", 1}, {FT_EXR, "exr", 1},
                                 {FT_Y4M, "y4m", 0}, {FT_DPX, "dpx", 1}, {FT_RGB, "rgb", 1}};      // hdr.h:59-68
// hdr.h:140-166 (index = transfer_characteristics code)
const TypeInfo kTransfers[] = {{0, "RESERVED0", 0}, {1, "BT709", 1}, {2, "UNSPECIFIED", 0}, {3, "RESERVED3", 0},
                               {4, "BT470M", 0}, {5, "BT470BG", 0}, {6, "BT601", 1}, {7, "SMPTE240M", 0}, {8, "LINEAR", 1},
                               {9, "LOG1", 0}, {10, "LOG2", 0}, {11, "IEC61966_2_4", 0}, {12, "XVYCC", 0}, {13, "SRGB", 0},
                               {14, "BT2020_10bit", 1}, {15, "BT2020_12bit", 1}, {16, "PQ", 1}, {17, "SMPTE428", 0},
                               {18, "RHO_GAMMA", 1}};
const int kNumTransfers = sizeof(kTransfers) / sizeof(kTransfers[0]);

struct Pic {             // the pic_t fields parse_options touches
    int width = 0, height = 0, bit_depth = 0, half_float_flag = 0, chroma_format_idc = 0, video_full_range_flag = 0;
    int colour_primaries = 0, transfer_characteristics = 0, matrix_coeffs = 0, chroma_sample_loc_type = 0;
};

struct Args {            // user_args_t (hdr.h:266-300) + this host's extensions
    const char *src_filename = nullptr, *dst_filename = nullptr, *ref_filename = nullptr;
    int sigma_compare = 0, src_start_frame = 0, n_frames = 0, verbose_level = 0, alpha_channel = 0;
    int cutout_hd = 0, cutout_qhd = 0, chroma_resampler_type = 0;
    int input_file_type = 0, output_file_type = 0;
    int devices = 1, batch_frames = 0;
    const char *dump_input = nullptr;
    Pic in, out;
};

const char *ext_of(const char *filename)
{
    const char *e = filename + strlen(filename);
    while (*e != '.' && e > filename) e--;
    return e + (*e == '.');
}

int type_of_ext(const char *ext, const TypeInfo *t, int n)
{
    for (int i = 0; i < n; i++) if (!strcasecmp(ext, t[i].name)) return t[i].idx;
    if (!strcasecmp(ext, "tif")) return FT_TIFF;
    return FT_UNDEFINED;
}

int transfer_of(const char *val, int fallback)
{
    char *end = nullptr;
    const long v = strtol(val, &end, 10);
    if (end != val && *end == 0) return (int)v;
    for (int i = 0; i < kNumTransfers; i++) if (!strcasecmp(val, kTransfers[i].name)) return kTransfers[i].idx;
    printf("WARNING: transfer characteristics (%s) unrecongized\n", val);
    return fallback;
}

void print_help()
{
    printf("Transfer chracteristics options:\n");
    for (int i = 0; i < kNumTransfers; i++)
        printf("%d: %s %s\n", i, kTransfers[i].name, kTransfers[i].supported ? "(SUPPORTED)" : "(NOT SUPPORTED)");
    printf("input file types:\n");
    for (int i = 0; i < 7; i++) printf("%d: %s %s\n", i, kInputTypes[i].name, kInputTypes[i].supported ? "(SUPPORTED)" : "(NOT SUPPORTED)");
    printf("output file types:\n");
    for (int i = 0; i < 7; i++) printf("%d: %s %s\n", i, kOutputTypes[i].name, kOutputTypes[i].supported ? "(SUPPORTED)" : "(NOT SUPPORTED)");
    printf("extensions of this host: --devices N (GPUs), --batch_frames B, --n_frames N over numbered files or raw frames\n");
}

// parse_options (hdr2yuv.cpp:56-580): same flags, same inheritance of unspecified dst fields, same sanity checks
void parse_options(Args *a, int argc, char *argv[])
{
    Pic &in = a->in, &out = a->out;
    out.chroma_format_idc = out.chroma_sample_loc_type = out.video_full_range_flag = -1;
    out.transfer_characteristics = out.matrix_coeffs = out.colour_primaries = -1;
    int arg_errors = 0;
    for (int i = 1; i < argc; i++) {
        const char *k = argv[i];
        const bool has = i + 1 < argc;
        const char *v = has ? argv[i + 1] : "";
        auto is = [&](const char *name) { return !strcmp(k, name); };
        if (is("--help")) { print_help(); continue; }
        if (!has) { printf("WARNING: argument (%s) unrecongized\n", k); continue; }
        if (is("--src_filename")) a->src_filename = v;
        else if (is("--dst_filename")) a->dst_filename = v;
        else if (is("--ref_filename")) a->ref_filename = v;
        else if (is("--sigma_compare")) a->sigma_compare = atoi(v);
        else if (is("--src_pic_width")) in.width = atoi(v);
        else if (is("--src_pic_height")) in.height = atoi(v);
        else if (is("--dst_pic_width")) out.width = atoi(v);
        else if (is("--dst_pic_height")) out.height = atoi(v);
        else if (is("--src_bit_depth")) in.bit_depth = atoi(v);
        else if (is("--src_half_float_flag")) in.half_float_flag = atoi(v);
        else if (is("--dst_bit_depth")) out.bit_depth = atoi(v);
        else if (is("--dst_half_float_flag")) out.half_float_flag = atoi(v);
        else if (is("--dst_chroma_format_idc")) out.chroma_format_idc = atoi(v);
        else if (is("--src_chroma_format_idc")) in.chroma_format_idc = atoi(v);
        else if (is("--src_start_frame")) a->src_start_frame = atoi(v);
        else if (is("--n_frames")) a->n_frames = atoi(v);
        else if (is("--verbose_level")) a->verbose_level = atoi(v);
        else if (is("--src_colour_primaries")) in.colour_primaries = atoi(v);
        else if (is("--dst_colour_primaries")) out.colour_primaries = atoi(v);
        else if (is("--src_matrix_coeffs")) in.matrix_coeffs = atoi(v);
        else if (is("--dst_matrix_coeffs")) out.matrix_coeffs = atoi(v);
        else if (is("--src_transfer_characteristics")) in.transfer_characteristics = transfer_of(v, in.transfer_characteristics);
        else if (is("--dst_transfer_characteristics")) out.transfer_characteristics = transfer_of(v, out.transfer_characteristics);
        else if (is("--alpha_channel")) a->alpha_channel = atoi(v);
        else if (is("--dst_video_full_range_flag")) out.video_full_range_flag = atoi(v);
        else if (is("--src_video_full_range_flag")) in.video_full_range_flag = atoi(v);
        else if (is("--cutout_hd")) a->cutout_hd = atoi(v);
        else if (is("--cutout_qhd")) a->cutout_qhd = atoi(v);
        else if (is("--chroma_resampler_type")) a->chroma_resampler_type = atoi(v);
        else if (is("--devices")) a->devices = atoi(v);
        else if (is("--batch_frames")) a->batch_frames = atoi(v);
        else if (is("--dump_input")) a->dump_input = v;
        else { printf("WARNING: argument (%s) unrecongized\n", k); continue; }
        i++;
    }
    const int vb = a->verbose_level;
    // unspecified destination fields inherit the source's (hdr2yuv.cpp:265-318)
    if (out.bit_depth == 0) { out.bit_depth = in.bit_depth; if (vb > 1) printf("output picture bit_depth not specified.  using input picture bit_depth(%d)\n", in.bit_depth); }
    if (out.width == 0) { out.width = in.width; if (vb > 1) printf("output picture width not specified.  using input picture width(%d)\n", in.width); }
    if (out.height == 0) { out.height = in.height; if (vb > 1) printf("output picture width not specified.  using input picture height(%d)\n", in.height); }
    if (out.chroma_format_idc == -1) { out.chroma_format_idc = in.chroma_format_idc; if (vb > 1) printf("output picture chroma_format_idc not specified.  using input picture chroma_format_idc(%d)\n", in.chroma_format_idc); }
    if (out.chroma_sample_loc_type == -1) out.chroma_sample_loc_type = in.chroma_sample_loc_type;
    if (out.video_full_range_flag == -1) { out.video_full_range_flag = in.video_full_range_flag; if (vb > 1) printf("output picture video_full_range_flag not specified.  using input picture video_full_range_flag(%d)\n", in.video_full_range_flag); }
    if (out.colour_primaries == -1) { out.colour_primaries = in.colour_primaries; if (vb > 1) printf("output picture colour_primaries not specified.  using input picture colour_primaries(%d)\n", in.colour_primaries); }
    if (out.transfer_characteristics == -1) { out.transfer_characteristics = in.transfer_characteristics; if (vb > 1) printf("output picture transfer_characteristics not specified.  using input picture transfer_characteristics(%d)\n", in.transfer_characteristics); }
    if (out.matrix_coeffs == -1) { out.matrix_coeffs = in.matrix_coeffs; if (vb > 1) printf("output picture matrix_coeffs not specified.  using input picture matrix_coeffs(%d)\n", in.matrix_coeffs); }

    if (!a->src_filename || !a->dst_filename) {
        printf("WARNING: --src_filename and --dst_filename are required\n");
        printf("TOO MANY ARGUMENT ERRORS. ABORTING PROGRAM. --help to show options\n\n");
        exit(0);
    }
    const char *ext = ext_of(a->src_filename);
    a->input_file_type = type_of_ext(ext, kInputTypes, 7);
    if (kInputTypes[a->input_file_type].supported != 1) {
        printf("WARNING: input file (%s) type extension (%s) idx(%d) is either not recongized or not supported\n", a->src_filename, ext, a->input_file_type);
        arg_errors++;
    }
    const bool int_in = a->input_file_type == FT_YUV || a->input_file_type == FT_TIFF || a->input_file_type == FT_RGB;
    if (int_in) {
        if (in.bit_depth < 10 || in.bit_depth > 16)
            printf("WARNING: src bit_depth(%d) outside range [10,16] for integer input file type(%s)\n", in.bit_depth, kInputTypes[a->input_file_type].name);
    } else if (in.chroma_format_idc != H2Y_CHROMA_444) {
        printf("file-type is 4:4:4.  Settig chroma_format_idc(%d) to  %d.\n", in.chroma_format_idc, H2Y_CHROMA_444);
        in.chroma_format_idc = H2Y_CHROMA_444;
    }
    ext = ext_of(a->dst_filename);
    a->output_file_type = type_of_ext(ext, kOutputTypes, 7);
    if (kOutputTypes[a->output_file_type].supported != 1) {
        printf("WARNING: output file (%s) type extension (%s) idx(%d) is either not recongized or not supported\n", a->dst_filename, ext, a->output_file_type);
        arg_errors++;
    }
    if (a->output_file_type == FT_EXR) {                         // hdr2yuv.cpp:412-428
        if (out.bit_depth != 16 && out.bit_depth != 32) {
            printf("WARNING: dst bit_depth(%d) must be 16 or 32 bits for float input file type(%s)\n", out.bit_depth, kOutputTypes[a->output_file_type].name);
            out.bit_depth = out.half_float_flag ? 32 : 16;
        }
    } else if (a->output_file_type == FT_DPX) {                  // hdr2yuv.cpp:430-443
        if (out.bit_depth != 32) {
            printf("WARNING: dst bit_depth(%d) must be 16 or 32 bits for float input file type(%s)\n", out.bit_depth, kOutputTypes[a->output_file_type].name);
            out.bit_depth = 32;
        }
    } else if (out.bit_depth < 10 || out.bit_depth > 16)
        printf("WARNING: dst bit_depth(%d) outside range [10,16] for integer input file type(%s)\n", out.bit_depth, kOutputTypes[a->output_file_type].name);

    printf("src_filename: %s (type: %s) %s\n", a->src_filename, kInputTypes[a->input_file_type].name, kInputTypes[a->input_file_type].supported ? "(SUPPORTED)" : "(NOT SUPPORTED)");
    printf("src_pic_width: %d\nsrc_pic_height: %d\nsrc_chroma_format_idc: %d\nsrc_bit_depth: %d\n", in.width, in.height, in.chroma_format_idc, in.bit_depth);
    printf("src_full_range_video_flag: %d\nsrc_colour_primaries: %d\n", in.video_full_range_flag, in.colour_primaries);
    const int ti = std::min(std::max(in.transfer_characteristics, 0), kNumTransfers - 1), to = std::min(std::max(out.transfer_characteristics, 0), kNumTransfers - 1);
    printf("src_transfer_characteristics: %d (type: %s) %s\n", in.transfer_characteristics, kTransfers[ti].name, kTransfers[ti].supported ? "(SUPPORTED)" : "(NOT SUPPORTED)");
    printf("src_matrix_coeffs: %d\n", in.matrix_coeffs);
    printf("dst_filename: %s (type: %s) %s\n", a->dst_filename, kOutputTypes[a->output_file_type].name, kOutputTypes[a->output_file_type].supported ? "(SUPPORTED)" : "(NOT SUPPORTED)");
    printf("dst_pic_width: %d\ndst_pic_height: %d\ndst_chroma_format_idc: %d\ndst_bit_depth: %d\n", out.width, out.height, out.chroma_format_idc, out.bit_depth);
    printf("dst_video_full_range_flag: %d\ndst_colour_primaries: %d\n", out.video_full_range_flag, out.colour_primaries);
    printf("dst_transfer_characteristics: %d (type: %s) %s\n", out.transfer_characteristics, kTransfers[to].name, kTransfers[to].supported ? "(SUPPORTED)" : "(NOT SUPPORTED)");
    printf("dst_matrix_coeffs: %d\n", out.matrix_coeffs);
    printf("verbose_level: %d\nsrc_start_frame: %d\nn_frames: %d\n", a->verbose_level, a->src_start_frame, a->n_frames);

    if (in.width < 2 || in.width > 10000) { printf("WARNING: pic_width(%d) outside range [0,10000]\n", in.width); arg_errors++; }
    if (in.height < 2 || in.height > 10000) { printf("WARNING: pic_height(%d) outside range [0,10000]\n", in.height); arg_errors++; }
    if (in.bit_depth < 8 || in.bit_depth > 32) { printf("WARNING: src bit_depth(%d) outside range [8,32]\n", in.bit_depth); arg_errors++; }
    if (in.chroma_format_idc != H2Y_CHROMA_444) {
        printf("WARNING: chroma_format_idc(%d) not %d, Only 4:4:4 input supported at this moment..\n", in.chroma_format_idc, H2Y_CHROMA_444);
        arg_errors++;
    }
    if (out.width < 2 || out.width > 10000) { printf("WARNING: pic_width(%d) outside range [0,10000]\n", out.width); arg_errors++; }
    if (out.height < 2 || out.height > 10000) { printf("WARNING: pic_height(%d) outside range [0,10000]\n", out.height); arg_errors++; }
    if (out.bit_depth < 8 || out.bit_depth > 32) { printf("WARNING: dst bit_depth(%d) outside range [32]\n", out.bit_depth); arg_errors++; }
    if (arg_errors) {
        printf("TOO MANY ARGUMENT ERRORS. ABORTING PROGRAM. --help to show options\n\n");
        exit(0);
    }
}

// What the reader of frame 0 establishes for the whole sequence
struct Source {
    int type = 0;
    int width = 0, height = 0, channels = 3;
    h2y_layout layout = H2Y_LAYOUT_RGB16;
    int clip_on_load = 0;
    size_t frame_bytes = 0;
    int crop_w = 0, crop_h = 0;
};

// dpx_read's three packings (dpx.cpp:316-341) as source layouts; 16-bit samples read with the half-float flag are halfs
h2y_layout dpx_layout(const h2yio::ImageInfo &info, int half_flag)
{
    if (info.bits == 16) return half_flag ? H2Y_LAYOUT_HALF_RGB : (info.big_endian ? H2Y_LAYOUT_DPX16_BE : H2Y_LAYOUT_DPX16_LE);
    if (info.bits == 32) return info.big_endian ? H2Y_LAYOUT_DPXF32_BE : H2Y_LAYOUT_DPXF32_LE;
    return info.big_endian ? H2Y_LAYOUT_DPX10_BE : H2Y_LAYOUT_DPX10_LE;
}

bool read_frame(const Args &a, const Source &s, int frame, uint8_t *dst, std::string *err)
{
    uint16_t *d16 = reinterpret_cast<uint16_t *>(dst);
    h2yio::ImageInfo info;
    switch (s.type) {
    case FT_TIFF: {
        const std::string name = h2yio::sequence_name(a.src_filename, a.src_start_frame + frame);
        if (!h2yio::tiff_read(name, d16, s.crop_w, s.crop_h, &info, err)) return false;
        if (info.width != s.width || info.height != s.height || info.channels != s.channels) { *err = name + ": geometry differs from the first frame"; return false; }
        return true;
    }
    case FT_EXR: {
        const std::string name = h2yio::sequence_name(a.src_filename, a.src_start_frame + frame);
        if (!h2yio::exr_read_half(name, d16, s.channels, &info, err)) return false;
        if (info.width != s.width || info.height != s.height) { *err = name + ": geometry differs from the first frame"; return false; }
        return true;
    }
    case FT_DPX: {
        const std::string name = h2yio::sequence_name(a.src_filename, a.src_start_frame + frame);
        if (!h2yio::dpx_read_words(name, reinterpret_cast<uint32_t *>(dst), &info, err)) return false;
        if (info.width != s.width || info.height != s.height) { *err = name + ": geometry differs from the first frame"; return false; }
        if (dpx_layout(info, a.in.half_float_flag) != s.layout) { *err = name + ": packing or byte order differs from the first frame"; return false; }
        if (s.layout == H2Y_LAYOUT_HALF_RGB && info.big_endian)      // half samples go to the GPU in the machine's order
            for (size_t i = 0; i < (size_t)s.width * s.height * 3; i++) d16[i] = (uint16_t)((d16[i] >> 8) | (d16[i] << 8));
        return true;
    }
    case FT_RGB:
        return h2yio::rgb_planar_read(a.src_filename, d16, s.width, s.height, a.src_start_frame + frame, err);
    default:   // .yuv 4:4:4: the three planes are the picture (hdr2yuv.cpp:641-643)
        return h2yio::file_read_at(a.src_filename, dst, s.frame_bytes, (uint64_t)(a.src_start_frame + frame) * s.frame_bytes, err);
    }
}

bool pwrite_all(int fd, const void *buf, size_t n, uint64_t off)
{
    const uint8_t *p = static_cast<const uint8_t *>(buf);
    while (n) {
        const ssize_t w = pwrite(fd, p, n, (off_t)off);
        if (w <= 0) return false;
        p += w; n -= (size_t)w; off += (uint64_t)w;
    }
    return true;
}

}   // namespace

int main(int argc, char *argv[])
{
    Args a;
    parse_options(&a, argc, argv);
    Pic &in = a.in, &out = a.out;

    // ---- what read_file() would establish (hdr2yuv.cpp:690-756): open frame 0 -------------------------------
    Source s;
    s.type = a.input_file_type;
    std::string err;
    h2yio::ImageInfo info;
    if (s.type == FT_TIFF) {
        const std::string first = h2yio::sequence_name(a.src_filename, a.src_start_frame);
        if (!h2yio::tiff_probe(first, &info, &err)) { printf("ERROR: unable top open %s (%s)\n", first.c_str(), err.c_str()); return 1; }
        printf("opened %s\n", first.c_str());
        // centre cut-outs of read_tiff (tiff.cpp:191-220): 3840 wide by default, HD / qHD on request
        s.crop_w = a.cutout_hd ? 1920 : (a.cutout_qhd ? 960 : (info.width > 3840 ? 3840 : 0));
        s.crop_h = a.cutout_hd ? 1080 : (a.cutout_qhd ? 540 : 0);
        s.width = s.crop_w && s.crop_w < info.width ? s.crop_w : info.width;
        s.height = s.crop_h && s.crop_h < info.height ? s.crop_h : info.height;
        s.channels = info.channels;
        s.layout = info.channels == 4 ? H2Y_LAYOUT_RGBA16 : H2Y_LAYOUT_RGB16;
        if (in.bit_depth != 16) printf("WARNING, read_tiff(): overriding used-specified bit_depth(%d) to tiff header value(%d)\n", in.bit_depth, 16);
        if (in.matrix_coeffs != H2Y_MATRIX_GBR) printf("WARNING, read_tiff(): matrix_coefs(%d) != MATRIX_GBR assumed for tiff input\n", in.matrix_coeffs);
        in.bit_depth = 16; in.matrix_coeffs = H2Y_MATRIX_GBR; in.chroma_format_idc = H2Y_CHROMA_444;
        s.clip_on_load = in.video_full_range_flag == 0;          // tiff.cpp:296-304
        printf("Frame size: %d x %d\n", s.width, s.height);
    } else if (s.type == FT_EXR) {
        const std::string first = h2yio::sequence_name(a.src_filename, a.src_start_frame);
        if (!h2yio::exr_probe(first, &info, &err)) { printf("ERROR: unable to open %s (%s)\n", first.c_str(), err.c_str()); return 1; }
        s.width = info.width; s.height = info.height; s.channels = 3;
        s.layout = H2Y_LAYOUT_HALF_RGB;
        if (in.bit_depth != 32) printf("read_exr(): overriding bit_depth(%d) to 32-bits (internal processing)\n", in.bit_depth);
        if (in.matrix_coeffs != H2Y_MATRIX_GBR) printf("read_exr(): overriding matrix_coeffs(%d) to MATRIX_GBR(%d)\n", in.matrix_coeffs, H2Y_MATRIX_GBR);
        if (in.video_full_range_flag != 1) printf("reading .exr file (%s):  setting input picture video_full_range_flag to 1", a.src_filename);
        in.bit_depth = 32; in.matrix_coeffs = H2Y_MATRIX_GBR; in.video_full_range_flag = 1; in.chroma_format_idc = H2Y_CHROMA_444;
    } else if (s.type == FT_DPX) {
        // hdr2yuv.cpp:700-737: a float picture (code / 1023.0), GBR, 4:4:4, 32 bits; the full-range flag is only reported
        const std::string first = h2yio::sequence_name(a.src_filename, a.src_start_frame);
        if (!h2yio::dpx_probe(first, &info, &err)) { printf(" %s, aborting\n", err.c_str()); return 1; }
        printf(" reading file %s width = %d, height = %d, cineon = 0\n", first.c_str(), info.width, info.height);
        s.width = info.width; s.height = info.height; s.channels = 3;
        s.layout = dpx_layout(info, in.half_float_flag);
        if (in.matrix_coeffs != H2Y_MATRIX_GBR) printf("reading .dpx file (%s):  setting input picture matrix_coef=%d (MATRIX_GBR)", a.src_filename, H2Y_MATRIX_GBR);
        if (in.chroma_format_idc != H2Y_CHROMA_444) printf("reading .dpx file (%s):  setting input picture chroma_format_idc=%d (CHROMA_444)", a.src_filename, H2Y_CHROMA_444);
        if (in.video_full_range_flag != 1) printf("reading .dpx file (%s):  setting input picture video_full_range_flag to 1", a.src_filename);
        in.bit_depth = 32; in.matrix_coeffs = H2Y_MATRIX_GBR; in.chroma_format_idc = H2Y_CHROMA_444;
    } else if (s.type == FT_RGB) {
        s.width = in.width; s.height = in.height; s.channels = 3; s.layout = H2Y_LAYOUT_RGB16;
        if (in.matrix_coeffs != H2Y_MATRIX_GBR) {
            printf("WARNING: RGB src matrix_coefs(%d) being overriden to MATRIX_GBR (%d)\n", in.matrix_coeffs, H2Y_MATRIX_GBR);
            in.matrix_coeffs = H2Y_MATRIX_GBR;
        }
    } else {
        s.width = in.width; s.height = in.height; s.channels = 3; s.layout = H2Y_LAYOUT_PLANAR_U16;
    }
    if (s.width != in.width) printf("overriding picture width(%d) to file header value(%d)\n", in.width, s.width);
    if (s.height != in.height) printf("overriding picture height(%d) to file header value(%d)\n", in.height, s.height);
    if (out.width == in.width) out.width = s.width;               // dst inherited the unspecified src size
    if (out.height == in.height) out.height = s.height;
    in.width = s.width; in.height = s.height;
    if (out.width != in.width || out.height != in.height) {
        printf("ERROR: picture resize (dst %dx%d != src %dx%d) is outside the accelerated path\n", out.width, out.height, in.width, in.height);
        return 1;
    }

    h2y_forward_params fp;
    memset(&fp, 0, sizeof(fp));
    fp.src.width = in.width; fp.src.height = in.height; fp.src.chroma_format_idc = in.chroma_format_idc;
    fp.src.transfer_characteristics = in.transfer_characteristics; fp.src.colour_primaries = in.colour_primaries;
    fp.src.matrix_coeffs = in.matrix_coeffs; fp.src.bit_depth = in.bit_depth;
    fp.src.video_full_range_flag = in.video_full_range_flag;
    fp.src.pic_buffer_type = (s.type == FT_EXR || s.type == FT_DPX) ? H2Y_PIC_TYPE_F32 : H2Y_PIC_TYPE_U16;
    fp.src.layout = s.layout;
    fp.dst.width = out.width; fp.dst.height = out.height; fp.dst.chroma_format_idc = out.chroma_format_idc;
    fp.dst.transfer_characteristics = out.transfer_characteristics; fp.dst.colour_primaries = out.colour_primaries;
    fp.dst.matrix_coeffs = out.matrix_coeffs; fp.dst.bit_depth = out.bit_depth;
    fp.dst.video_full_range_flag = out.video_full_range_flag;
    fp.dst.pic_buffer_type = H2Y_PIC_TYPE_U16; fp.dst.layout = H2Y_LAYOUT_PLANAR_U16;
    fp.chroma_resampler_type = a.chroma_resampler_type;
    fp.clip_on_load = s.clip_on_load;
    s.frame_bytes = h2y_src_frame_bytes(&fp.src);
    const size_t out_bytes = h2y_yuv_frame_bytes(out.width, out.height, out.chroma_format_idc);
    if (!s.frame_bytes || !out_bytes) { printf("ERROR: unsupported picture geometry\n"); return 1; }

    // ---- frame range -----------------------------------------------------------------------------------------------
    int nframes = a.n_frames > 0 ? a.n_frames : 1;
    if (s.type == FT_RGB || s.type == FT_YUV) {
        const uint64_t have = h2yio::file_size(a.src_filename) / s.frame_bytes;
        if ((uint64_t)a.src_start_frame + (uint64_t)nframes > have) {
            printf("ERROR, read_planar_integer_file(): %s holds %llu frames, asked for frames %d..%d\n", a.src_filename,
                   (unsigned long long)have, a.src_start_frame, a.src_start_frame + nframes - 1);
            return 1;
        }
    }

    if (a.dump_input) {      // decoder check without a GPU: the interleaved samples exactly as they would go to the device
        std::vector<uint8_t> buf(s.frame_bytes);
        FILE *f = fopen(a.dump_input, "wb");
        if (!f) { printf("ERROR: unable to create %s\n", a.dump_input); return 1; }
        for (int i = 0; i < nframes; i++) {
            if (!read_frame(a, s, i, buf.data(), &err)) { printf("ERROR: %s\n", err.c_str()); return 1; }
            fwrite(buf.data(), 1, buf.size(), f);
        }
        fclose(f);
        printf("dumped %d frame(s) of %dx%d, layout %d, %zu bytes each\n", nframes, s.width, s.height, (int)s.layout, s.frame_bytes);
        return 0;
    }

    if (a.output_file_type == FT_RGB) {          // the reference has no writer for it either (hdr2yuv.cpp:959-960)
        printf("WARNING: don't know what file type to write to.\n");
        return 0;
    }
    if (a.output_file_type == FT_EXR || a.output_file_type == FT_DPX) {
        // F32 destinations (hdr2yuv.cpp:826-850, 935-957): an RGB picture, 4:4:4; matrix_convert's F32-output twin is the
        // whole conversion; the planes go to write_exr_file's RGBA halfs or dpx_write_float's floats
        if (out.chroma_format_idc != H2Y_CHROMA_444) {
            printf("WARNING: out is RGB pic: chroma_format_idc(%d) != CHROMA_444 (%d)", out.chroma_format_idc, H2Y_CHROMA_444);
            out.chroma_format_idc = H2Y_CHROMA_444;
        }
        // (the reference also forces out.matrix_coeffs to GBR here, but only after matrix_convert has run with the
        // user's value, hdr2yuv.cpp:797-823 vs 845-849: the conversion keeps the requested matrix)
        if (out.matrix_coeffs != H2Y_MATRIX_GBR)
            printf("WARNING: out is RGB pic: matrix_coeffs(%d) != MATRIX_GBR (%d)", out.matrix_coeffs, H2Y_MATRIX_GBR);
        fp.dst.chroma_format_idc = H2Y_CHROMA_444;
        fp.dst.pic_buffer_type = H2Y_PIC_TYPE_F32; fp.dst.layout = H2Y_LAYOUT_PLANAR_F32;
        h2y_ctx *ctx = nullptr;
        h2y_status st = h2y_ctx_create(0, &ctx);
        if (st != H2Y_OK) { printf("ERROR: %s\n", h2y_status_string(st)); return 1; }
        const size_t npix = (size_t)s.width * s.height, ob = npix * 12;
        uint8_t *hin = (uint8_t *)h2y_host_alloc(s.frame_bytes);
        float *hout = (float *)h2y_host_alloc(ob);
        if (!hin || !hout) { printf("ERROR: pinned host allocation failed\n"); return 1; }
        for (int i = 0; i < nframes; i++) {
            if (!read_frame(a, s, i, hin, &err)) { printf("ERROR: %s\n", err.c_str()); return 1; }
            st = h2y_forward_f32_host(ctx, &fp, hin, s.frame_bytes, hout, ob, 1);
            if (st == H2Y_ERR_UNSUPPORTED) {
                printf("ERROR: a float destination needs an integer source here: with a float source the reference sizes its tmp picture at 32 bits "
                       "and set_pic_clip shifts by 32 (hdr2yuv.cpp:805-808, common.cpp:303-311)\n");
                return 1;
            }
            if (st != H2Y_OK) { printf("%s (h2y_status %d)\n", h2y_status_string(st), (int)st); return st == H2Y_ERR_PRECONDITION ? 1 : 2; }
            const std::string name = nframes > 1 ? h2yio::sequence_name(a.dst_filename, i) : std::string(a.dst_filename);
            const bool ok = a.output_file_type == FT_EXR
                                ? h2yio::exr_write_rgba_from_float(name, hout, hout + npix, hout + 2 * npix, s.width, s.height, 3, &err)
                                : h2yio::dpx_write_float(name, hout, hout + npix, hout + 2 * npix, s.width, s.height, &err);
            if (!ok) { printf("%s\n", err.c_str()); return 1; }
        }
        h2y_host_free(hin); h2y_host_free(hout);
        h2y_ctx_destroy(ctx);
        printf("wrote %d %s frame(s)\n", nframes, a.output_file_type == FT_EXR ? "exr" : "dpx");
        return 0;
    }
    if (a.output_file_type == FT_TIFF) {
        // .yuv (4:4:4) -> .tiff: matrix_inverse into a tmp picture of the source depth, write_tiff at the dst depth
        if (s.type != FT_YUV) { printf("ERROR: a .tiff destination needs a .yuv source (matrix_inverse, hdr2yuv.cpp:818-819)\n"); return 1; }
        h2y_ctx *ctx = nullptr;
        h2y_status st = h2y_ctx_create(0, &ctx);
        if (st != H2Y_OK) { printf("ERROR: %s\n", h2y_status_string(st)); return 1; }
        const size_t fb = (size_t)s.width * s.height * 6;
        uint8_t *hin = (uint8_t *)h2y_host_alloc(fb), *hout = (uint8_t *)h2y_host_alloc(fb);
        if (!hin || !hout) { printf("ERROR: pinned host allocation failed\n"); return 1; }
        for (int i = 0; i < nframes; i++) {
            if (!read_frame(a, s, i, hin, &err)) { printf("ERROR: %s\n", err.c_str()); return 1; }
            uint32_t invalid = 0;
            st = h2y_inverse444_host(ctx, &fp.src, out.bit_depth, hin, fb, hout, fb, 1, &invalid);
            if (st == H2Y_ERR_MATRIX) { printf("Can't determine color difference to use?\n\n"); return 0; }   // exit(0), convert.cpp:1737
            if (st != H2Y_OK) { printf("%s (h2y_status %d)\n", h2y_status_string(st), (int)st); return 1; }
            const std::string name = nframes > 1 ? h2yio::sequence_name(a.dst_filename, i) : std::string(a.dst_filename);
            if (!h2yio::tiff_write_rgb16(name, reinterpret_cast<const uint16_t *>(hout), s.width, s.height, 3, &err)) {
                printf("unable to open %s.  Exiting\n", name.c_str());
                return 0;
            }
            if (a.verbose_level > 0) printf("%s: invalid pixels %u\n", name.c_str(), invalid);
        }
        h2y_host_free(hin); h2y_host_free(hout);
        h2y_ctx_destroy(ctx);
        printf("wrote %d tiff frame(s)\n", nframes);
        return 0;
    }

    // ---- output: frame i at (existing size) + i * frame_bytes, as the reference's append would leave it ---------
    const int fd = open(a.dst_filename, O_WRONLY | O_CREAT, 0644);
    if (fd < 0) { printf("ERROR: unable to open %s for writing\n", a.dst_filename); return 1; }
    const uint64_t base = (uint64_t)lseek(fd, 0, SEEK_END);

    const int ndev = std::max(1, std::min(a.devices, nframes));
    // frames per h2y_forward_host call = decode threads per device (file decoding is what bounds this host)
    int batch = a.batch_frames > 0 ? a.batch_frames : (int)std::max<size_t>(1, std::min<size_t>(16, (448u << 20) / s.frame_bytes));
    std::vector<int> rc(ndev, 0);
    std::vector<std::thread> workers;
    // seconds each device's worker spent decoding (thread body), inside h2y_forward_host, writing (thread body), setting up
    std::vector<double> t_decode(ndev, 0.0), t_gpu(ndev, 0.0), t_write(ndev, 0.0), t_setup(ndev, 0.0);
    auto now = []() { return std::chrono::steady_clock::now(); };
    auto since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); };
    const auto t_start = std::chrono::steady_clock::now();
    for (int d = 0; d < ndev; d++) {
        workers.emplace_back([&, d]() {
            int lo = 0, hi = 0;
            h2y_frame_range(d, ndev, nframes, &lo, &hi);
            const auto t_s = now();
            h2y_ctx *ctx = nullptr;
            h2y_status st = h2y_ctx_create(d, &ctx);
            if (st != H2Y_OK) { printf("ERROR: device %d: %s\n", d, h2y_status_string(st)); rc[d] = 1; return; }
            const int nb = std::min(batch, hi - lo);
            // three stages overlap: batch k+1 is decoded and batch k-1 is written while batch k is on the GPU
            uint8_t *hin[2], *hout[2];
            for (int i = 0; i < 2; i++) {
                hin[i] = (uint8_t *)h2y_host_alloc(s.frame_bytes * nb);
                hout[i] = (uint8_t *)h2y_host_alloc(out_bytes * nb);
            }
            if (!hin[0] || !hin[1] || !hout[0] || !hout[1]) { printf("ERROR: pinned host allocation failed\n"); rc[d] = 1; return; }
            std::string rerr[2];
            bool rok[2] = {true, true};
            t_setup[d] = since(t_s);
            auto load = [&](int slot, int f0, int n) {      // decode with a few threads: file decoding is the host's bottleneck
                const auto t_l = now();
                rok[slot] = true;
                std::vector<std::thread> th;
                std::vector<std::string> e(n);
                std::vector<char> ok(n, 1);
                for (int i = 0; i < n; i++)
                    th.emplace_back([&, i]() { ok[i] = read_frame(a, s, f0 + i, hin[slot] + (size_t)i * s.frame_bytes, &e[i]); });
                for (auto &t : th) t.join();
                for (int i = 0; i < n; i++) if (!ok[i]) { rok[slot] = false; rerr[slot] = e[i]; }
                t_decode[d] += since(t_l);
            };
            int slot = 0;
            bool wok = true;
            std::thread writer;
            load(slot, lo, std::min(nb, hi - lo));
            for (int f0 = lo; f0 < hi && rc[d] == 0; f0 += nb) {
                const int n = std::min(nb, hi - f0);
                if (!rok[slot]) { printf("ERROR: %s\n", rerr[slot].c_str()); rc[d] = 1; break; }
                std::thread next;
                const int nf0 = f0 + nb;
                if (nf0 < hi) next = std::thread(load, slot ^ 1, nf0, std::min(nb, hi - nf0));     // overlap decode with the GPU
                const auto t_g = now();
                st = h2y_forward_host(ctx, &fp, hin[slot], s.frame_bytes, hout[slot], out_bytes, n);
                t_gpu[d] += since(t_g);
                if (writer.joinable()) writer.join();                                               // batch k-1 is on disk
                if (st != H2Y_OK) {
                    printf("%s (h2y_status %d)\n", h2y_status_string(st), (int)st);
                    rc[d] = st == H2Y_ERR_PRECONDITION ? 1 : 2;
                } else if (!wok) {
                    printf("ERROR: write to %s failed\n", a.dst_filename);
                    rc[d] = 1;
                } else {
                    const uint8_t *wsrc = hout[slot];
                    // one pwrite per frame, side by side: a single writer is bounded by the page cache's per-thread copy rate
                    writer = std::thread([&, wsrc, n, f0]() {
                        const auto t_w = now();
                        std::vector<std::thread> th;
                        std::vector<char> ok(n, 1);
                        for (int i = 0; i < n; i++)
                            th.emplace_back([&, i]() {
                                ok[i] = pwrite_all(fd, wsrc + (size_t)i * out_bytes, out_bytes, base + (uint64_t)(f0 + i) * out_bytes);
                            });
                        for (auto &t : th) t.join();
                        for (int i = 0; i < n; i++) if (!ok[i]) wok = false;
                        t_write[d] += since(t_w);
                    });
                    if (a.verbose_level > 0)
                        printf("device %d: frames %d..%d converted, %llu kernel launches so far\n", d, f0, f0 + n - 1,
                               (unsigned long long)h2y_kernel_launches(ctx));
                }
                if (next.joinable()) next.join();
                slot ^= 1;
            }
            if (writer.joinable()) writer.join();
            if (!wok && rc[d] == 0) { printf("ERROR: write to %s failed\n", a.dst_filename); rc[d] = 1; }
            for (int i = 0; i < 2; i++) { h2y_host_free(hin[i]); h2y_host_free(hout[i]); }
            h2y_ctx_destroy(ctx);
        });
    }
    for (auto &t : workers) t.join();
    close(fd);
    for (int d = 0; d < ndev; d++) if (rc[d]) return rc[d];
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    printf("wrote %d frame(s) of %zu bytes to %s\n", nframes, out_bytes, a.dst_filename);
    // files in, file out, context creation included: the end-to-end figure of this host
    printf("%d device(s): %.3f s, %.1f frames/s, %.1f Mpixel/s, source %.2f GB/s, destination %.2f GB/s\n", ndev, secs, nframes / secs,
           (double)nframes * s.width * s.height / secs * 1e-6, (double)nframes * s.frame_bytes / secs * 1e-9,
           (double)nframes * out_bytes / secs * 1e-9);
    for (int d = 0; d < ndev; d++)
        printf("device %d: setup %.3f s, decode %.3f s, h2y_forward_host %.3f s, write %.3f s (stages overlap)\n", d, t_setup[d], t_decode[d],
               t_gpu[d], t_write[d]);
    return 0;
}
