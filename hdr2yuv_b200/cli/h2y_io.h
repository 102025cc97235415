// h2y_io.h -- minimal native readers/writers for the command-line hosts.
//
// libtiff and OpenEXR are not available in this image; these cover what the reference's readers accept on
// the hot path and nothing more:
//   TIFF  baseline, uncompressed, 16 bits per sample, 3 or 4 samples per pixel, chunky, strips of any height,
//         either byte order (tiff.cpp:54-362 reads raw strips the same way); centre cut-outs as tiff.cpp:191-220
//   EXR   single-part scanline files, HALF (or FLOAT, narrowed to half like Imf::RgbaInputFile) channels R,G,B[,A],
//         compression NONE / ZIPS / ZIP via zlib (exr.cpp:138-261 goes through RgbaInputFile)
//   DPX   10-bit packed RGB, one 32-bit word per pixel, either byte order (dpx.cpp:210-360, 506-531); the words go
//         to the GPU as they are in the file (H2Y_LAYOUT_DPX10_BE / _LE), the unpack and /1023.0 happen there
//   raw   planar .rgb (R,G,B planes) and .yuv frames of u16 (hdr2yuv.cpp:582-656)
// Every reader delivers the decoder's natural interleaved layout; the de-interleave to G,B,R planes, the
// on-read clip and the half->float widening happen on the GPU (h2y_layout).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace h2yio {

struct ImageInfo {
    int width = 0, height = 0;
    int channels = 0;          // 3 or 4 (interleaved R,G,B[,A])
    int bits = 0;              // 16
    bool is_half = false;      // EXR: half bit patterns; TIFF: integer codes
    bool big_endian = false;   // DPX: byte order of the pixel words
    uint32_t data_offset = 0;  // DPX: offset of the image data
};

// --- TIFF -----------------------------------------------------------------------------------------------
// Reads the header only.
bool tiff_probe(const std::string &path, ImageInfo *info, std::string *err);
// Decodes into `dst` (width*height*channels u16, native endian).  crop_w/crop_h > 0 select a centred window
// (tiff.cpp:191-220); `dst` must hold the cropped size.
bool tiff_read(const std::string &path, uint16_t *dst, int crop_w, int crop_h, ImageInfo *info, std::string *err);
// One strip per row, PHOTOMETRIC_RGB, 16 bits, chunky (yuv2tiff.cpp:326-342, tiff.cpp:585-600).
bool tiff_write_rgb16(const std::string &path, const uint16_t *src, int width, int height, int channels, std::string *err);

// --- EXR ------------------------------------------------------------------------------------------------
bool exr_probe(const std::string &path, ImageInfo *info, std::string *err);
// dst: width*height*out_channels half bit patterns, interleaved r,g,b[,a]; a missing A is written as 1.0
bool exr_read_half(const std::string &path, uint16_t *dst, int out_channels, ImageInfo *info, std::string *err);
// Test helper and --dst .exr is out of scope: writes an uncompressed or ZIP scanline file of HALF channels.
bool exr_write_half(const std::string &path, const uint16_t *src, int width, int height, int channels, int compression,
                    std::string *err);

// --- DPX ------------------------------------------------------------------------------------------------
// Header fields dpx_read uses: magic (byte order), image offset at byte 4, width / height at 772 / 776, bit size at
// 803 (dpx.cpp:268-348).  The three packings dpx_read knows are served: 10-bit packed words (4 bytes per pixel), 16-bit
// samples (6) and 32-bit floats (12); 12-bit and everything else is refused with dpx_read's message.  info->bits says which.
bool dpx_probe(const std::string &path, ImageInfo *info, std::string *err);
// dst: the pixel payload exactly as stored, width*height*(4|6|12) bytes (no byte swap: the layout tells the GPU the order)
bool dpx_read_words(const std::string &path, uint32_t *dst, ImageInfo *info, std::string *err);
// dpx_write_float (dpx.cpp:719-920): a little-endian 32-bit float DPX with that function's 2048-byte header; planes in
// pic_t.fbuf order G, B, R (planar_float_to_muxed_dpx_buf, common.cpp:32-47, puts R first)
bool dpx_write_float(const std::string &path, const float *g, const float *b, const float *r, int width, int height, std::string *err);
// Test helper: 16-bit (bits = 16, u16 samples) or float (bits = 32) DPX with a generic header in either byte order
bool dpx_write_raw(const std::string &path, const void *rgb, int bits, int width, int height, bool big_endian, std::string *err);
// float planes G, B, R -> an RGBA half EXR with A = 0, what write_exr_file stores (exr.cpp:99-133; the container is ZIP or
// uncompressed scanlines here, PIZ there)
bool exr_write_rgba_from_float(const std::string &path, const float *g, const float *b, const float *r, int width, int height,
                               int compression, std::string *err);
// Test helper: a 2048-byte generic header + packed words, like dpx_write_10bit_from_float's (dpx.cpp:554-716).
bool dpx_write_10bit(const std::string &path, const uint16_t *rgb10, int width, int height, bool big_endian, std::string *err);

// --- raw planar -----------------------------------------------------------------------------------------
// .rgb: planes R,G,B of width*height u16 each; returns interleaved R,G,B in dst (frame index `frame`)
bool rgb_planar_read(const std::string &path, uint16_t *dst, int width, int height, long frame, std::string *err);
bool file_read_at(const std::string &path, void *dst, size_t bytes, uint64_t offset, std::string *err);
uint64_t file_size(const std::string &path);

// --- frame sequences ------------------------------------------------------------------------------------
// Name of frame `index` of a one-file-per-frame sequence: a printf pattern (%d / %05d) is formatted with
// start+index; otherwise the LAST run of digits in the base name is incremented by index (width kept).
std::string sequence_name(const std::string &first, int index);

uint16_t float_to_half(float f);

}   // namespace h2yio
