/*
 * ref_io_stubs.cpp -- the two EXR entry points hdr2yuv.cpp links against (hdr.h:416-417; the DPX ones come from the
 * reference's own dpx.cpp, compiled with a stand-in for OpenEXR's `half`) when the
 * reference's whole program is built as oracle/_ref/hdr2yuv_ref.  TEST INFRASTRUCTURE ONLY.  OpenEXR is not in this
 * image, so the EXR routes of the reference's main() cannot run here: they stop loudly.  Everything else
 * (option parsing and inheritance, the file-type table, the validation exits, .rgb / .yuv / .tiff sources, .yuv / .tiff
 * destinations) is the reference's own code.
 */
#include <cstdio>
#include <cstdlib>

#include "hdr.h"

static void gone(const char *what)
{
    fprintf(stderr, "hdr2yuv_ref: %s is not available in the oracle build (no OpenEXR)\n", what);
    exit(3);
}
void read_exr(pic_t *, char *) { gone("read_exr"); }
int write_exr_file(char *, int, int, int, pic_t *) { gone("write_exr_file"); return 1; }
