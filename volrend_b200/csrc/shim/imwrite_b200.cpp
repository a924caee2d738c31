// imwrite_b200.cpp -- volrend::internal::write_png_file (include/volrend/internal/imwrite.hpp:9-10)
// over vr_write_png, so that `volrend_headless -o <dir>` (main_headless.cpp:216-222) writes images
// even though libpng is not available: same pixel format and compression level as
// src/imwrite.cpp:27-29,45-47 (8-bit RGBA, level 0, filter NONE).  Link this INSTEAD of the
// reference's src/imwrite.cpp, whose body compiles to a warning without VOLREND_PNG.
#include <cstdint>
#include <cstdio>
#include <string>

#include "volrend/internal/imwrite.hpp"
#include "volrend_b200.h"

namespace volrend {
namespace internal {

bool write_png_file(const std::string& filename, uint8_t* ptr, int width, int height) {
    if (!ptr) {
        fprintf(stderr, "PNG write failed\n");
        return false;
    }
    if (vr_write_png(filename.c_str(), ptr, width, height) != VR_OK) {
        fprintf(stderr, "PNG destination could not be opened\n");
        return false;
    }
    return true;
}

}  // namespace internal
}  // namespace volrend
