// Extras of the GL-free VolumeRenderer that the reference surface has no call for.
#pragma once
#include <cstdint>

namespace volrend { struct VolumeRenderer; }

// Copies the renderer's current RGBA8 frame (row 0 = top) to host memory; false if nothing
// has been rendered yet.  (The reference blits to the GL default framebuffer instead,
// src/cuda_renderer.cpp:121-124.)
bool volrend_b200_read_pixels(volrend::VolumeRenderer& r, uint8_t* rgba_host);
