// Hot-pixel prefilter (src/hotpixel.hpp:19-64) — GPU implementation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace pgs
{
static int hotpixel_filter_u16(uint16_t *, uint32_t, uint32_t, uint32_t, double, int, std::string &err)
{
    err = "hot-pixel prefilter is not available on the GPU path yet";
    return 3; // PGS_ERR_UNSUPPORTED
}
} // namespace pgs
