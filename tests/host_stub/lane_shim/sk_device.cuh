// TEST INFRASTRUCTURE ONLY -- host stand-in for sickle_b200/csrc/sk_device.cuh.
//
// tests/test_lane_logic.py (and test_copy_logic.py, for sk_copy.cuh) copies sickle_b200/csrc/trim_lane.cuh (the device function that trims one
// read with one lane, shared by the fused kernel and K2) next to this file and k1_index.cuh,
// and compiles it with g++: the CUDA qualifiers become empty macros, the integer intrinsics are
// restated here.  The integer logic that decides the cut points is then checked against the oracle on
// CPU.  Nothing here is used by the product.
#pragma once

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>

#define __device__
#define __forceinline__ inline
#define __noinline__
#define __restrict__

struct alignas(16) uint4 {
    uint32_t x, y, z, w;
};
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
inline void __stcs(uint4 *p, uint4 v) { *p = v; }

using std::max;
using std::min;

namespace sk {

// same members as the real DevParams (only names matter here)
struct DevParams {
    int32_t qoff, qmin, qmax;
    int32_t qthr;
    int32_t lthr;
    int32_t no_fiveprime;
    int32_t trunc_n;
    int32_t mode;
    int32_t emu_threads;
    int32_t has_singles;
};

}  // namespace sk

inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    sh &= 31u;
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> sh);
}
inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
    sh &= 31u;
    return (uint32_t)(((((uint64_t)hi << 32) | lo) << sh) >> 32);
}
inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c) {
    for (int k = 0; k < 4; ++k) c += ((a >> (8 * k)) & 255u) * ((b >> (8 * k)) & 255u);
    return c;
}

