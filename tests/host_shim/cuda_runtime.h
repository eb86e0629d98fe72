// Host stand-ins for the handful of CUDA intrinsics elba_b200/csrc/common.cuh uses, so that the
// 2-bit parse / rolling canonical k-mer device code can be compiled with g++ and checked on a CPU
// (tests/test_host_parse.py).  Test infrastructure only.
#pragma once
#include <stdint.h>
#include <algorithm>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
using std::min; using std::max;
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { uint64_t v = ((uint64_t)hi << 32) | lo; return (uint32_t)(v >> (sh & 31)); }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
    uint64_t v = ((uint64_t)b << 32) | a; uint32_t r = 0;
    for (int i = 0; i < 4; ++i) { uint32_t s = (sel >> (4 * i)) & 7; r |= (uint32_t)((v >> (8 * s)) & 0xff) << (8 * i); }
    return r;
}
static inline uint64_t __umul64hi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
