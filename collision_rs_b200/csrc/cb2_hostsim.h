// cb2_hostsim.h — lets the per-lane device code (cb2_math/cells/gjk/pair/epa_lane .cuh) compile
// with plain g++ so that tests/host_sim can run the very same source one lane at a time on the
// CPU and compare it with the oracle.  TEST INFRASTRUCTURE: only defined under CB2_HOST_SIM, which
// the product build (nvcc, Makefile) never sets; there is no CPU fallback in the library.
//
// g++ must be given -ffp-contract=off -fno-fast-math: the __f*_rn shims below are then single
// IEEE-754 binary32 operations, like the intrinsics they stand for.
#pragma once
#ifdef CB2_HOST_SIM
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline uint32_t __float_as_uint(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
static inline float __uint_as_float(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (s & 31u));
}
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
// one lane at a time: a warp-wide vote / shuffle sees only the calling lane
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline bool __all_sync(unsigned, bool p) { return p; }
static inline bool __any_sync(unsigned, bool p) { return p; }
template <class T>
static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) {
    const uint32_t o = *p;
    *p = o + v;
    return o;
}
using std::max;
using std::min;
#define CB2_NOINLINE_DEVICE static inline
#else
#define CB2_NOINLINE_DEVICE static __device__ __noinline__
#endif
