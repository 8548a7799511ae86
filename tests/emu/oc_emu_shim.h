// TEST INFRASTRUCTURE ONLY.  Minimal host stand-ins for the CUDA vocabulary used by
// gym_comm_b200/csrc/oc_device.cuh, so the *same* device source can be executed lane by lane
// on the CPU in this GPU-less container (tests/emu/oc_emu.cpp).  Never linked into the product
// library; the product has no CPU path.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __grid_constant__
using std::min;
using std::max;
struct uint4 { uint32_t x, y, z, w; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct float2 { float x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __saturatef(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
static inline float __fmaf_rn(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline void __stcs(float4* p, float4 v) { *p = v; }
static inline void __syncwarp() {}
struct EmuDim { int x; };
static thread_local EmuDim threadIdx{0}, blockDim{1};
