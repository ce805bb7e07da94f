// Shared device/host helpers for the svs_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

typedef unsigned long long u64;

namespace svsb {

// ---------------------------------------------------------------------------------------------
// 64-bit selection keys.
//
// A row's key is (order-preserving transform of its fp32 score) << 32 | ~row.  Comparing keys as
// unsigned integers is exactly the engine's total order: larger score first and, for bit-equal
// scores, smaller row first (ascending row == ascending embeddings.id for a rowid scan,
// reference src/svs/kb.py:603-609).  Keys of distinct rows are distinct, so "the k largest keys"
// is a unique set: selection needs no tie handling of its own.
// NaN scores map above +inf, i.e. they are selected first -- np.argpartition also treats NaN as
// the largest value (reference src/svs/util.py:202).
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f32_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t b; memcpy(&b, &f, 4); return b;
#endif
}
__host__ __device__ __forceinline__ float bits_f32(uint32_t b) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}
__host__ __device__ __forceinline__ uint32_t f32_to_ordered(float f) {
    uint32_t b = f32_bits(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_f32(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return bits_f32(b);
}
__host__ __device__ __forceinline__ u64 make_key(float score, uint32_t row) {
    return ((u64)f32_to_ordered(score) << 32) | (u64)(uint32_t)(~row);
}
__host__ __device__ __forceinline__ uint32_t key_row(u64 key) { return ~(uint32_t)key; }
__host__ __device__ __forceinline__ float key_score(u64 key) { return ordered_to_f32((uint32_t)(key >> 32)); }

// Score stored for a tombstoned row: the bit pattern 0xffffffff orders below every other float under f32_to_ordered
// (key high word 0), -inf and negative-sign NaNs included.
__host__ __device__ __forceinline__ float dead_score() { return bits_f32(0xffffffffu); }

// Largest k the single-CTA selection kernel handles; larger k takes the full-sort path.
constexpr int K_FAST_MAX = 2048;
// Target upper bound on the number of row groups whose maxima the selection kernel scans.
constexpr int64_t GROUPS_TARGET = 16384;   // fits the selection kernel's shared memory (128 KB of keys)
constexpr int GROUP_SHIFT_MIN = 6;     // 64 rows per group at least

// rows per group = 1 << shift, chosen so that ceil(n / rows) <= GROUPS_TARGET
__host__ __device__ inline int group_shift_for(int64_t n) {
    int s = GROUP_SHIFT_MIN;
#ifndef __CUDA_ARCH__
    // measurement knob: SVSB_GROUP_SHIFT_MIN = 4..12 (rows per group = 1 << shift); fewer, larger groups make the
    // selection kernel's key staging / threshold phases cheaper and its candidate rescan dearer
    static const int env_min = [] { const char* v = getenv("SVSB_GROUP_SHIFT_MIN"); const int x = v ? atoi(v) : 0;
                                    return (x >= 4 && x <= 12) ? x : GROUP_SHIFT_MIN; }();
    s = env_min;
#endif
    while (((n + ((int64_t)1 << s) - 1) >> s) > GROUPS_TARGET) ++s;
    return s;
}

#ifdef __CUDACC__
// streaming 128-bit load: read-only path, do not allocate in L1 (every matrix byte is used once)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ u64 warp_max_u64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { u64 t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}
__device__ __forceinline__ u64 warp_min_u64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { u64 t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    return v;
}
// system-scope publication of data other GPUs read over NVLink peer memory (and its acquire side)
__device__ __forceinline__ void st_release_sys(u64* p, u64 v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ld_acquire_sys(const u64* p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fma4(float4& acc, const float4& a, const float4& b) {
    acc.x = fmaf(a.x, b.x, acc.x); acc.y = fmaf(a.y, b.y, acc.y);
    acc.z = fmaf(a.z, b.z, acc.z); acc.w = fmaf(a.w, b.w, acc.w);
}
#endif

}  // namespace svsb
