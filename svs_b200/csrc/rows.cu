// K0 -- load-path kernels: row L2 norms (validate / normalise) and the synthetic row generator.
//
// The reference stores embedding blobs verbatim (src/svs/kb.py:611-615) and only *checks*
// | ||v|| - 1 | <= 1e-3 when a vector enters the system (src/svs/embeddings/util.py:33-38, kb.py:58).
// The load path therefore always computes the row norms on the device (statistics + optional
// normalisation); by default rows stay verbatim so scores equal the reference's raw dot products.
#include "kernels.cuh"

namespace svsb {

// one warp per row; stats[0] = bits of max | ||row|| - 1 |, stats[1] = rows with deviation > tol
__global__ void __launch_bounds__(256)
row_norm_kernel(float* __restrict__ M, int64_t n, int d4, int normalize, float tol,
                float* __restrict__ norms, u64* __restrict__ stats)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float local_max = 0.f; unsigned local_bad = 0;
    for (int64_t r = warp_global; r < n; r += nwarps) {
        float4* row = reinterpret_cast<float4*>(M) + r * (int64_t)d4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = lane; c < d4; c += 32) { const float4 v = row[c]; fma4(acc, v, v); }
        const float ss = warp_sum((acc.x + acc.y) + (acc.z + acc.w));
        const float nrm = sqrtf(ss);
        const float dev = fabsf(nrm - 1.0f);
        if (lane == 0) {
            if (norms) norms[r] = nrm;
            if (dev > local_max || dev != dev) local_max = dev != dev ? __int_as_float(0x7f800000) : dev;
            if (!(dev <= tol)) ++local_bad;
        }
        if (normalize && nrm > 0.f) {
            for (int c = lane; c < d4; c += 32) {
                float4 v = row[c];
                v.x = __fdiv_rn(v.x, nrm); v.y = __fdiv_rn(v.y, nrm); v.z = __fdiv_rn(v.z, nrm); v.w = __fdiv_rn(v.w, nrm);
                row[c] = v;
            }
        }
    }
    if (lane == 0) {
        if (local_max > 0.f) atomicMax(reinterpret_cast<unsigned long long*>(&stats[0]), (u64)__float_as_uint(local_max));
        if (local_bad) atomicAdd(reinterpret_cast<unsigned long long*>(&stats[1]), (u64)local_bad);
    }
}

cudaError_t launch_row_norms(cudaStream_t st, int device, float* M, int64_t n, int d, int ld, int normalize,
                             float tol, float* norms_or_null, u64* stats)
{
    (void)d;
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 7) / 8;                    // 8 warps per block
    const int64_t cap = (int64_t)sm_count(device) * 8;
    if (blocks > cap) blocks = cap;
    row_norm_kernel<<<(unsigned)blocks, 256, 0, st>>>(M, n, ld / 4, normalize, tol, norms_or_null, stats);
    count_launch();
    return cudaGetLastError();
}

// ---- synthetic rows: element (row, col) = f(seed, row, col); must match oracle/svs_oracle.py -----
__device__ __forceinline__ u64 mix64(u64 z) {
    z ^= z >> 33; z *= 0xFF51AFD7ED558CCDull;
    z ^= z >> 33; z *= 0xC4CEB9FE1A85EC53ull;
    z ^= z >> 33; return z;
}

__global__ void __launch_bounds__(256)
synth_kernel(float* __restrict__ M, int64_t n, int d, int ld, u64 seed, int64_t global_row0,
             int64_t* __restrict__ ids, int64_t id0, int64_t id_step)
{
    const int64_t total = n * (int64_t)ld;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / ld; const int c = (int)(i - r * ld);
        float v = 0.f;
        if (c < d) {
            const u64 grow = (u64)(global_row0 + r);
            const u64 ctr = (grow * (u64)d + (u64)c) * 0x9E3779B97F4A7C15ull + seed * 0xC4CEB9FE1A85EC53ull + 1ull;
            v = (float)(mix64(ctr) >> 40) * 5.9604644775390625e-8f;      // 2^-24
        }
        M[i] = v;
        if (c == 0 && ids) ids[r] = id0 + (global_row0 + r) * id_step;
    }
}

cudaError_t launch_synth(cudaStream_t st, int device, float* M, int64_t n, int d, int ld, uint64_t seed,
                         int64_t global_row0, int64_t* ids, int64_t id0, int64_t id_step)
{
    if (n <= 0) return cudaSuccess;
    const int64_t total = n * (int64_t)ld;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count(device) * 16;
    if (blocks > cap) blocks = cap;
    synth_kernel<<<(unsigned)blocks, 256, 0, st>>>(M, n, d, ld, (u64)seed, global_row0, ids, id0, id_step);
    count_launch();
    return cudaGetLastError();
}

}  // namespace svsb
