// layout_probe.cu -- development microbenchmark: the store-layout decision of DESIGN.md section 3, as an A/B.
//   A  what the engine uses: 48 B per song, rows PRE-NORMALISED (the inverse norm folded into the row), two songs
//      interleaved so a 128-bit load yields two FFMA2 operands; 12 FFMA2 per song pair and query.
//   B  what north_star (2) sketches: rows padded to 16 floats = 64 B per song with a stored inverse norm; the dot
//      product runs over the raw features and is scaled by the song's inverse norm afterwards: 13 FFMA2 per song
//      pair and query, 8 more registers per thread, a third more bytes per pass over the store.
// Two regimes: "stream" = one query over a 10 M-song store (HBM-bound: bytes decide), "loop" = 128 queries over a
// register-resident tile, repeated (FP32-bound: instructions decide).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
constexpr int kF = 12;
__constant__ float4 c_q[128 * 3];

// one filter pass of S songs (S/2 packed pairs) against query ql; PADDED adds the per-song inverse-norm scaling
template <int S, bool PADDED>
__device__ __forceinline__ uint32_t filter(const float2 (&fp)[S / 2][kF], const float2 (&inv)[S / 2], int ql, float t)
{
    const float4 *r = c_q + ql * 3;
    const float4 q0 = r[0], q1 = r[1], q2 = r[2];
    const float q[kF] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
    float2 acc[S / 2];
#pragma unroll
    for (int p = 0; p < S / 2; ++p) acc[p] = PADDED ? make_float2(0.f, 0.f) : make_float2(t, t);
#pragma unroll
    for (int j = 0; j < kF; ++j)
#pragma unroll
        for (int p = 0; p < S / 2; ++p) acc[p] = __ffma2_rn(fp[p][j], make_float2(q[j], q[j]), acc[p]);
    if (PADDED) {
#pragma unroll
        for (int p = 0; p < S / 2; ++p) acc[p] = __ffma2_rn(acc[p], inv[p], make_float2(t, t));
    }
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int p = 0; p < S / 2; ++p) m &= __float_as_uint(acc[p].x) & __float_as_uint(acc[p].y);
    return m;
}

// FP32-bound regime: a register-resident tile, 128 queries, `reps` rounds
template <int S, int THREADS, bool PADDED>
__global__ void __launch_bounds__(THREADS, 1) loop_kernel(const float *store, int nq, int reps, unsigned long long *hits)
{
    const int64_t row0 = (int64_t)blockIdx.x * S * THREADS + threadIdx.x;
    float2 fp[S / 2][kF], inv[S / 2];
#pragma unroll
    for (int p = 0; p < S / 2; ++p) {
#pragma unroll
        for (int j = 0; j < kF; ++j)
            fp[p][j] = make_float2(store[(row0 + (2 * p) * THREADS) * 16 + j], store[(row0 + (2 * p + 1) * THREADS) * 16 + j]);
        inv[p] = make_float2(store[(row0 + (2 * p) * THREADS) * 16 + 12], store[(row0 + (2 * p + 1) * THREADS) * 16 + 12]);
    }
    unsigned long long myhits = 0;
    for (int rep = 0; rep < reps; ++rep) {
        uint32_t signs = 0;
#pragma unroll 16
        for (int ql = 0; ql < nq; ++ql) signs = __funnelshift_l(filter<S, PADDED>(fp, inv, ql, -3.0f), signs, 1);
        myhits += __popc(~signs);
    }
    if (myhits) atomicAdd(hits, myhits);
}

// HBM-bound regime: ONE query over the whole store, persistent grid, 128-bit loads
//   A: pair-interleaved 48-B rows (six loads per pair)   B: 64-B padded rows (four loads per song)
template <int S, int THREADS, bool PADDED>
__global__ void __launch_bounds__(THREADS, 2) stream_kernel(const float4 *store, int64_t n_tiles, unsigned long long *hits)
{
    unsigned long long myhits = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        float2 fp[S / 2][kF], inv[S / 2];
        if (!PADDED) {
            const float4 *src = store + (tile * (S / 2) * THREADS + threadIdx.x) * 6;
#pragma unroll
            for (int p = 0; p < S / 2; ++p) {
#pragma unroll
                for (int c4 = 0; c4 < 6; ++c4) {
                    const float4 v = __ldg(src + (int64_t)p * THREADS * 6 + c4);
                    fp[p][2 * c4] = make_float2(v.x, v.y);
                    fp[p][2 * c4 + 1] = make_float2(v.z, v.w);
                }
                inv[p] = make_float2(1.f, 1.f);
            }
        } else {
            const float4 *src = store + (tile * S * THREADS + threadIdx.x) * 4;
#pragma unroll
            for (int p = 0; p < S / 2; ++p) {
                float4 a[4], b[4];
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    a[c4] = __ldg(src + (int64_t)(2 * p) * THREADS * 4 + c4);
                    b[c4] = __ldg(src + (int64_t)(2 * p + 1) * THREADS * 4 + c4);
                }
#pragma unroll
                for (int c4 = 0; c4 < 3; ++c4) {
                    fp[p][4 * c4] = make_float2(a[c4].x, b[c4].x);
                    fp[p][4 * c4 + 1] = make_float2(a[c4].y, b[c4].y);
                    fp[p][4 * c4 + 2] = make_float2(a[c4].z, b[c4].z);
                    fp[p][4 * c4 + 3] = make_float2(a[c4].w, b[c4].w);
                }
                inv[p] = make_float2(a[3].x, b[3].x);
            }
        }
        const uint32_t m = filter<S, PADDED>(fp, inv, 0, -3.0f);
        myhits += (m >> 31) ^ 1u;
    }
    if (myhits) atomicAdd(hits, myhits);
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int64_t n = 10000000 / 2048 * 2048;  // whole tiles of 8 x 256 songs
    float *d_store; unsigned long long *d_hits;
    CK(cudaMalloc(&d_store, (size_t)n * 16 * 4));
    std::vector<float> h((size_t)n * 16);
    uint32_t x = 12345u;
    for (auto &v : h) { x = x * 1664525u + 1013904223u; v = (float)(x >> 8) / 16777216.0f * 0.28f; }
    CK(cudaMemcpy(d_store, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_hits, 8)); CK(cudaMemset(d_hits, 0, 8));
    std::vector<float> q(128 * 12);
    for (size_t i = 0; i < q.size(); ++i) q[i] = 0.25f + 0.001f * (float)(i % 53);
    CK(cudaMemcpyToSymbol(c_q, q.data(), q.size() * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time_it = [&](auto launch) {
        launch(); CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int t = 0; t < 5; ++t) {
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        return best;
    };
    // ---- HBM-bound: one query over 10 M songs
    const float tA = time_it([&] { stream_kernel<8, 256, false><<<sms * 2, 256>>>((const float4 *)d_store, n / 2048, d_hits); });
    const float tB = time_it([&] { stream_kernel<8, 256, true><<<sms * 2, 256>>>((const float4 *)d_store, n / 2048, d_hits); });
    printf("stream, 1 query x %lld songs:  A 48-B pre-normalised interleaved rows %7.1f us (%6.1f GB/s)   B 64-B padded rows + inverse norm %7.1f us (%6.1f GB/s)   B/A = %.3f\n",
           (long long)n, tA * 1e3, 48.0 * n / (tA * 1e-3) / 1e9, tB * 1e3, 64.0 * n / (tB * 1e-3) / 1e9, tB / tA);
    // ---- FP32-bound: 128 queries over register-resident tiles
    const int reps = 400;
    const float lA = time_it([&] { loop_kernel<8, 512, false><<<sms, 512>>>(d_store, 128, reps, d_hits); });
    const float lB = time_it([&] { loop_kernel<8, 512, true><<<sms, 512>>>(d_store, 128, reps, d_hits); });
    const double pairs = (double)sms * 8 * 512 * 128.0 * reps;
    cudaFuncAttributes fa, fb;
    CK(cudaFuncGetAttributes(&fa, loop_kernel<8, 512, false>)); CK(cudaFuncGetAttributes(&fb, loop_kernel<8, 512, true>));
    printf("loop, 128 queries per tile:    A 12 FFMA2 per pair %7.3f ms = %5.2f TFLOP/s algorithmic (%d regs)   B 13 FFMA2 per pair %7.3f ms = %5.2f TFLOP/s algorithmic (%d regs)   B/A = %.3f\n",
           lA, pairs * 24 / (lA * 1e-3) / 1e12, fa.numRegs, lB, pairs * 24 / (lB * 1e-3) / 1e12, fb.numRegs, lB / lA);
    return 0;
}
