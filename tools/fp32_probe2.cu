// fp32_probe2.cu -- development microbenchmark #2: which resource bounds the scan inner loop?
// (register-file read bandwidth / operand reuse vs LDS latency vs issue slots)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
constexpr int kF = 12;
__constant__ float4 c_rec[128 * 7];

// ---- register-only patterns ------------------------------------------------
// MODE 0: FFMA2, q pair reused over 4 song pairs (the scan pattern, everything in registers)
// MODE 1: FFMA , q reused over 8 songs
// MODE 2: FFMA2, no reuse possible (q differs every instruction)
template <int MODE, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) reg_kernel(const float *src, float *out, int iters)
{
    float2 f[4][kF];
    float2 q[kF];
    const float2 *s2 = reinterpret_cast<const float2 *>(src);
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int j = 0; j < kF; ++j) f[p][j] = s2[(threadIdx.x * 4 + p) * kF + j];
#pragma unroll
    for (int j = 0; j < kF; ++j) q[j] = s2[4096 + j];
    float2 acc[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[p] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < kF; ++j)
#pragma unroll
                for (int p = 0; p < 4; ++p) acc[p] = __ffma2_rn(f[p][j], q[j], acc[p]);
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < kF; ++j)
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    acc[p].x = fmaf(f[p][j].x, q[j].x, acc[p].x);
                    acc[p].y = fmaf(f[p][j].y, q[j].x, acc[p].y);
                }
        } else {
#pragma unroll
            for (int j = 0; j < kF; ++j)
#pragma unroll
                for (int p = 0; p < 4; ++p) acc[p] = __ffma2_rn(f[p][j], q[(j + p * 3) % kF], acc[p]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) s += acc[p].x + acc[p].y;
    if (s == 12345.678f) out[0] = s;
}

// ---- scan-shaped loops, queries from constant memory ------------------------------------
// VAR 0: scalar FFMA, UR operands (record = 4 float4: {-T,q0,q1,q2},{q3..q6},{q7..q10},{q11,..})
// VAR 1: packed FFMA2, record = 7 float4 (dup'd), LDC.64 into registers
template <int VAR, int S, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) const_kernel(const float *store, int nq, int reps, unsigned long long *hits)
{
    const int64_t row0 = (int64_t)blockIdx.x * S * THREADS + threadIdx.x;
    unsigned long long myhits = 0;
    if (VAR == 0) {
        float f[S][kF];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int j = 0; j < kF; ++j) f[s][j] = store[(row0 + s * THREADS) * 12 + j];
        for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
            for (int ql = 0; ql < nq; ++ql) {
                const float4 *r = c_rec + ql * 4;
                const float4 q0 = r[0], q1 = r[1], q2 = r[2], q3 = r[3];
                const float q[kF] = {q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x};
                float acc[S];
#pragma unroll
                for (int s = 0; s < S; ++s) acc[s] = q0.x;
#pragma unroll
                for (int j = 0; j < kF; ++j)
#pragma unroll
                    for (int s = 0; s < S; ++s) acc[s] = fmaf(f[s][j], q[j], acc[s]);
                uint32_t m = 0xffffffffu;
#pragma unroll
                for (int s = 0; s < S; ++s) m &= __float_as_uint(acc[s]);
                if ((int)m >= 0) {
#pragma unroll
                    for (int s = 0; s < S; ++s) myhits += (acc[s] >= 0.f);
                }
            }
        }
    } else {
        float2 fp[S / 2][kF];
#pragma unroll
        for (int p = 0; p < S / 2; ++p)
#pragma unroll
            for (int j = 0; j < kF; ++j)
                fp[p][j] = make_float2(store[(row0 + (2 * p) * THREADS) * 12 + j], store[(row0 + (2 * p + 1) * THREADS) * 12 + j]);
        for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
            for (int ql = 0; ql < nq; ++ql) {
                const float4 *r = c_rec + ql * 7;
                const float4 t = r[0];
                float2 acc[S / 2];
#pragma unroll
                for (int p = 0; p < S / 2; ++p) acc[p] = make_float2(t.x, t.y);
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const float4 v = r[1 + c];
                    const float2 qa = make_float2(v.x, v.y), qb = make_float2(v.z, v.w);
#pragma unroll
                    for (int p = 0; p < S / 2; ++p) acc[p] = __ffma2_rn(fp[p][2 * c], qa, acc[p]);
#pragma unroll
                    for (int p = 0; p < S / 2; ++p) acc[p] = __ffma2_rn(fp[p][2 * c + 1], qb, acc[p]);
                }
                uint32_t m = 0xffffffffu;
#pragma unroll
                for (int p = 0; p < S / 2; ++p) m &= __float_as_uint(acc[p].x) & __float_as_uint(acc[p].y);
                if ((int)m >= 0) {
#pragma unroll
                    for (int p = 0; p < S / 2; ++p) myhits += (acc[p].x >= 0.f) + (acc[p].y >= 0.f);
                }
            }
        }
    }
    if (myhits) atomicAdd(hits, myhits);
}

static float *d_store;
static unsigned long long *d_hits;
static int g_sms;

template <int MODE, int THREADS, int MINB>
void run_reg(const char *name)
{
    float *d_out; CK(cudaMalloc(&d_out, 4));
    const int iters = 20000;
    const int grid = g_sms * MINB;
    auto kern = reg_kernel<MODE, THREADS, MINB>;
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<<<grid, THREADS>>>(d_store, d_out, 10);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int t = 0; t < 3; ++t) {
        CK(cudaEventRecord(e0));
        kern<<<grid, THREADS>>>(d_store, d_out, iters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    double fl = (double)grid * THREADS * (double)iters * 12 * 8 * 2.0;
    double tf = fl / (best * 1e-3) / 1e12;
    printf("%-40s regs=%3d warps/SMSP=%.1f %8.3f ms %7.2f TFLOP/s (%5.1f%% of 74.4)\n", name, fa.numRegs,
           THREADS * MINB / 128.0, best, tf, 100.0 * tf / 74.4);
    fflush(stdout);
}

template <int VAR, int S, int THREADS, int MINB>
void run_const(const char *name, int nq, int reps)
{
    const int REC = (VAR == 1) ? 28 : 16;
    std::vector<float> recs((size_t)128 * 28, 0.f);
    for (int q = 0; q < nq; ++q) {
        float *r = &recs[(size_t)q * REC];
        if (VAR == 1) {
            r[0] = r[1] = -3.0f;
            for (int j = 0; j < 12; ++j) r[4 + 2 * j] = r[5 + 2 * j] = 0.25f + 0.001f * (float)((q * 7 + j) % 50);
        } else {
            r[0] = -3.0f;
            for (int j = 0; j < 12; ++j) r[1 + j] = 0.25f + 0.001f * (float)((q * 7 + j) % 50);
        }
    }
    CK(cudaMemcpyToSymbol(c_rec, recs.data(), 128 * 28 * 4));
    auto kern = const_kernel<VAR, S, THREADS, MINB>;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, 0));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    int grid = g_sms * occ;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaMemset(d_hits, 0, 8));
    kern<<<grid, THREADS>>>(d_store, nq, 2, d_hits);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int t = 0; t < 3; ++t) {
        CK(cudaEventRecord(e0));
        kern<<<grid, THREADS>>>(d_store, nq, reps, d_hits);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    unsigned long long h; CK(cudaMemcpy(&h, d_hits, 8, cudaMemcpyDeviceToHost));
    double pairs = (double)grid * S * THREADS * (double)nq * reps;
    double tf = pairs * 24.0 / (best * 1e-3) / 1e12;
    printf("%-40s regs=%3d occ=%d grid=%4d %8.3f ms %7.2f TFLOP/s (%5.1f%% of 74.4) hits=%llu\n", name, fa.numRegs, occ, grid,
           best, tf, 100.0 * tf / 74.4, h);
    fflush(stdout);
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    g_sms = p.multiProcessorCount;
    const size_t rows = (size_t)g_sms * 16 * 512 * 2 + 8192;
    std::vector<float> h(rows * 12);
    uint32_t x = 12345u;
    for (auto &v : h) { x = x * 1664525u + 1013904223u; v = (float)(x >> 8) / 16777216.0f * 0.28f; }
    CK(cudaMalloc(&d_store, h.size() * 4));
    CK(cudaMemcpy(d_store, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_hits, 8));
    run_reg<0, 128, 1>("reg FFMA2 reuse4  1 warp/SMSP");
    run_reg<0, 256, 1>("reg FFMA2 reuse4  2 warps/SMSP");
    run_reg<0, 256, 2>("reg FFMA2 reuse4  4 warps/SMSP");
    run_reg<1, 128, 1>("reg FFMA  reuse8  1 warp/SMSP");
    run_reg<1, 256, 1>("reg FFMA  reuse8  2 warps/SMSP");
    run_reg<1, 256, 2>("reg FFMA  reuse8  4 warps/SMSP");
    run_reg<2, 128, 1>("reg FFMA2 noreuse 1 warp/SMSP");
    run_reg<2, 256, 2>("reg FFMA2 noreuse 4 warps/SMSP");
    const int nq = 128, reps = 400;
    run_const<0, 8, 256, 2>("const scalar-UR S8 T256x2", nq, reps);
    run_const<0, 8, 128, 4>("const scalar-UR S8 T128x4", nq, reps);
    run_const<0, 8, 512, 1>("const scalar-UR S8 T512x1", nq, reps);
    run_const<0, 12, 256, 1>("const scalar-UR S12 T256x1", nq, reps);
    run_const<0, 12, 384, 1>("const scalar-UR S12 T384x1", nq, reps);
    run_const<0, 16, 256, 1>("const scalar-UR S16 T256x1", nq, reps);
    run_const<0, 4, 256, 4>("const scalar-UR S4 T256x4", nq, reps);
    run_const<1, 8, 256, 2>("const packed S8 T256x2", nq, reps);
    run_const<1, 8, 512, 1>("const packed S8 T512x1", nq, reps);
    run_const<1, 12, 256, 1>("const packed S12 T256x1", nq, reps);
    run_const<1, 16, 256, 1>("const packed S16 T256x1", nq, reps);
    return 0;
}
