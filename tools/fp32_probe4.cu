// fp32_probe4.cu -- development microbenchmark #4: does sharing the song operand between
// consecutive FFMA2 of DIFFERENT queries (register-reuse cache) lift the uniform-operand loop
// past the register-file read limit?  NQ queries per iteration, song pair read once per NQ FFMA2.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
constexpr int kF = 12;
__constant__ float4 c_q[1365 * 3];

__device__ __forceinline__ float2 ffma2_pin(float2 a, float b, float2 c)
{
    unsigned long long ra, rb, rc, rd;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(rb) : "f"(b));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}

template <int S, int THREADS, int MINB, int NQ, int PIN, int UNROLL>
__global__ void __launch_bounds__(THREADS, MINB) ur_kernel(const float *store, const float *thr, int nq, int q_base, int reps,
                                                          unsigned long long *hits)
{
    __shared__ float s_thr[128];
    for (int i = threadIdx.x; i < nq; i += THREADS) s_thr[i] = thr[i];
    __syncthreads();
    const int64_t row0 = (int64_t)blockIdx.x * S * THREADS + threadIdx.x;
    unsigned long long myhits = 0;
    float2 fp[S / 2][kF];
#pragma unroll
    for (int p = 0; p < S / 2; ++p)
#pragma unroll
        for (int j = 0; j < kF; ++j)
            fp[p][j] = make_float2(store[(row0 + (2 * p) * THREADS) * 12 + j], store[(row0 + (2 * p + 1) * THREADS) * 12 + j]);
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll UNROLL
        for (int ql = 0; ql < nq; ql += NQ) {
            float q[NQ][kF];
            float2 acc[NQ][S / 2];
#pragma unroll
            for (int g = 0; g < NQ; ++g) {
                const float4 *r = c_q + (q_base + ql + g) * 3;
                const float4 q0 = r[0], q1 = r[1], q2 = r[2];
                q[g][0] = q0.x; q[g][1] = q0.y; q[g][2] = q0.z; q[g][3] = q0.w;
                q[g][4] = q1.x; q[g][5] = q1.y; q[g][6] = q1.z; q[g][7] = q1.w;
                q[g][8] = q2.x; q[g][9] = q2.y; q[g][10] = q2.z; q[g][11] = q2.w;
                const float t = s_thr[ql + g];
#pragma unroll
                for (int p = 0; p < S / 2; ++p) acc[g][p] = make_float2(t, t);
            }
#pragma unroll
            for (int j = 0; j < kF; ++j)
#pragma unroll
                for (int p = 0; p < S / 2; ++p)
#pragma unroll
                    for (int g = 0; g < NQ; ++g)
                        acc[g][p] = PIN ? ffma2_pin(fp[p][j], q[g][j], acc[g][p])
                                        : __ffma2_rn(fp[p][j], make_float2(q[g][j], q[g][j]), acc[g][p]);
#pragma unroll
            for (int g = 0; g < NQ; ++g) {
                uint32_t m = 0xffffffffu;
#pragma unroll
                for (int p = 0; p < S / 2; ++p) m &= __float_as_uint(acc[g][p].x) & __float_as_uint(acc[g][p].y);
                if ((int)m >= 0) {
#pragma unroll
                    for (int p = 0; p < S / 2; ++p) myhits += (acc[g][p].x >= 0.f) + (acc[g][p].y >= 0.f);
                }
            }
        }
    }
    if (myhits) atomicAdd(hits, myhits);
}

static float *d_store, *d_thr;
static unsigned long long *d_hits;
static int g_sms;

template <int S, int THREADS, int MINB, int NQ, int PIN, int UNROLL>
void run_ur(const char *name, int nq, int reps)
{
    if (getenv("PROBE_ONLY")) {
        char buf[512]; strncpy(buf, getenv("PROBE_ONLY"), 511); buf[511] = 0;
        bool ok = false;
        for (char *t = strtok(buf, ","); t; t = strtok(nullptr, ",")) ok |= strstr(name, t) != nullptr;
        if (!ok) return;
    }
    auto kern = ur_kernel<S, THREADS, MINB, NQ, PIN, UNROLL>;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, 0));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    int grid = g_sms * occ;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaMemset(d_hits, 0, 8));
    kern<<<grid, THREADS>>>(d_store, d_thr, nq, 0, 2, d_hits);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int t = 0; t < 3; ++t) {
        CK(cudaEventRecord(e0));
        kern<<<grid, THREADS>>>(d_store, d_thr, nq, 0, reps, d_hits);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    unsigned long long h; CK(cudaMemcpy(&h, d_hits, 8, cudaMemcpyDeviceToHost));
    double pairs = (double)grid * S * THREADS * (double)nq * reps;
    double tf = pairs * 24.0 / (best * 1e-3) / 1e12;
    printf("%-36s nq=%3d regs=%3d occ=%d grid=%4d %8.3f ms %7.2f TFLOP/s (%5.1f%% of 74.4) hits=%llu\n", name, nq, fa.numRegs,
           occ, grid, best, tf, 100.0 * tf / 74.4, h);
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
    std::vector<float> q(1365 * 12);
    for (size_t i = 0; i < q.size(); ++i) q[i] = 0.25f + 0.001f * (float)(i % 53);
    CK(cudaMemcpyToSymbol(c_q, q.data(), q.size() * 4));
    std::vector<float> thr(128, -3.0f);
    CK(cudaMalloc(&d_thr, 512));
    CK(cudaMemcpy(d_thr, thr.data(), 512, cudaMemcpyHostToDevice));
    const int reps = getenv("PROBE_REPS") ? atoi(getenv("PROBE_REPS")) : 400;
    run_ur<8, 512, 1, 1, 0, 2>("A S8 T512x1 nq1 u2", 128, reps);
    run_ur<8, 512, 1, 1, 0, 16>("A S8 T512x1 nq1 u16", 128, reps);
    run_ur<8, 512, 1, 2, 0, 1>("B S8 T512x1 nq2 u1", 128, reps);
    run_ur<8, 512, 1, 2, 0, 4>("B S8 T512x1 nq2 u4", 128, reps);
    run_ur<8, 512, 1, 4, 0, 1>("C S8 T512x1 nq4 u1", 128, reps);
    run_ur<8, 512, 1, 2, 1, 1>("D S8 T512x1 nq2 pin u1", 128, reps);
    run_ur<8, 512, 1, 2, 1, 4>("D S8 T512x1 nq2 pin u4", 128, reps);
    run_ur<8, 512, 1, 4, 1, 1>("E S8 T512x1 nq4 pin u1", 128, reps);
    run_ur<8, 512, 1, 4, 1, 2>("E S8 T512x1 nq4 pin u2", 128, reps);
    run_ur<8, 256, 2, 2, 1, 2>("F S8 T256x2 nq2 pin u2", 128, reps);
    run_ur<8, 256, 2, 4, 1, 1>("F S8 T256x2 nq4 pin u1", 128, reps);
    return 0;
}
