// sr_device.cuh -- device-side building blocks of the cosine top-K engine
// (sm_100a only).  See DESIGN.md for the data layout and the proof sketch of the
// filter-with-margin + exact re-score scheme.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sr {

constexpr int kF = 12;            // reference Song.h:12 FEATURE_COUNT
constexpr int kQTMax = 256;       // queries resident in shared memory per CTA
constexpr int kSortCap = 4096;    // keys sorted per pass by one CTA in finalize / sample
constexpr int kKMax = 1024;       // largest supported top-K
constexpr int kRowPad = 36864;    // store rows are padded to a multiple of every tile size

// The fused filter value differs from the oracle's score by at most ~54 u
// (u = 2^-24, 3.2e-6) for regular rows and queries: DESIGN.md "filter slack".
// Every filter threshold is the exact running K-th best minus kEps.
constexpr float kEps = 5.0e-6f;
// bound pass: a filter score computed from a zero start is within 26u + 14u < 2.4e-6 of the
// oracle's score of the same pair; lowered by this before it is used as a K-th-best bound
constexpr float kBoundSlack = 4.0e-6f;
// rows / queries whose exact norm is outside [kNormLo, kNormHi] (and not 0) are
// "irregular": the bound above assumes no under/overflow, so they carry NaN in
// the normalised store / record, always pass the filter and are scored exactly.
constexpr float kNormLo = 1.0e-3f;
constexpr float kNormHi = 1.0e15f;


// ---- order-preserving key encoding ----------------------------------------
// key = orderable(score) << 32 | (0xFFFFFFFF - id): larger key == better under
// (score descending, id ascending).  key 0 is the "no entry" sentinel.
__host__ __device__ __forceinline__ uint32_t f2ord(float f)
{
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o)
{
    uint32_t u = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
constexpr uint32_t kOrdNegInf = 0x007FFFFFu;  // f2ord(-inf)

__device__ __forceinline__ uint64_t make_key(float score, uint32_t id)
{
    // +0.0f folds -0 into +0 so that signed zeros tie (oracle compares with ==)
    return ((uint64_t)f2ord(__fadd_rn(score, 0.0f)) << 32) | (uint64_t)(0xFFFFFFFFu - id);
}
__device__ __forceinline__ uint32_t key_id(uint64_t k) { return 0xFFFFFFFFu - (uint32_t)k; }
__device__ __forceinline__ float key_score(uint64_t k) { return ord2f((uint32_t)(k >> 32)); }

// ---- the oracle's arithmetic, instruction for instruction ------------------
// reference Recommender.cu:263-271: unfused multiply/add in feature order,
// den = sqrtf(norm) * queryNorm, IEEE divide, std::min/std::max clamp.
// __fmul_rn/__fadd_rn are never contracted into FMA by nvcc.
__device__ __forceinline__ float exact_norm(const float *f)
{
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < kF; ++j) acc = __fadd_rn(acc, __fmul_rn(f[j], f[j]));
    return __fsqrt_rn(acc);
}

// IEEE-754 round-to-nearest FP32 division a / b for b > 0, WITHOUT the subroutine call
// nvcc emits for `/` and __fdiv_rn (its slow path is a CALL, and one CALL anywhere in a
// kernel makes ptxas stop keeping warp-uniform operands in uniform registers -- which is
// what the scan's FFMA2 stream lives on).  The quotient is formed in double precision
// from MUFU.RCP64H + two Newton steps + one residual correction (relative error < 2^-51)
// and rounded once to FP32.  That equals the correctly rounded FP32 quotient: the exact
// quotient of two 24-bit significands is never closer than 2^-49 (relative) to a rounding
// boundary of the 24-bit (or a subnormal) format, so the double result always rounds the
// same way.  tests/test_engine_gpu.py::test_division_matches_ieee pins it against the
// host's divss bit for bit.
__device__ __forceinline__ float ieee_div_pos(float a, float b)
{
    if (!(fabsf(a) <= 3.402823466e+38f) || !(b <= 3.402823466e+38f)) {
        // a is NaN/inf or b is +inf/NaN: finite/inf = signed 0, everything else NaN or inf
        if (a != a || b != b) return __int_as_float(0x7fc00000);
        if (fabsf(a) > 3.402823466e+38f) return (b > 3.402823466e+38f) ? __int_as_float(0x7fc00000) : a;
        return copysignf(0.0f, a);
    }
    if (a == 0.0f) return a;  // keeps the sign of a zero numerator
    const double bd = (double)b, ad = (double)a;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(bd));
    r = fma(r, fma(-bd, r, 1.0), r);
    r = fma(r, fma(-bd, r, 1.0), r);
    double q = ad * r;
    q = fma(fma(-bd, q, ad), r, q);
    return __double2float_rn(q);
}

__device__ __forceinline__ float exact_finish(float dot, float nf, float qn)
{
    float den = __fmul_rn(nf, qn);
    float s = 0.0f;
    if (den > 1e-8f) {
        s = ieee_div_pos(dot, den);
        s = (s < 1.0f) ? s : 1.0f;     // std::min(1.0f, s)
        s = (-1.0f < s) ? s : -1.0f;   // std::max(-1.0f, s)
    }
    return s;
}

__device__ __forceinline__ float exact_score(const float *f, float nf, const float *q, float qn)
{
    float dot = 0.0f;
#pragma unroll
    for (int j = 0; j < kF; ++j) dot = __fadd_rn(dot, __fmul_rn(q[j], f[j]));
    return exact_finish(dot, nf, qn);
}

__device__ __forceinline__ void load_row12(const float *base, int64_t row, float *out)
{
    const float4 *p = reinterpret_cast<const float4 *>(base) + row * 3;
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
    out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
    out[8] = c.x; out[9] = c.y; out[10] = c.z; out[11] = c.w;
}

// ---- block-wide bitonic sort, descending, of L (power of two) 64-bit keys ----
template <int THREADS>
__device__ __forceinline__ void bitonic_desc(uint64_t *s, int L)
{
    for (int k2 = 2; k2 <= L; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < L; i += THREADS) {
                int p = i ^ j;
                if (p > i) {
                    uint64_t a = s[i], b = s[p];
                    bool desc = ((i & k2) == 0);
                    if ((a < b) == desc) { s[i] = b; s[p] = a; }
                }
            }
            __syncthreads();
        }
    }
}

template <int THREADS>
__device__ __forceinline__ void bitonic_desc_u32(uint32_t *s, int L)
{
    for (int k2 = 2; k2 <= L; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < L; i += THREADS) {
                int p = i ^ j;
                if (p > i) {
                    uint32_t a = s[i], b = s[p];
                    bool desc = ((i & k2) == 0);
                    if ((a < b) == desc) { s[i] = b; s[p] = a; }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int next_pow2(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace sr
