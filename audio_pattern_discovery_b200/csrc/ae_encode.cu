// ae_encode.cu -- the embedding step in front of the DTW path on the device (SURVEY.md
// section 8 row f3): NDSequence::encoded (src/spectrogram.rs:103-121) = AutoEncoder::predict
// (src/neural.rs:55-71) frame by frame, written straight into the padded sequence arena, so
// the reference's host loop (one 1 x n_bins Mat per frame, scalar Mat::mul) and the host
// round trip of the embeddings disappear.
//
// Arithmetic restated operation for operation in f32 (file:line in /root/reference/):
//   Mat::mul     src/numerics.rs:305-319  acc = 0.0; acc += x[k] * w[k*cols+j], k ascending; the
//                                         product is rounded before the add (no FMA)
//   add_col      src/numerics.rs:246-257  acc += b[j]
//   sigmoid      src/numerics.rs:222-232  1.0 / (1.0 + f32::exp(-acc))
//   scale        src/numerics.rs:296-301  * 255.0
//   mean, std    src/numerics.rs:12-29    sequential sums over the n_latent values
//   sigma = f32::max(std, 1.0); z_score (src/numerics.rs:71-73) (v - mu) / sigma
// f32::exp is the platform libm's expf.  glibc's expf (the one a Linux build of the reference
// calls) evaluates 2^(k/32) * p(r) in double precision and rounds once to f32; expf_glibc()
// below is that algorithm with the same table and coefficients, so it returns the same f32
// wherever the double-precision intermediate is not within ~2^-29 (relative) of an f32 rounding
// boundary -- 0 differences against the C library on 3.2e8 arguments (tests/test_ae_encode.py
// checks a corpus).  One thread per frame; the work is tiny (5 M frames x 10 latents at C3 scale).
#include "ae_encode.cuh"

#include "../../include/apd.h"

namespace apd {

namespace {

__constant__ unsigned long long kExp2fTab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

// glibc 2.27+ sysdeps/ieee754/flt-32/e_expf.c (EXP2F_TABLE_BITS = 5), round-to-nearest mode.
__device__ __forceinline__ float expf_glibc(float x)
{
    const unsigned int ax = __float_as_uint(x) & 0x7fffffffu;
    if (ax >= 0x42b00000u) {  // |x| >= 88 or NaN
        if (__float_as_uint(x) == 0xff800000u) return 0.0f;
        if (ax >= 0x7f800000u) return __fadd_rn(x, x);
        if (x > 0x1.62e42ep6f) return __int_as_float(0x7f800000);   // > log(2^128): overflow
        if (x < -0x1.9fe368p6f) return 0.0f;                       // < log(2^-150): underflow
    }
    const double InvLn2N = 0x1.71547652b82fep+0 * 32.0, Shift = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0, C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0,
                 C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
    const double xd = (double)x;
    double z = __dmul_rn(InvLn2N, xd);
    double kd = __dadd_rn(z, Shift);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dadd_rn(kd, -Shift);
    const double r = __dadd_rn(z, -kd);
    unsigned long long t = kExp2fTab[ki & 31u];
    t += ki << (52 - 5);
    const double s = __longlong_as_double((long long)t);
    z = __fma_rn(C0, r, C1);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(C2, r, 1.0);
    y = __fma_rn(z, r2, y);
    y = __dmul_rn(y, s);
    return __double2float_rn(y);
}

__global__ void __launch_bounds__(128) ae_encode_kernel(const float* __restrict__ raw, const uint64_t* __restrict__ src_off,
                                                        const uint32_t* __restrict__ off, const uint32_t* __restrict__ len,
                                                        uint32_t n, uint32_t n_bins, uint32_t n_latent, uint32_t dpad,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        float* __restrict__ arena)
{
    __shared__ float ws[APD_AE_MAX_BINS * APD_MAX_DIM];
    __shared__ float bs[APD_MAX_DIM];
    for (uint32_t k = threadIdx.x; k < n_bins * n_latent; k += blockDim.x) ws[k] = w[k];
    for (uint32_t k = threadIdx.x; k < n_latent; k += blockDim.x) bs[k] = b[k];
    __syncthreads();
    float x[APD_AE_MAX_BINS];
    float p[APD_MAX_DIM];
    for (uint32_t s = blockIdx.x; s < n; s += gridDim.x) {
        const float* src = raw + src_off[s];
        float* dst = arena + (size_t)off[s] * dpad;
        const uint32_t T = len[s];
        for (uint32_t t = threadIdx.x; t < T; t += blockDim.x) {
            for (uint32_t k = 0; k < n_bins; k++) x[k] = src[(size_t)t * n_bins + k];
            for (uint32_t j = 0; j < n_latent; j++) {
                float acc = 0.0f;
                for (uint32_t k = 0; k < n_bins; k++) acc = __fadd_rn(acc, __fmul_rn(x[k], ws[k * n_latent + j]));
                acc = __fadd_rn(acc, bs[j]);
                const float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf_glibc(-acc)));
                p[j] = __fmul_rn(sg, 255.0f);
            }
            float mu = 0.0f;
            for (uint32_t j = 0; j < n_latent; j++) mu = __fadd_rn(mu, p[j]);
            mu = __fdiv_rn(mu, (float)n_latent);
            float sd = 0.0f;
            for (uint32_t j = 0; j < n_latent; j++) {
                const float d = __fadd_rn(p[j], -mu);
                sd = __fadd_rn(sd, __fmul_rn(d, d));
            }
            sd = __fsqrt_rn(__fdiv_rn(sd, (float)n_latent));
            const float sigma = fmaxf(sd, 1.0f);  // f32::max: NaN loses
            for (uint32_t j = 0; j < n_latent; j++) dst[(size_t)t * dpad + j] = __fdiv_rn(__fadd_rn(p[j], -mu), sigma);
        }
    }
}

}  // namespace

cudaError_t ae_encode_launch(const float* d_raw, const uint64_t* d_src_off, const uint32_t* d_off, const uint32_t* d_len,
                             uint32_t n, uint32_t n_bins, uint32_t n_latent, uint32_t dpad, const float* d_w,
                             const float* d_b, float* d_arena, int sm_count, cudaStream_t stream)
{
    if (!n) return cudaSuccess;
    const int grid = (int)(n < (uint32_t)sm_count * 8 ? n : (uint32_t)sm_count * 8);
    ae_encode_kernel<<<grid, 128, 0, stream>>>(d_raw, d_src_off, d_off, d_len, n, n_bins, n_latent, dpad, d_w, d_b, d_arena);
    return cudaGetLastError();
}

}  // namespace apd
