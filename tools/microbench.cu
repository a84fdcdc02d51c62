// microbench.cu -- issue-rate probes for the FP32 instructions the DTW kernels use, on
// the actual B200 (SURVEY.md section 6 asks for the FP32 roofline to be confirmed on the
// box).  Each probe runs ILP independent dependency chains per thread, 8 warps x 4 CTAs
// per SM, and reports warp-instructions per clock per SM (4.0 = one per SMSP per clock).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define ILP 8
#define ITERS 4096

enum Op { FFMA, FFMA2, FADD, FADD2, FMUL2, FMNMX, SETP_SEL, RSQ, SQRT_APPROX, FFMA_FMNMX_MIX, FADD_FMNMX_MIX, LDS128_BCAST, N_OPS };
static const char* kNames[] = {"ffma", "ffma2 (f32x2)", "fadd", "fadd2 (f32x2)", "fmul2 (f32x2)", "fmnmx", "fsetp+sel",
                               "mufu.rsq", "mufu.sqrt", "ffma+fmnmx 1:1", "fadd+fmnmx 1:1", "lds.128 broadcast"};
static const int kInstrPerIter[] = {1, 1, 1, 1, 1, 1, 2, 1, 1, 2, 2, 1};

template <int OP>
__global__ void __launch_bounds__(256) probe(float* out, long long* cycles, float seed)
{
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    float a[ILP], b[ILP];
    unsigned long long p[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) { a[k] = seed + k; b[k] = seed * 0.5f + k; p[k] = ((unsigned long long)__float_as_uint(a[k]) << 32) | __float_as_uint(b[k]); }
    const float c0 = seed * 1.0001f, c1 = seed * 0.9999f;
    const unsigned long long c2 = ((unsigned long long)__float_as_uint(c0) << 32) | __float_as_uint(c1);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(c0), "f"(c1));
            if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[k]) : "l"(c2));
            if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(c0));
            if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(c2));
            if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(c2));
            if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(b[k]));
            if (OP == SETP_SEL) asm volatile("{.reg .pred q; setp.lt.f32 q, %0, %1; selp.f32 %0, %1, %2, q;}" : "+f"(a[k]) : "f"(b[k]), "f"(c0));
            if (OP == RSQ) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
            if (OP == SQRT_APPROX) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
            if (OP == FFMA_FMNMX_MIX) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(c0), "f"(c1)); asm volatile("min.f32 %0, %0, %1;" : "+f"(b[k]) : "f"(c0)); }
            if (OP == FADD_FMNMX_MIX) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(c0)); asm volatile("min.f32 %0, %0, %1;" : "+f"(b[k]) : "f"(c0)); }
            if (OP == LDS128_BCAST) { float4 v = sm[(it + k) & 63]; a[k] += v.x; b[k] += v.w; }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += a[k] + b[k] + __uint_as_float((unsigned)(p[k] >> 32)) + __uint_as_float((unsigned)p[k]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, float* d_out, long long* d_cyc)
{
    const int ctas = sms * 4;
    probe<OP><<<ctas, 256>>>(d_out, d_cyc, 1.0f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<OP><<<ctas, 256>>>(d_out, d_cyc, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long* h = (long long*)malloc(sizeof(long long) * ctas);
    cudaMemcpy(h, d_cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < ctas; i++) avg += (double)h[i];
    avg /= ctas;
    free(h);
    // 4 CTAs x 8 warps per SM run concurrently for ~avg cycles
    double winstr = 32.0 * ITERS * ILP * kInstrPerIter[OP];
    double per_clk_sm = winstr / avg;
    double total = (double)ctas * 8 * ITERS * ILP * kInstrPerIter[OP];
    printf("%-20s %8.3f warp-instr/clk/SM   %8.1f Gwarp-instr/s (event)   eff clock %.0f MHz\n", kNames[OP], per_clk_sm,
           total / (ms * 1e-3) / 1e9, (total / (ms * 1e-3)) / (per_clk_sm * sms) / 1e6);
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, sms);
    float* d_out; long long* d_cyc;
    cudaMalloc(&d_out, sizeof(float) * sms * 4 * 256);
    cudaMalloc(&d_cyc, sizeof(long long) * sms * 4);
    run<FFMA>(sms, d_out, d_cyc);
    run<FFMA2>(sms, d_out, d_cyc);
    run<FADD>(sms, d_out, d_cyc);
    run<FADD2>(sms, d_out, d_cyc);
    run<FMUL2>(sms, d_out, d_cyc);
    run<FMNMX>(sms, d_out, d_cyc);
    run<SETP_SEL>(sms, d_out, d_cyc);
    run<RSQ>(sms, d_out, d_cyc);
    run<SQRT_APPROX>(sms, d_out, d_cyc);
    run<FFMA_FMNMX_MIX>(sms, d_out, d_cyc);
    run<FADD_FMNMX_MIX>(sms, d_out, d_cyc);
    run<LDS128_BCAST>(sms, d_out, d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
