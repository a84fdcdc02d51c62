// microbench2.cu -- can one warp keep the FMA pipe busy with packed f32x2 ops while its
// ALU-pipe instructions (FMNMX / FSETP / FSEL) issue in the gaps?  Patterns of independent
// chains, 8 warps x 4 CTAs per SM (and 1 warp per SMSP with --one), reported as
// warp-instructions per clock per SM from CUDA-event time at the measured clock.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define ITERS 400000

enum Pat { P_FFMA2, P_FFMA, P_MNMX, P_MNMX3, P_SETPSEL, P_FFMA2_MNMX_1_1, P_FFMA2_MNMX_2_1, P_FFMA_MNMX_1_1,
           P_FFMA2_SETPSEL, P_FFMA2_DPMIX, P_FADD_MNMX, P_FFMA2_LDS, N_PAT };
static const char* kNames[] = {"ffma2", "ffma", "fmnmx (min/max alternating)", "fmnmx3", "fsetp+fsel",
                               "ffma2 : fmnmx = 1:1", "ffma2 : fmnmx = 2:1", "ffma : fmnmx = 1:1",
                               "ffma2 : (fsetp+fsel) = 2:2", "20 ffma2 : 2x(mnmx,setp,setp,sel,fadd)",
                               "fadd : fmnmx = 1:1", "4 ffma2 : 1 lds.128 bcast"};

template <int P>
__global__ void __launch_bounds__(256) probe(float* out, float seed, int* cnt)
{
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    unsigned long long p[8];
    float a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        a[k] = seed + k; b[k] = seed * 0.5f + k;
        p[k] = ((unsigned long long)__float_as_uint(a[k]) << 32) | __float_as_uint(b[k]);
    }
    const float c0 = seed * 1.0001f, c1 = seed * 0.9999f, c3 = seed * 3.0f;
    const unsigned long long c2 = ((unsigned long long)__float_as_uint(c0) << 32) | __float_as_uint(c1);
#define F2OP(k) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[k]) : "l"(c2))
#define F1OP(k) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(c0), "f"(c1))
#define MIN(k) asm volatile("min.f32 %0, %0, %1;" : "+f"(b[k]) : "f"(c3))
#define MAX(k) asm volatile("max.f32 %0, %0, %1;" : "+f"(b[k]) : "f"(c1))
#define MIN3(k) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(b[k]) : "f"(c3), "f"(a[k]))
#define SETPSEL(k) asm volatile("{.reg .pred q; setp.lt.f32 q, %0, %1; selp.f32 %0, %1, %2, q;}" : "+f"(b[k]) : "f"(a[k]), "f"(c0))
#define FADD(k) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(c0))
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (P == P_FFMA2) { F2OP(0); F2OP(1); F2OP(2); F2OP(3); F2OP(4); F2OP(5); F2OP(6); F2OP(7); }
        if (P == P_FFMA) { F1OP(0); F1OP(1); F1OP(2); F1OP(3); F1OP(4); F1OP(5); F1OP(6); F1OP(7); }
        if (P == P_MNMX) { MIN(0); MIN(1); MIN(2); MIN(3); MIN(4); MIN(5); MIN(6); MIN(7); MAX(0); MAX(1); MAX(2); MAX(3); MAX(4); MAX(5); MAX(6); MAX(7); }
        if (P == P_MNMX3) { MIN3(0); MIN3(1); MIN3(2); MIN3(3); MIN3(4); MIN3(5); MIN3(6); MIN3(7); }
        if (P == P_SETPSEL) { SETPSEL(0); SETPSEL(1); SETPSEL(2); SETPSEL(3); SETPSEL(4); SETPSEL(5); SETPSEL(6); SETPSEL(7); }
        if (P == P_FFMA2_MNMX_1_1) { F2OP(0); MIN(0); F2OP(1); MIN(1); F2OP(2); MIN(2); F2OP(3); MIN(3); F2OP(4); MAX(0); F2OP(5); MAX(1); F2OP(6); MAX(2); F2OP(7); MAX(3); }
        if (P == P_FFMA2_MNMX_2_1) { F2OP(0); F2OP(1); MIN(0); F2OP(2); F2OP(3); MIN(1); F2OP(4); F2OP(5); MAX(0); F2OP(6); F2OP(7); MAX(1); }
        if (P == P_FFMA_MNMX_1_1) { F1OP(0); MIN(0); F1OP(1); MIN(1); F1OP(2); MIN(2); F1OP(3); MIN(3); F1OP(4); MAX(0); F1OP(5); MAX(1); F1OP(6); MAX(2); F1OP(7); MAX(3); }
        if (P == P_FFMA2_SETPSEL) { F2OP(0); SETPSEL(0); F2OP(1); SETPSEL(1); F2OP(2); SETPSEL(2); F2OP(3); SETPSEL(3); }
        if (P == P_FFMA2_DPMIX) {
            F2OP(0); MIN(0); F2OP(1); SETPSEL(1); F2OP(2); MIN(2); F2OP(3); SETPSEL(3); F2OP(4); FADD(4); F2OP(5); SETPSEL(5);
            F2OP(6); FADD(6); F2OP(7); SETPSEL(7); F2OP(0); F2OP(1); F2OP(2); F2OP(3); F2OP(4); F2OP(5); F2OP(6); F2OP(7);
            F2OP(0); F2OP(1); F2OP(2); F2OP(3);
        }
        if (P == P_FADD_MNMX) { FADD(0); MIN(0); FADD(1); MIN(1); FADD(2); MIN(2); FADD(3); MIN(3); FADD(4); MAX(0); FADD(5); MAX(1); FADD(6); MAX(2); FADD(7); MAX(3); }
        if (P == P_FFMA2_LDS) { float4 v = sm[it & 63]; F2OP(0); F2OP(1); F2OP(2); F2OP(3); a[0] += v.x; }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k] + b[k] + __uint_as_float((unsigned)(p[k] >> 32)) + __uint_as_float((unsigned)p[k]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static const int kInstr[] = {8, 8, 16, 8, 16, 16, 12, 16, 12, 20 + 2 * 2 + 4 * 2 + 2, 16, 6};

template <int P>
void run(int sms, float* d_out, int threads, double clk_ghz)
{
    const int ctas = sms * (threads == 256 ? 4 : 1);
    probe<P><<<ctas, threads>>>(d_out, 1.0f, nullptr);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<P><<<ctas, threads>>>(d_out, 1.0f, nullptr);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double warps = (double)ctas * threads / 32;
    double total = warps * ITERS * kInstr[P];
    printf("%-42s %7.1f Gwarp-instr/s  = %5.2f /clk/SM at %.3f GHz  (%.3f ms)\n", kNames[P], total / (ms * 1e-3) / 1e9,
           total / (ms * 1e-3) / (sms * clk_ghz * 1e9), clk_ghz, ms);
}

int main(int argc, char** argv)
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double clk = khz / 1e6;
    int threads = (argc > 1 && !strcmp(argv[1], "--one")) ? 128 : 256;  // --one: 1 warp per SMSP
    printf("%s, %d SMs, max clock %.3f GHz, %s\n", prop.name, sms, clk, threads == 128 ? "1 warp / SMSP" : "8 warps / SMSP");
    float* d_out;
    cudaMalloc(&d_out, sizeof(float) * sms * 4 * 256);
    run<P_FFMA2>(sms, d_out, threads, clk);
    run<P_FFMA>(sms, d_out, threads, clk);
    run<P_MNMX>(sms, d_out, threads, clk);
    run<P_MNMX3>(sms, d_out, threads, clk);
    run<P_SETPSEL>(sms, d_out, threads, clk);
    run<P_FFMA2_MNMX_1_1>(sms, d_out, threads, clk);
    run<P_FFMA2_MNMX_2_1>(sms, d_out, threads, clk);
    run<P_FFMA_MNMX_1_1>(sms, d_out, threads, clk);
    run<P_FFMA2_SETPSEL>(sms, d_out, threads, clk);
    run<P_FFMA2_DPMIX>(sms, d_out, threads, clk);
    run<P_FADD_MNMX>(sms, d_out, threads, clk);
    run<P_FFMA2_LDS>(sms, d_out, threads, clk);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
