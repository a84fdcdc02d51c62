// dtw_inst.cu -- instantiates the DTW unit kernels for one padded frame width.
// Compiled once per width with -DAPD_DPAD=<4|8|...|32> (see ../build.py) so the
// eight widths build in parallel.
#include "dtw_kernels.cuh"

#ifndef APD_DPAD
#error "compile with -DAPD_DPAD=<padded frame width>"
#endif

namespace apd {

#define APD_CAT2(a, b) a##b
#define APD_CAT(a, b) APD_CAT2(a, b)

template <bool STRICT, bool UNITW, int RING>
static cudaError_t launch_one(const KernelArgs& a, int grid, size_t smem, cudaStream_t stream)
{
    if (RING == RING_WIDE) {
        auto kern = dtw_units_wide_kernel<APD_DPAD, STRICT, UNITW>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, a.carveout);
        if (e != cudaSuccess) return e;
        kern<<<grid, 32 * WIDE_WARPS, smem, stream>>>(a);
        return cudaGetLastError();
    }
    if (RING == RING_TMEM) {
        auto kern = dtw_units_tmem_kernel<APD_DPAD, STRICT, UNITW>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, a.carveout);
        if (e != cudaSuccess) return e;
        kern<<<grid, 32 * TMEM_WARPS, smem, stream>>>(a);
        return cudaGetLastError();
    }
    auto kern = dtw_units_kernel<APD_DPAD, STRICT, UNITW, (RING == RING_TMEM || RING == RING_WIDE) ? RING_SMEM : RING>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, a.carveout);
    if (e != cudaSuccess) return e;
    kern<<<grid, 32, smem, stream>>>(a);
    return cudaGetLastError();
}

template <bool STRICT, bool UNITW, int RING>
static cudaError_t occupancy_one(size_t smem, int* blocks_per_sm)
{
    if (RING == RING_WIDE) {
        // one CTA per SM by construction (it allocates all 512 tensor-memory columns)
        auto kern = dtw_units_wide_kernel<APD_DPAD, STRICT, UNITW>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        *blocks_per_sm = 1;
        return cudaSuccess;
    }
    if (RING == RING_TMEM) {
        auto kern = dtw_units_tmem_kernel<APD_DPAD, STRICT, UNITW>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        // The occupancy calculator answers 1 for a kernel that allocates tensor memory (it cannot
        // see the column count); two allocations of TMEM_COLS columns fill the 512 columns of an
        // SM, so residency is bounded by that and by the register file.
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, kern);
        if (e != cudaSuccess) return e;
        const int regs_per_cta = ((fa.numRegs + 7) & ~7) * 32 * TMEM_WARPS;
        int by_regs = regs_per_cta > 0 ? 65536 / regs_per_cta : 1;
        int occ = 512 / TMEM_COLS;
        if (by_regs < occ) occ = by_regs;
        int dev = 0, smem_sm = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) == cudaSuccess) {
            const int by_smem = (int)(smem_sm / (smem + 1024));   // 1 KB per resident CTA is reserved by the system
            if (by_smem < occ) occ = by_smem;
        }
        *blocks_per_sm = occ < 1 ? 1 : occ;
        return cudaSuccess;
    }
    auto kern = dtw_units_kernel<APD_DPAD, STRICT, UNITW, (RING == RING_TMEM || RING == RING_WIDE) ? RING_SMEM : RING>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, 32, smem);
}

#define APD_DISPATCH_RING(FN, S, U, ...)                                          \
    do {                                                                          \
        if (ring == RING_WIDE) return FN<S, U, RING_WIDE>(__VA_ARGS__);           \
        if (ring == RING_TMEM) return FN<S, U, RING_TMEM>(__VA_ARGS__);           \
        if (ring == RING_GLOBAL) return FN<S, U, RING_GLOBAL>(__VA_ARGS__);       \
        return FN<S, U, RING_SMEM>(__VA_ARGS__);                                  \
    } while (0)

#define APD_DISPATCH(FN, ...)                                                     \
    do {                                                                          \
        if (strict) {                                                             \
            if (unitw) APD_DISPATCH_RING(FN, true, true, __VA_ARGS__);            \
            APD_DISPATCH_RING(FN, true, false, __VA_ARGS__);                      \
        }                                                                         \
        if (unitw) APD_DISPATCH_RING(FN, false, true, __VA_ARGS__);               \
        APD_DISPATCH_RING(FN, false, false, __VA_ARGS__);                         \
    } while (0)

cudaError_t APD_CAT(dtw_launch_, APD_DPAD)(const KernelArgs& a, bool strict, bool unitw,
                                           int ring, int grid, size_t smem,
                                           cudaStream_t stream)
{
    APD_DISPATCH(launch_one, a, grid, smem, stream);
}

cudaError_t APD_CAT(dtw_occupancy_, APD_DPAD)(bool strict, bool unitw, int ring, size_t smem,
                                              int* blocks_per_sm)
{
    APD_DISPATCH(occupancy_one, smem, blocks_per_sm);
}

}  // namespace apd
