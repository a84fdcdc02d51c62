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

template <bool STRICT, bool UNITW, bool GSTATE>
static cudaError_t launch_one(const KernelArgs& a, int grid, size_t smem, cudaStream_t stream)
{
    auto kern = dtw_units_kernel<APD_DPAD, STRICT, UNITW, GSTATE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, 32, smem, stream>>>(a);
    return cudaGetLastError();
}

template <bool STRICT, bool UNITW, bool GSTATE>
static cudaError_t occupancy_one(size_t smem, int* blocks_per_sm)
{
    auto kern = dtw_units_kernel<APD_DPAD, STRICT, UNITW, GSTATE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, 32, smem);
}

#define APD_DISPATCH(FN, ...)                                                     \
    do {                                                                          \
        if (strict) {                                                             \
            if (unitw) { if (gstate) return FN<true, true, true>(__VA_ARGS__);    \
                         return FN<true, true, false>(__VA_ARGS__); }             \
            if (gstate) return FN<true, false, true>(__VA_ARGS__);                \
            return FN<true, false, false>(__VA_ARGS__);                           \
        }                                                                         \
        if (unitw) { if (gstate) return FN<false, true, true>(__VA_ARGS__);       \
                     return FN<false, true, false>(__VA_ARGS__); }                \
        if (gstate) return FN<false, false, true>(__VA_ARGS__);                   \
        return FN<false, false, false>(__VA_ARGS__);                              \
    } while (0)

cudaError_t APD_CAT(dtw_launch_, APD_DPAD)(const KernelArgs& a, bool strict, bool unitw,
                                           bool gstate, int grid, size_t smem,
                                           cudaStream_t stream)
{
    APD_DISPATCH(launch_one, a, grid, smem, stream);
}

cudaError_t APD_CAT(dtw_occupancy_, APD_DPAD)(bool strict, bool unitw, bool gstate, size_t smem,
                                              int* blocks_per_sm)
{
    APD_DISPATCH(occupancy_one, smem, blocks_per_sm);
}

}  // namespace apd
