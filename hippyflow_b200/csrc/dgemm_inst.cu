// Kernel instantiations of the DMMA GEMM for one operand layout (-DHFB_GEMM_LAYOUT=0|1|2), split into
// three translation units so they compile in parallel.
#include "../../include/hfb200.h"
#include "dgemm_dmma.cuh"

#ifndef HFB_GEMM_LAYOUT
#error "compile with -DHFB_GEMM_LAYOUT=0|1|2"
#endif

namespace hfb {

template <int LAYOUT, int NT, bool PEER>
static int launch_k(const CUtensorMap& mapA, const CUtensorMap& mapB, const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<LAYOUT, NT>;
    static bool configured[64] = {false};  // per device: the attribute belongs to the device's copy of the kernel
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(dgemm_dmma_kernel<LAYOUT, NT, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = true;
    }
    const long long grid = (long long)p.m_tiles * p.n_tiles * p.splits * p.batch;
    dgemm_dmma_kernel<LAYOUT, NT, PEER><<<(unsigned)grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(mapA, mapB, p);
    ++g_launch_count;
    return (int)cudaGetLastError();
}

// The peer-store epilogue (fused reduce-scatter of the lift Y = Xt^T W) exists for the TN layout only.
template <int LAYOUT, int NT>
static int launch(const CUtensorMap& mapA, const CUtensorMap& mapB, const GemmParams& p, cudaStream_t stream) {
    if (p.peer_rows > 0) {
        if constexpr (LAYOUT == 1) return launch_k<LAYOUT, NT, true>(mapA, mapB, p, stream);
        else return HFB_E_UNSUPPORTED;
    }
    return launch_k<LAYOUT, NT, false>(mapA, mapB, p, stream);
}

template <int LAYOUT>
static int dispatch_nt(int nt, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t s) {
    switch (nt) {
        case 4: return launch<LAYOUT, 4>(a, b, p, s);
        case 9: return launch<LAYOUT, 9>(a, b, p, s);
        case 10: return launch<LAYOUT, 10>(a, b, p, s);
        case 14: return launch<LAYOUT, 14>(a, b, p, s);
        case 16: return launch<LAYOUT, 16>(a, b, p, s);
        case 17: return launch<LAYOUT, 17>(a, b, p, s);
        case 18: return launch<LAYOUT, 18>(a, b, p, s);
    }
    return HFB_E_UNSUPPORTED;
}


#if HFB_GEMM_LAYOUT == 0
int dgemm_launch_nn(int nt, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t s) {
    return dispatch_nt<0>(nt, a, b, p, s);
}
#elif HFB_GEMM_LAYOUT == 1
int dgemm_launch_tn(int nt, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t s) {
    return dispatch_nt<1>(nt, a, b, p, s);
}
#else
int dgemm_launch_nt(int nt, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t s) {
    return dispatch_nt<2>(nt, a, b, p, s);
}
#endif

}  // namespace hfb
