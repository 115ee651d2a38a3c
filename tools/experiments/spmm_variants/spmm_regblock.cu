// CSR SpMM, register-blocked variant (v5):  C[n x m] = Mat * B[n x m], dense row-major.
//
// Same host-packed clusters as the TMA variant (spmm_blob.cuh): <= R mesh-neighbouring rows touching <= max_cols distinct
// columns.  A CTA owns one cluster; warp w owns the 64-column panel w of the result.  The cluster's matrix entries are
// scattered once into a DENSE local block D[ncol][R] in shared memory (zeros where a row does not touch a column), and
// the product of the cluster is then   acc[r][lane's 2 columns] += D[j][r] * B[cols[j]][lane's 2 columns]   over the
// distinct columns j:
//   * every distinct B row is loaded ONCE per cluster, straight from global memory into registers (one coalesced
//     LDG.128 = 512 contiguous bytes per warp) -- B is never staged in shared memory, which was the limiter of the
//     cp.async-panel kernel (LSU data pipe 70 % busy, profiles/r01_spmm_tma_vs_staged.md);
//   * the accumulators of all R rows live in registers with static indices; the matrix values arrive as broadcast
//     LDS.128 (two rows per load): R/2 shared-memory wavefronts per distinct column instead of 5 per matrix entry;
//   * a software ring of DEPTH loads per warp keeps ~DEPTH*512 B in flight per warp for the HBM latency.
// The price is R*ncol DFMA pairs per lane and panel instead of nnz (P1 mass matrix, R = 16: ~4x the useful flops;
// ~0.11 ms of FP64 pipe at cfg2 against the 0.175 ms HBM floor of the algorithmic bytes).
// Narrow tail panels (<= 16 column pairs, e.g. the last 10 of m = 266 columns) fold several row groups into one warp:
// lane = (row group g, column pair c) and a lane accumulates rows g, g+G, g+2G, ... so the tail costs ~1/G of a panel.
#include "../../include/hfb200.h"
#include "hfb_common.cuh"
#include "spmm_blob.cuh"
#include <cstdlib>

namespace hfb {

constexpr int RB_MAX_WARPS = 8;

template <int R, int DEPTH>
__global__ void __launch_bounds__(RB_MAX_WARPS * 32)
    csr_spmm_regblock_kernel(int m, int max_cols, SpmmBlobLayout L, const unsigned char* __restrict__ blobs,
                             const double* __restrict__ B, unsigned ldb2 /* ldb / 2 */, double* __restrict__ C, long long ldc) {
    extern __shared__ __align__(16) unsigned char smem_rb[];
    unsigned char* sBlob = smem_rb;                                   // raw cluster record (L.stride bytes)
    double* sD = reinterpret_cast<double*>(smem_rb + L.stride);       // dense local block [max_cols][R]
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;

    {   // record -> shared memory (fixed size: one dependent global latency for all metadata), zero the dense block
        const int4* src = reinterpret_cast<const int4*>(blobs + (size_t)blockIdx.x * L.stride);
        int4* dst = reinterpret_cast<int4*>(sBlob);
        for (int i = tid; i < (L.stride >> 4); i += nthr) dst[i] = __ldg(src + i);
        double2* z = reinterpret_cast<double2*>(sD);
        for (int i = tid; i < (max_cols * R) >> 1; i += nthr) z[i] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    const int* hdr = reinterpret_cast<const int*>(sBlob);
    const int nrow = hdr[0], ncol = hdr[1], nent = hdr[2];
    {
        const int4* ent = reinterpret_cast<const int4*>(sBlob + L.off_ent);
        for (int e = tid; e < nent; e += nthr) {
            const int4 en = ent[e];  // {value lo, value hi, local column, local row}
            sD[en.z * R + en.w] = __hiloint2double(en.y, en.x);
        }
    }
    __syncthreads();
    const int* sCols = reinterpret_cast<const int*>(sBlob + L.off_cols);
    const int* sOut = reinterpret_cast<const int*>(sBlob + L.off_outrow);

    const int npair = (m + 1) >> 1;          // column pairs of the block
    const int npanel = (npair + 31) >> 5;
    for (int panel = warp; panel < npanel; panel += nwarps) {
        const int t = min(32, npair - panel * 32);   // column pairs of this panel
        if (t > 16) {
            // ---------------------------------------------------------------- full panel: lane = column pair
            const int c = panel * 64 + 2 * lane;
            // ldb >= m rounded up to even: a lane with c < m may read the full pair (the padding column of an odd m feeds
            // acc.y only, which is never stored); lanes past the block read column 0 and store nothing -> no branches
            const bool in2 = c + 1 < m, in1 = c < m;
            const double2* Bp = reinterpret_cast<const double2*>(B + (in1 ? c : 0));
            auto loadb = [&](int j) -> double2 { return __ldg(Bp + (unsigned long long)(unsigned)sCols[j] * ldb2); };
            double2 acc[R];
#pragma unroll
            for (int i = 0; i < R; ++i) acc[i] = make_double2(0.0, 0.0);
            double2 bq[DEPTH];
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) bq[d] = (d < ncol) ? loadb(d) : make_double2(0.0, 0.0);
            for (int j0 = 0; j0 < ncol; j0 += DEPTH) {
#pragma unroll
                for (int d = 0; d < DEPTH; ++d) {
                    const int j = j0 + d;
                    if (j < ncol) {
                        const double2 b = bq[d];
                        if (j + DEPTH < ncol) bq[d] = loadb(j + DEPTH);
                        const double2* dv = reinterpret_cast<const double2*>(sD + j * R);
#pragma unroll
                        for (int q = 0; q < R / 2; ++q) {
                            const double2 v = dv[q];  // same address in all lanes: one broadcast LDS.128, rows 2q, 2q+1
                            acc[2 * q].x = fma(v.x, b.x, acc[2 * q].x);
                            acc[2 * q].y = fma(v.x, b.y, acc[2 * q].y);
                            acc[2 * q + 1].x = fma(v.y, b.x, acc[2 * q + 1].x);
                            acc[2 * q + 1].y = fma(v.y, b.y, acc[2 * q + 1].y);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < R; ++i) {
                if (i < nrow) {
                    double* cp = C + (long long)sOut[i] * ldc + c;
                    if (in2) {
                        *reinterpret_cast<double2*>(cp) = acc[i];
                    } else if (in1) {
                        cp[0] = acc[i].x;
                    }
                }
            }
        } else {
            // ---------------------------------------------------------------- narrow panel: lane = (row group, column pair)
            constexpr int RT = (R + 1) / 2;
            const int G = 32 / t;                     // row groups folded into the warp (>= 2)
            const int g = lane / t, cc = lane - g * t;
            const bool active = g < G;
            const int c = panel * 64 + 2 * cc;
            const bool in2 = active && c + 1 < m, in1 = active && c < m;
            const int cnt = (nrow + G - 1) / G;       // rows per lane (<= RT)
            const double2* Bp = reinterpret_cast<const double2*>(B + (in1 ? c : 0));
            auto loadb = [&](int j) -> double2 { return __ldg(Bp + (unsigned long long)(unsigned)sCols[j] * ldb2); };
            double2 acc[RT];
#pragma unroll
            for (int i = 0; i < RT; ++i) acc[i] = make_double2(0.0, 0.0);
            double2 bq[DEPTH];
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) bq[d] = (d < ncol) ? loadb(d) : make_double2(0.0, 0.0);
            for (int j0 = 0; j0 < ncol; j0 += DEPTH) {
#pragma unroll
                for (int d = 0; d < DEPTH; ++d) {
                    const int j = j0 + d;
                    if (j < ncol) {
                        const double2 b = bq[d];
                        if (j + DEPTH < ncol) bq[d] = loadb(j + DEPTH);
                        const double* dv = sD + j * R;
#pragma unroll
                        for (int i = 0; i < RT; ++i) {
                            if (i < cnt) {
                                const int row = g + i * G;
                                const double v = row < R ? dv[row] : 0.0;
                                acc[i].x = fma(v, b.x, acc[i].x);
                                acc[i].y = fma(v, b.y, acc[i].y);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < RT; ++i) {
                const int row = g + i * G;
                if (i < cnt && row < nrow) {
                    double* cp = C + (long long)sOut[row] * ldc + c;
                    if (in2) {
                        *reinterpret_cast<double2*>(cp) = acc[i];
                    } else if (in1) {
                        cp[0] = acc[i].x;
                    }
                }
            }
        }
    }
}

template <int R, int DEPTH>
static int launch_regblock(int64_t nclusters, int m, int max_cols, const SpmmBlobLayout& L, const void* blobs, const double* B,
                           int64_t ldb, double* C, int64_t ldc, cudaStream_t stream) {
    const int npanel = ((m + 1) / 2 + 31) / 32;
    const int nwarps = npanel < RB_MAX_WARPS ? npanel : RB_MAX_WARPS;
    const size_t smem = (size_t)L.stride + (size_t)max_cols * R * sizeof(double);
    if (smem > 48 * 1024) {
        static size_t configured = 0;
        if (smem > configured) {
            cudaError_t e = cudaFuncSetAttribute(csr_spmm_regblock_kernel<R, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem);
            if (e != cudaSuccess) return (int)e;
            configured = smem;
        }
    }
    csr_spmm_regblock_kernel<R, DEPTH><<<(unsigned)nclusters, nwarps * 32, smem, stream>>>(
        m, max_cols, L, static_cast<const unsigned char*>(blobs), B, (unsigned)(ldb >> 1), C, ldc);
    ++g_launch_count;
    return (int)cudaGetLastError();
}

}  // namespace hfb

using namespace hfb;

extern "C" int hfb_csr_spmm_regblock(int64_t nclusters, int64_t m, const void* blobs, int32_t max_rows, int32_t max_cols,
                                     int32_t max_entries, const double* B, int64_t ldb, double* C, int64_t ldc, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nclusters <= 0 || m <= 0 || !blobs || !B || !C || B == C || max_rows <= 0 || max_cols <= 0 || max_cols > 128 ||
        max_entries <= 0)
        return HFB_E_BADARG;
    if (ldb < m + (m & 1) || ldc < m || ldb > 0x7fffffffLL) return HFB_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(B) & 15) || (reinterpret_cast<uintptr_t>(C) & 15) || (reinterpret_cast<uintptr_t>(blobs) & 15) ||
        (ldb & 1) || (ldc & 1))
        return HFB_E_ALIGN;
    if (nclusters > 0x7fffffffLL || m > 0x3fffffffLL) return HFB_E_UNSUPPORTED;
    if (max_rows > 32) return HFB_E_UNSUPPORTED;   // accumulators of a cluster must fit in registers
    const SpmmBlobLayout L = blob_layout(max_rows, max_cols, max_entries);
    if ((size_t)L.stride + (size_t)max_cols * 32 * sizeof(double) > 200 * 1024) return HFB_E_UNSUPPORTED;
    // loads in flight per warp (software ring); HFB_SPMM_RB_DEPTH = 4 / 8 selects the alternatives kept for tuning runs
    const char* dep_env = getenv("HFB_SPMM_RB_DEPTH");
    const int dep = dep_env ? atoi(dep_env) : 0;
#define HFB_RB_LAUNCH(R_, D_) launch_regblock<R_, D_>(nclusters, (int)m, max_cols, L, blobs, B, ldb, C, ldc, stream)
    if (max_rows <= 8) return dep == 4 ? HFB_RB_LAUNCH(8, 4) : dep == 12 ? HFB_RB_LAUNCH(8, 12) : HFB_RB_LAUNCH(8, 8);
    if (max_rows <= 12) return dep == 4 ? HFB_RB_LAUNCH(12, 4) : dep == 8 ? HFB_RB_LAUNCH(12, 8) : HFB_RB_LAUNCH(12, 6);
    if (max_rows <= 16) return dep == 4 ? HFB_RB_LAUNCH(16, 4) : dep == 8 ? HFB_RB_LAUNCH(16, 8) : HFB_RB_LAUNCH(16, 6);
    if (max_rows <= 24) return HFB_RB_LAUNCH(24, 4);
    return HFB_RB_LAUNCH(32, 4);
#undef HFB_RB_LAUNCH
}
