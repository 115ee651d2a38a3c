// CSR SpMM, cluster-dense DMMA variant (v6):  C[n x m] = Mat * B[n x m], dense row-major.
//
// Host-packed clusters (spmm_blob.cuh; the same clusters feed spmm_runs.cu): <= 8*RH mesh-neighbouring rows touching
// <= 4*MAXKS distinct columns.  A CTA owns one cluster.  The cluster's entries are scattered once into a dense local block
// D[rows][cols] (zeros where a row does not touch a column); each thread then keeps ITS DMMA A-fragments of D in registers
// for the whole kernel (MAXKS*RH doubles).  The distinct B rows of the cluster are staged panel by panel (64*NT columns)
// into a double-buffered shared-memory tile with cp.async (hardware-asynchronous: no scoreboard / register pressure, the
// limiter of the register-blocked variant), and the product of the cluster is   C_cluster = D * B_staged   as
// DMMA.8x8x4 tiles: warp w owns one 8-row half of the cluster and 8*RH*NT columns of the panel; a k-step is one
// conflict-free LDS (row pitch 64*NT + 4 doubles) feeding one DMMA per n-tile; all-zero 8x4 blocks of D are skipped.
//
// Why tensor-core tiles for an HBM-bound kernel: not for flops (DMMA issues at the DFMA rate) but for shared-memory
// traffic.  The cp.async-panel kernel reads one 16-byte matrix entry and one B pair per (entry, lane): ~5 shared
// wavefronts per matrix entry, LSU data pipe 70 % busy (profiles/r01_spmm_tma_vs_staged.md).  Here every staged B element
// is read from shared memory exactly once per cluster and the matrix never is: 8 wavefronts per staged 512-byte row
// segment (4 written by cp.async, 4 read as fragments) instead of ~26.
// Non-finite inputs: like any dense-block formulation, 0 * B[k][j] is evaluated for (row, k) pairs the matrix does not
// couple, so an Inf/NaN in row k of B reaches every row of the clusters that touch column k (SciPy would confine it to the
// coupled rows).  The stored bases and sketches of this path are finite.
#include "../../include/hfb200.h"
#include "hfb_common.cuh"
#include "spmm_blob.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace hfb {

constexpr int DM_WARPS = 8;

template <int RH, int MAXKS, int NT>
__global__ void __launch_bounds__(DM_WARPS * 32)
    csr_spmm_dmma_kernel(int m, SpmmBlobLayout L, const unsigned char* __restrict__ blobs, const double* __restrict__ B,
                         long long ldb, double* __restrict__ C, long long ldc) {
    constexpr int ROWS = 8 * RH, COLS = 4 * MAXKS;
    constexpr int CP = (COLS + 11) / 16 * 16 + 4;   // dense-block pitch >= COLS with CP mod 16 == 4: A-fragment LDS.64 conflict-free
    constexpr int PANEL = 64 * NT;          // columns per staged panel
    constexpr int PITCH = PANEL + 4;        // staged-row pitch: B-fragment LDS.64 conflict-free (2*PITCH mod 32 == 8)
    static_assert(CP >= COLS && CP % 16 == 4 && PITCH % 16 == 4, "fragment reads must be bank-conflict free");
    extern __shared__ __align__(16) unsigned char smem_dm[];
    unsigned char* sBlob = smem_dm;                                          // raw cluster record
    double* sD = reinterpret_cast<double*>(smem_dm + L.stride);              // [ROWS][CP]
    double* sB0 = sD + ROWS * CP;                                            // [COLS][PITCH] x 2
    double* sB1 = sB0 + COLS * PITCH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    {   // record -> shared memory (fixed size: one dependent global latency for all metadata); zero the dense block
        const int4* src = reinterpret_cast<const int4*>(blobs + (size_t)blockIdx.x * L.stride);
        int4* dst = reinterpret_cast<int4*>(sBlob);
        for (int i = tid; i < (L.stride >> 4); i += DM_WARPS * 32) dst[i] = __ldg(src + i);
        double2* z = reinterpret_cast<double2*>(sD);
        for (int i = tid; i < (ROWS * CP) >> 1; i += DM_WARPS * 32) z[i] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    const int* hdr = reinterpret_cast<const int*>(sBlob);
    const int nrow = hdr[0], ncol = hdr[1], nent = hdr[2];
    const int* sCols = reinterpret_cast<const int*>(sBlob + L.off_cols);
    const int* sOut = reinterpret_cast<const int*>(sBlob + L.off_outrow);
    const int KS = (ncol + 3) >> 2;                   // k-steps of this cluster
    const int width = m + (m & 1);                    // readable columns of a B row (ldb >= width)
    const int npanel = (m + PANEL - 1) / PANEL;

    // Staging map: warp w copies the distinct rows j = w, w + 8, ... (JW per warp); a lane moves 16 bytes, so one warp
    // instruction is 512 contiguous bytes of one row.  The row base pointers are formed once and kept in registers.
    // Rows >= ncol (k padding up to a multiple of 4) and columns >= width are zero-filled (src-size 0).
    constexpr int JW = (COLS + DM_WARPS - 1) / DM_WARPS;
    // Addresses are made opaque (asm volatile mov) so that they stay in registers: left alone, the compiler re-derives
    // every source/destination address from scratch for each copy (~28 instructions per cp.async in the first version).
    unsigned long long rowa[JW];                      // global address of this lane's 16 bytes of row j, next panel to stage
    unsigned rowok = 0, rowuse = 0;                   // bit i: row j = warp + 8 i is a real column / has to be written at all
#pragma unroll
    for (int i = 0; i < JW; ++i) {
        const int j = warp + DM_WARPS * i;
        const bool real = j < ncol;
        const double* ptr = B + (real ? (long long)sCols[j] * ldb : 0) + 2 * lane;
        asm volatile("mov.u64 %0, %1;" : "=l"(rowa[i]) : "l"(reinterpret_cast<unsigned long long>(ptr)));
        rowok |= (real ? 1u : 0u) << i;
        rowuse |= (j < (KS << 2) ? 1u : 0u) << i;
    }
    uint32_t dst0;
    asm volatile("mov.u32 %0, %1;" : "=r"(dst0) : "r"(smem_u32(sB0 + warp * PITCH + 2 * lane)));
    constexpr uint32_t BUF_BYTES = COLS * PITCH * 8;
    int cstage = 2 * lane;                            // this lane's first column in the next panel to stage
    auto stage = [&](int which) {                     // stages panels 0, 1, 2, ... in call order
        const uint32_t dst = dst0 + which * BUF_BYTES;
#pragma unroll
        for (int q = 0; q < NT; ++q) {
            const bool cvalid = cstage + 64 * q < width;
#pragma unroll
            for (int i = 0; i < JW; ++i) {
                if (rowuse >> i & 1u) {
                    const int bytes = (cvalid && (rowok >> i & 1u)) ? 16 : 0;      // 0: zero-fill, global memory untouched
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(
                                     dst + (uint32_t)(i * DM_WARPS * PITCH + 64 * q) * 8u),
                                 "l"(rowa[i] + 512ull * q), "r"(bytes)
                                 : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll
        for (int i = 0; i < JW; ++i) rowa[i] += PANEL * 8;
        cstage += PANEL;
    };
    stage(0);                                         // panel 0 is in flight while the dense block is built

    {
        const int4* ent = reinterpret_cast<const int4*>(sBlob + L.off_ent);
        for (int e = tid; e < nent; e += DM_WARPS * 32) {
            const int4 en = ent[e];                   // {value lo, value hi, local column, local row}
            sD[en.w * CP + en.z] = __hiloint2double(en.y, en.x);
        }
    }
    __syncthreads();
    // Work split: warp w owns row half h = w % RH (rows 8h .. 8h+7 of the cluster) and column group cg = w / RH of the
    // panel (TPW = RH*NT n-tiles = 8*TPW columns).  Its A fragments a[ks] = D[g + 8h][4 ks + t] stay in registers for all
    // panels.  nz: bit ks set when that 8 x 4 block of D has a nonzero -- k padding and the blocks this row half does not
    // reach (hfb_csr_pack_clusters numbers the local columns "upper half only, shared, lower half only") issue nothing.
    constexpr int TPW = RH * NT;
    static_assert(TPW == 1 || TPW == 2, "one or two n-tiles per warp and panel");
    const int h = warp % RH, cg = warp / RH;
    double a[MAXKS];
    unsigned nz = 0;
#pragma unroll
    for (int ks = 0; ks < MAXKS; ++ks) {
        a[ks] = sD[(g + 8 * h) * CP + 4 * ks + t];
        nz |= (__ballot_sync(0xffffffffu, a[ks] != 0.0) ? 1u : 0u) << ks;
    }
    // With two n-tiles per warp the pair is column-interleaved: tile A takes the even columns of the 16-column group, tile B
    // the odd ones, so one conflict-free LDS.128 at Bstaged[4 ks + t][16 cg + 2 g] feeds both DMMAs, and a lane ends up
    // with four consecutive result columns 16 cg + 4 t + {0: A.c0, 1: B.c0, 2: A.c1, 3: B.c1} of row g + 8h.
    const int r = g + 8 * h;
    const bool rowok_out = r < nrow;
    double* outp = C + (rowok_out ? (long long)sOut[r] * ldc : 0) + cg * TPW * 8 + (TPW == 2 ? 4 * t : 2 * t);
    uint32_t bfrag0;
    asm volatile("mov.u32 %0, %1;" : "=r"(bfrag0) : "r"(smem_u32(sB0 + t * PITCH + cg * TPW * 8 + (TPW == 2 ? 2 * g : g))));

    for (int panel = 0; panel < npanel; ++panel) {
        // single barrier per panel: after it, panel `panel` has landed for everybody AND everybody is done with the buffer
        // of panel - 1, which the copy of panel + 1 may now overwrite while this panel is multiplied
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (panel + 1 < npanel) stage((panel + 1) & 1);
        const int cbase = panel * PANEL + cg * TPW * 8;
        if (cbase < m) {
            uint32_t bcur;
            asm volatile("mov.u32 %0, %1;" : "=r"(bcur) : "r"(bfrag0 + (panel & 1) * BUF_BYTES));
            double* cp = outp + panel * PANEL;
            if (TPW == 2) {
                double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < MAXKS; ++ks) {
                    if (nz >> ks & 1u) {
                        const double2 b = lds128(bcur + (uint32_t)(ks * 4 * PITCH) * 8u);
                        dmma884(a0, a1, a[ks], b.x);
                        dmma884(b0, b1, a[ks], b.y);
                    }
                }
                if (rowok_out) {
                    const int c = cbase + 4 * t;
                    if (c + 1 < m) {
                        *reinterpret_cast<double2*>(cp) = make_double2(a0, b0);
                    } else if (c < m) {
                        cp[0] = a0;
                    }
                    if (c + 3 < m) {
                        *reinterpret_cast<double2*>(cp + 2) = make_double2(a1, b1);
                    } else if (c + 2 < m) {
                        cp[2] = a1;
                    }
                }
            } else {
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < MAXKS; ++ks) {
                    if (nz >> ks & 1u) {
                        const double b = lds64(bcur + (uint32_t)(ks * 4 * PITCH) * 8u);
                        dmma884(a0, a1, a[ks], b);
                    }
                }
                if (rowok_out) {
                    const int c = cbase + 2 * t;
                    if (c + 1 < m) {
                        *reinterpret_cast<double2*>(cp) = make_double2(a0, a1);
                    } else if (c < m) {
                        cp[0] = a0;
                    }
                }
            }
        }
    }
}

// column groups per warp of the fragment kernel (W <= 64 * SLAB_NG) and the cp.async group wait it uses
constexpr int SLAB_NG = 5;          // column groups per warp: W <= 64 * SLAB_NG

template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int RH, int MAXKS, int NT>
static int launch_dmma(int64_t nclusters, int m, const SpmmBlobLayout& L, const void* blobs, const double* B, int64_t ldb,
                       double* C, int64_t ldc, cudaStream_t stream) {
    constexpr int ROWS = 8 * RH, COLS = 4 * MAXKS, CP = (COLS + 11) / 16 * 16 + 4, PITCH = 64 * NT + 4;
    const size_t smem = (size_t)L.stride + sizeof(double) * ((size_t)ROWS * CP + 2 * (size_t)COLS * PITCH);
    static size_t configured[64] = {0};  // per device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (smem > 48 * 1024 && smem > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(csr_spmm_dmma_kernel<RH, MAXKS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = smem;
    }
    csr_spmm_dmma_kernel<RH, MAXKS, NT><<<(unsigned)nclusters, DM_WARPS * 32, smem, stream>>>(
        m, L, static_cast<const unsigned char*>(blobs), B, ldb, C, ldc);
    ++g_launch_count;
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------- fragment-blob variant
// The row-slab kernel without its prologue.  hfb_csr_pack_clusters_frag stores, per cluster, the dense block ALREADY in DMMA
// A-fragment order (afrag[ks][h][lane] = D[(lane >> 2) + 8h][4 ks + (lane & 3)]), the nonzero-block masks of the two row
// halves, the distinct columns and the result rows.  A thread reads its fragments with one coalesced load per k-step straight
// from global memory (the 8 warps of a CTA read the same 256-byte lines: L1 hits), so there is no record staging, no
// zero-fill / scatter of a dense block and no barrier before the B rows are requested: the cp.asyncs of a cluster depend on
// a single global-memory round trip (its column list).  Costs ~0.2 KB of extra matrix bytes per cluster row.
struct FragBlobLayout {
    int off_outrow, off_cols, off_afrag, stride;   // bytes; header = int32[8] {nrow, ncol, nz[0], nz[1], 0, 0, 0, 0}
};
static FragBlobLayout frag_layout(int rh, int maxks) {
    FragBlobLayout F;
    F.off_outrow = 32;
    F.off_cols = F.off_outrow + 4 * 8 * rh;
    F.off_afrag = round_up(F.off_cols + 4 * 4 * maxks, 256);     // fragment lines 256-byte aligned
    F.stride = round_up(F.off_afrag + 8 * 32 * rh * maxks, 256);
    return F;
}
static bool frag_shape(int max_rows, int max_cols, int& rh, int& maxks) {
    if (max_rows <= 0 || max_cols <= 0 || max_rows > 16 || max_cols > 48) return false;
    rh = max_rows <= 8 ? 1 : 2;
    maxks = max_cols <= 16 ? 4 : max_cols <= 24 ? 6 : max_cols <= 32 ? 8 : 12;
    return true;
}

// NG = column groups per warp (W <= 64 * NG): 5 for whole rows of up to 320 columns (3 CTAs/SM), 3 and 2 for chunks of up to
// 192 / 128 columns, whose smaller accumulator sets and row buffers let 4 / 5 CTAs share an SM.
template <int RH, int MAXKS, int NG>
__global__ void __launch_bounds__(DM_WARPS * 32, NG >= 5 ? 3 : NG == 3 ? 4 : 5)
    csr_spmm_dmma_frag_kernel(int m, int W, int st256, FragBlobLayout F, const unsigned char* __restrict__ blobs,
                              const double* __restrict__ B, long long ldb, double* __restrict__ C, long long ldc) {
    constexpr int COLS = 4 * MAXKS;
    constexpr int TILEW = 8 * RH;
    constexpr int JW = (COLS + DM_WARPS - 1) / DM_WARPS;
    constexpr int NA = 2 * RH;
    extern __shared__ __align__(16) unsigned char smem_dm[];
    double* sB = reinterpret_cast<double*>(smem_dm);                         // [COLS][pitch]
    const int pitch = W + 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int h = warp % RH, cg = warp / RH;
    const int c0 = blockIdx.y * W;
    const int wcur = min(W, m - c0);
    const unsigned char* blob = blobs + (size_t)blockIdx.x * F.stride;
    const int* hdr = reinterpret_cast<const int*>(blob);
    const int* gcols = reinterpret_cast<const int*>(blob + F.off_cols);
    const int nrow = __ldg(hdr), ncol = __ldg(hdr + 1);
    const unsigned nz = (unsigned)__ldg(hdr + 2 + h);
    const int KS = (ncol + 3) >> 2;
    const int width = m + (m & 1);
    const int npiece = (min(W, width - c0) + 1) >> 1;
    const int npiece_all = (((wcur + TILEW - 1) / TILEW) * TILEW) >> 1;
    int mycol[JW];
#pragma unroll
    for (int i = 0; i < JW; ++i) mycol[i] = __ldg(gcols + warp + DM_WARPS * i);      // padded with zeros past ncol
#pragma unroll
    for (int i = 0; i < JW; ++i) {
        const int j = warp + DM_WARPS * i;
        if (j < (KS << 2)) {
            const bool real = j < ncol;
            const double* src = B + (real ? (long long)mycol[i] * ldb + c0 : 0);
            const uint32_t dst = smem_u32(sB + j * pitch);
            for (int p = lane; p < npiece_all; p += 32) {
                const bool valid = real && p < npiece;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 16u * p),
                             "l"(src + (valid ? 2 * p : 0)), "r"(valid ? 16 : 0)
                             : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const double* afrag = reinterpret_cast<const double*>(blob + F.off_afrag) + h * 32 + lane;
    double a[MAXKS];
#pragma unroll
    for (int ks = 0; ks < MAXKS; ++ks) a[ks] = (nz >> ks & 1u) ? __ldg(afrag + ks * RH * 32) : 0.0;
    const int r = g + 8 * h;
    const int orow = r < nrow ? __ldg(reinterpret_cast<const int*>(blob + F.off_outrow) + r) : -1;

    constexpr int GSTRIDE = (DM_WARPS / RH) * TILEW;
    const int ngrp = (wcur > cg * TILEW) ? (wcur - cg * TILEW + GSTRIDE - 1) / GSTRIDE : 0;
    double acc[NG][NA];
#pragma unroll
    for (int i = 0; i < NG; ++i)
#pragma unroll
        for (int q = 0; q < NA; ++q) acc[i][q] = 0.0;
    uint32_t bfrag;
    asm volatile("mov.u32 %0, %1;" : "=r"(bfrag) : "r"(smem_u32(sB + t * pitch + cg * TILEW + (RH == 2 ? 2 * g : g))));
    const uint32_t kstep_bytes = (uint32_t)(4 * pitch) * 8u;
    auto ksteps = [&](int ks) {
        if (nz >> ks & 1u) {
            const uint32_t bk = bfrag + ks * kstep_bytes;
#pragma unroll
            for (int i = 0; i < NG; ++i) {
                if (i < ngrp) {
                    if (RH == 2) {
                        const double2 b = lds128(bk + (uint32_t)(i * GSTRIDE) * 8u);
                        dmma884(acc[i][0], acc[i][1], a[ks], b.x);
                        dmma884(acc[i][NA - 2], acc[i][NA - 1], a[ks], b.y);
                    } else {
                        const double b = lds64(bk + (uint32_t)(i * GSTRIDE) * 8u);
                        dmma884(acc[i][0], acc[i][1], a[ks], b);
                    }
                }
            }
        }
    };
#pragma unroll
    for (int i = 0; i < JW; ++i) {
        switch (JW - 1 - i) {
            case 0: cp_async_wait_group<0>(); break;
            case 1: cp_async_wait_group<1>(); break;
            case 2: cp_async_wait_group<2>(); break;
            case 3: cp_async_wait_group<3>(); break;
            case 4: cp_async_wait_group<4>(); break;
            default: cp_async_wait_group<5>(); break;
        }
        __syncthreads();
        if (2 * i < MAXKS) ksteps(2 * i);
        if (2 * i + 1 < MAXKS) ksteps(2 * i + 1);
    }
    if (orow >= 0) {
        double* outp = C + (long long)orow * ldc + c0 + cg * TILEW + (RH == 2 ? 4 * t : 2 * t);
#pragma unroll
        for (int i = 0; i < NG; ++i) {
            if (i < ngrp) {
                double* cp = outp + i * GSTRIDE;
                const int cc = c0 + cg * TILEW + i * GSTRIDE + (RH == 2 ? 4 * t : 2 * t);
                if (RH == 2) {
                    if (st256 && cc + 3 < m) {
                        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(cp)),
                                     "d"(acc[i][0]), "d"(acc[i][NA - 2]), "d"(acc[i][1]), "d"(acc[i][NA - 1])
                                     : "memory");
                    } else {
                        if (cc + 1 < m) {
                            *reinterpret_cast<double2*>(cp) = make_double2(acc[i][0], acc[i][NA - 2]);
                        } else if (cc < m) {
                            cp[0] = acc[i][0];
                        }
                        if (cc + 3 < m) {
                            *reinterpret_cast<double2*>(cp + 2) = make_double2(acc[i][1], acc[i][NA - 1]);
                        } else if (cc + 2 < m) {
                            cp[2] = acc[i][1];
                        }
                    }
                } else {
                    if (cc + 1 < m) {
                        *reinterpret_cast<double2*>(cp) = make_double2(acc[i][0], acc[i][1]);
                    } else if (cc < m) {
                        cp[0] = acc[i][0];
                    }
                }
            }
        }
    }
}

template <int RH, int MAXKS, int NG>
static int launch_dmma_frag_ng(int64_t nclusters, int m, int W, const FragBlobLayout& F, const void* blobs, const double* B,
                               int64_t ldb, double* C, int64_t ldc, cudaStream_t stream) {
    constexpr int COLS = 4 * MAXKS;
    const size_t smem = sizeof(double) * (size_t)COLS * (W + 4);
    if (smem > 227 * 1024 || W > 64 * NG) return HFB_E_UNSUPPORTED;
    static size_t configured[64] = {0};  // per device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (smem > 48 * 1024 && smem > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(csr_spmm_dmma_frag_kernel<RH, MAXKS, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = smem;
    }
    const int nchunk = (m + W - 1) / W;
    if (nchunk > 65535) return HFB_E_UNSUPPORTED;
    const int st256 = ((reinterpret_cast<uintptr_t>(C) & 31) == 0 && (ldc & 3) == 0) ? 1 : 0;
    dim3 grid((unsigned)nclusters, (unsigned)nchunk);
    csr_spmm_dmma_frag_kernel<RH, MAXKS, NG><<<grid, DM_WARPS * 32, smem, stream>>>(
        m, W, st256, F, static_cast<const unsigned char*>(blobs), B, ldb, C, ldc);
    ++g_launch_count;
    return (int)cudaGetLastError();
}

template <int RH, int MAXKS>
static int launch_dmma_frag(int64_t nclusters, int m, int W, const FragBlobLayout& F, const void* blobs, const double* B,
                            int64_t ldb, double* C, int64_t ldc, cudaStream_t stream) {
    if (W <= 128) return launch_dmma_frag_ng<RH, MAXKS, 2>(nclusters, m, W, F, blobs, B, ldb, C, ldc, stream);
    if (W <= 192) return launch_dmma_frag_ng<RH, MAXKS, 3>(nclusters, m, W, F, blobs, B, ldb, C, ldc, stream);
    return launch_dmma_frag_ng<RH, MAXKS, 5>(nclusters, m, W, F, blobs, B, ldb, C, ldc, stream);
}


// ---------------------------------------------------------------------------------------------------- ring-pipelined variant
// The fragment-record kernel with the per-CTA latency chain (column list -> row copies -> DMMAs -> stores) taken apart by
// warp specialisation.  One resident CTA per SM walks clusters blockIdx.x, blockIdx.x + gridDim.x, ... :
//   * one producer warp keeps a ring of 2-8 cluster buffers full: for every cluster it posts the byte count on the slot's FULL
//     mbarrier and issues one TMA linear copy (cp.async.bulk, SASS UBLKCP) for the fragment record and one per distinct B row
//     (lane j copies row j; the column lists are fetched three clusters ahead, in three register sets used in turn by a loop
//     unrolled by hand -- rotating them with moves would expose one DRAM latency per cluster);
//   * RING_CONSUMERS = 16 consumer warps (2 row halves x 8 column groups) wait for a slot, take their A fragments, masks and
//     result rows from the record in shared memory, run the k-steps, release the slot with one arrive per warp on EMPTY, and
//     store.
// Measured on B200 (profiles/r02_spmm_ring.md, clock64 sections): the TMA engine retires one small linear copy per ~65 cycles
// per SM (33 requests = ~2150 cycles per cluster whatever the row length), a DMMA.8x8x4 that accumulates into the result of
// the previous one issues ~250 cycles after it, and at m = 266 the 16 consumer warps keep the FP64 pipe ~56 % busy in the
// multiply phase (2150 cycles per cluster for 408 DMMAs); the period of the ring is the larger of the two, ~2900 cycles.
// Splitting the k-steps over two accumulator sets (RING_KSPLIT = 2) or the consumers over two clusters was measured slower
// (registers / fewer warps per cluster).  With RING_KSPLIT = 1 the summation order is that of the per-cluster kernel: results
// are bitwise equal to hfb_csr_spmm_dmma_frag.  TMA copies cannot zero-fill: the ring is zeroed once, and rows past a
// cluster's column count keep finite earlier data that only meets zero fragment entries.
constexpr int RING_CONSUMERS = 16;
constexpr int RING_THREADS = (RING_CONSUMERS + 1) * 32;
constexpr int RING_CG = RING_CONSUMERS / 2;   // column groups (two row halves)
constexpr int RING_NG = 3;                    // 16-column tiles per consumer warp: W <= 16 * RING_CG * RING_NG = 384
constexpr int RING_KSPLIT = 1;                // independent accumulator sets (k-step parity); 2 measured slower (spills)
constexpr int RING_MAX_SLOTS = 8;

template <int MAXKS>
__global__ void __launch_bounds__(RING_THREADS, 1)
    csr_spmm_ring_kernel(int m, int W, int st256, int nclusters, int nslots, int slot_bytes, FragBlobLayout F,
                         const unsigned char* __restrict__ blobs, const double* __restrict__ B, long long ldb,
                         double* __restrict__ C, long long ldc, unsigned long long* prof) {
    constexpr int RH = 2;
    constexpr int COLS = 4 * MAXKS;
    constexpr int TILEW = 8 * RH;
    extern __shared__ __align__(128) unsigned char smem_ring[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_ring);
    uint64_t* empty = full + RING_MAX_SLOTS;
    unsigned char* slots = smem_ring + 128;
    const int pitch = W + 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < nslots; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], RING_CONSUMERS);
        }
        mbar_fence_init();
    }
    for (int i = tid; i < nslots * slot_bytes / 8; i += blockDim.x) reinterpret_cast<double*>(slots)[i] = 0.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int width = m + (m & 1);
    const int npiece = (min(W, width) + 1) >> 1;                       // 16-byte pieces of a B row that exist

    if (warp == RING_CONSUMERS) {
        // ------------------------------------------------------------------ producer
        int c = blockIdx.x;
        auto fetch = [&](int cl, int& col, int& col2, int& nc) {
            col = col2 = nc = 0;
            if (cl < nclusters) {
                const unsigned char* bl = blobs + (size_t)cl * F.stride;
                nc = __ldg(reinterpret_cast<const int*>(bl) + 1);
                col = __ldg(reinterpret_cast<const int*>(bl + F.off_cols) + lane);
                if (COLS > 32 && lane + 32 < COLS) col2 = __ldg(reinterpret_cast<const int*>(bl + F.off_cols) + lane + 32);
            }
        };
        int colA, colB, colC, c2A, c2B, c2C, nA, nB, nC;
        fetch(c, colA, c2A, nA);
        fetch(c + (int)gridDim.x, colB, c2B, nB);
        fetch(c + 2 * (int)gridDim.x, colC, c2C, nC);
        int it = 0;
        const uint32_t rowbytes = (uint32_t)npiece * 16u;
        auto one_cluster = [&](int& colX, int& col2X, int& nX) {
            const int s = it % nslots;
            const uint32_t ph = (uint32_t)(it / nslots) & 1u;
            const int cur_col = colX, cur_col2 = col2X, cur_ncol = nX;
            const unsigned char* blob = blobs + (size_t)c * F.stride;
            fetch(c + 3 * (int)gridDim.x, colX, col2X, nX);
            const long long tp0 = prof ? clock64() : 0;
            mbar_wait(&empty[s], ph ^ 1u);
            if (prof && lane == 0) atomicAdd(prof + 0, (unsigned long long)(clock64() - tp0));
            unsigned char* st = slots + (size_t)s * slot_bytes;
            double* sB = reinterpret_cast<double*>(st + F.stride);
            if (lane == 0) {
                mbar_arrive_expect_tx(&full[s], (uint32_t)F.stride + (uint32_t)cur_ncol * rowbytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(st)),
                             "l"(reinterpret_cast<uint64_t>(blob)), "r"((uint32_t)F.stride), "r"(smem_u32(&full[s]))
                             : "memory");
            }
            __syncwarp();
            if (lane < cur_ncol)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(sB + lane * pitch)),
                             "l"(reinterpret_cast<uint64_t>(B + (long long)cur_col * ldb)), "r"(rowbytes), "r"(smem_u32(&full[s]))
                             : "memory");
            if (COLS > 32 && lane + 32 < cur_ncol)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(sB + (lane + 32) * pitch)),
                             "l"(reinterpret_cast<uint64_t>(B + (long long)cur_col2 * ldb)), "r"(rowbytes), "r"(smem_u32(&full[s]))
                             : "memory");
        };
        while (true) {
            if (c >= nclusters) break;
            one_cluster(colA, c2A, nA);
            c += gridDim.x;
            ++it;
            if (c >= nclusters) break;
            one_cluster(colB, c2B, nB);
            c += gridDim.x;
            ++it;
            if (c >= nclusters) break;
            one_cluster(colC, c2C, nC);
            c += gridDim.x;
            ++it;
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    const int g = lane >> 2, t = lane & 3;
    const int h = warp & 1, cg = warp >> 1;
    const int ntiles = (min(W, m) + TILEW - 1) / TILEW;
    const uint32_t kstep_bytes = (uint32_t)(4 * pitch) * 8u;
    int it = 0;
    for (int c = blockIdx.x; c < nclusters; c += gridDim.x, ++it) {
        const int s = it % nslots;
        const uint32_t ph = (uint32_t)(it / nslots) & 1u;
        const long long tc0 = prof ? clock64() : 0;
        mbar_wait(&full[s], ph);
        const long long tc1 = prof ? clock64() : 0;
        const unsigned char* st = slots + (size_t)s * slot_bytes;
        const int* hdr = reinterpret_cast<const int*>(st);
        const int nrow = hdr[0];
        const unsigned nz = (unsigned)hdr[2 + h];
        const double* afrag = reinterpret_cast<const double*>(st + F.off_afrag) + h * 32 + lane;
        double a[MAXKS];
#pragma unroll
        for (int ks = 0; ks < MAXKS; ++ks) a[ks] = (nz >> ks & 1u) ? afrag[ks * RH * 32] : 0.0;
        const int r = g + 8 * h;
        const int orow = r < nrow ? reinterpret_cast<const int*>(st + F.off_outrow)[r] : -1;
        // 16-column tiles of this warp: tile = i * RING_CG + ((cg + it) mod RING_CG); the rotation spreads the odd tile of a
        // width like 272 = 17 tiles over the warps from cluster to cluster
        const int rot = (cg + it) & (RING_CG - 1);
        double acc[RING_KSPLIT][RING_NG][4];
#pragma unroll
        for (int q = 0; q < RING_KSPLIT; ++q)
#pragma unroll
            for (int i = 0; i < RING_NG; ++i) acc[q][i][0] = acc[q][i][1] = acc[q][i][2] = acc[q][i][3] = 0.0;
        const uint32_t bbase = smem_u32(reinterpret_cast<const double*>(st + F.stride) + t * pitch + rot * TILEW + 2 * g);
        // software pipeline: the B fragments of k-step ks + 1 are requested before the DMMAs of k-step ks are issued
        // (lds128 / dmma884 are volatile asm and keep their program order, so the order below is the issue order)
        bool live[RING_NG];
#pragma unroll
        for (int i = 0; i < RING_NG; ++i) live[i] = i * RING_CG + rot < ntiles;
        double2 bcur[RING_NG], bnxt[RING_NG];
#pragma unroll
        for (int i = 0; i < RING_NG; ++i) {
            bcur[i] = make_double2(0.0, 0.0);
            if ((nz & 1u) && live[i]) bcur[i] = lds128(bbase + (uint32_t)(i * RING_CG * TILEW) * 8u);
        }
#pragma unroll
        for (int ks = 0; ks < MAXKS; ++ks) {
            if (ks + 1 < MAXKS) {
                const uint32_t bk = bbase + (ks + 1) * kstep_bytes;
#pragma unroll
                for (int i = 0; i < RING_NG; ++i) {
                    bnxt[i] = make_double2(0.0, 0.0);
                    if ((nz >> (ks + 1) & 1u) && live[i]) bnxt[i] = lds128(bk + (uint32_t)(i * RING_CG * TILEW) * 8u);
                }
            }
            if (nz >> ks & 1u) {
#pragma unroll
                for (int i = 0; i < RING_NG; ++i) {
                    if (live[i]) {
                        dmma884(acc[ks % RING_KSPLIT][i][0], acc[ks % RING_KSPLIT][i][1], a[ks], bcur[i].x);
                        dmma884(acc[ks % RING_KSPLIT][i][2], acc[ks % RING_KSPLIT][i][3], a[ks], bcur[i].y);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < RING_NG; ++i) bcur[i] = bnxt[i];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);      // everything this warp needs from the slot is in registers now
        if (orow >= 0) {
#pragma unroll
            for (int i = 0; i < RING_NG; ++i) {
                if (live[i]) {
                    double v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        v[e] = acc[0][i][e];
#pragma unroll
                        for (int q = 1; q < RING_KSPLIT; ++q) v[e] += acc[q][i][e];
                    }
                    const int cc = (i * RING_CG + rot) * TILEW + 4 * t;
                    double* cp = C + (long long)orow * ldc + cc;
                    if (st256 && cc + 3 < m) {
                        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(cp)),
                                     "d"(v[0]), "d"(v[2]), "d"(v[1]), "d"(v[3])
                                     : "memory");
                    } else {
                        if (cc + 1 < m) {
                            *reinterpret_cast<double2*>(cp) = make_double2(v[0], v[2]);
                        } else if (cc < m) {
                            cp[0] = v[0];
                        }
                        if (cc + 3 < m) {
                            *reinterpret_cast<double2*>(cp + 2) = make_double2(v[1], v[3]);
                        } else if (cc + 2 < m) {
                            cp[2] = v[1];
                        }
                    }
                }
            }
        }
        if (prof && lane == 0 && warp == 0) {
            atomicAdd(prof + 1, (unsigned long long)(tc1 - tc0));            // consumer: waiting for data
            atomicAdd(prof + 2, (unsigned long long)(clock64() - tc1));      // consumer: fragments + k-steps + stores
            atomicAdd(prof + 3, 1ull);
        }
    }
}

template <int MAXKS>
static int launch_ring(int64_t nclusters, int m, const FragBlobLayout& F, const void* blobs, const double* B, int64_t ldb,
                       double* C, int64_t ldc, cudaStream_t stream) {
    constexpr int COLS = 4 * MAXKS;
    const int W = (m + 15) / 16 * 16;
    if (W > 16 * RING_CG * RING_NG) return HFB_E_UNSUPPORTED;
    const int slot_bytes = round_up(F.stride + 8 * COLS * (W + 4), 128);
    int nslots = (227 * 1024 - 256) / slot_bytes;
    static const int max_slots = getenv("HFB_RING_SLOTS") ? atoi(getenv("HFB_RING_SLOTS")) : RING_MAX_SLOTS;
    if (nslots > max_slots) nslots = max_slots;
    if (nslots > RING_MAX_SLOTS) nslots = RING_MAX_SLOTS;
    if (nslots < 2) return HFB_E_UNSUPPORTED;
    const size_t smem = 128 + (size_t)nslots * slot_bytes;
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && smem > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(csr_spmm_ring_kernel<MAXKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = smem;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long gx = sms;
    if (gx > nclusters) gx = nclusters;
    const int st256 = ((reinterpret_cast<uintptr_t>(C) & 31) == 0 && (ldc & 3) == 0) ? 1 : 0;
    static unsigned long long* prof_dev = nullptr;
    unsigned long long* prof_buf = nullptr;
    if (getenv("HFB_RING_PROF")) {   // debug aid: cycles the producer waits for a slot / a consumer warp waits for data / works
        if (!prof_dev) cudaMalloc(&prof_dev, 4 * sizeof(unsigned long long));
        cudaMemsetAsync(prof_dev, 0, 4 * sizeof(unsigned long long), stream);
        prof_buf = prof_dev;
    }
    csr_spmm_ring_kernel<MAXKS><<<(unsigned)gx, RING_THREADS, smem, stream>>>(
        m, W, st256, (int)nclusters, nslots, slot_bytes, F, static_cast<const unsigned char*>(blobs), B, ldb, C, ldc, prof_buf);
    ++g_launch_count;
    if (prof_buf) {
        unsigned long long h[4];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[ring prof] m=%d slots=%d: producer wait-empty %.0f cyc/cluster, consumer wait-full %.0f, consumer work %.0f (%llu samples)\n",
                m, nslots, (double)h[0] / (double)nclusters, (double)h[1] / (double)(h[3] ? h[3] : 1), (double)h[2] / (double)(h[3] ? h[3] : 1), h[3]);
    }
    return (int)cudaGetLastError();
}

}  // namespace hfb

using namespace hfb;

extern "C" int64_t hfb_csr_cluster_blob_stride(int32_t max_rows, int32_t max_cols, int32_t max_entries) {
    if (max_rows <= 0 || max_cols <= 0 || max_entries <= 0) return HFB_E_BADARG;
    return blob_layout(max_rows, max_cols, max_entries).stride;
}

// HOST function: packs the clusters of (order, cluster_ptr) into the blob format above.  O(nnz).
extern "C" int hfb_csr_pack_clusters(int64_t n, const int32_t* rowptr, const int32_t* colind, const double* val,
                                     const int32_t* order, const int32_t* cluster_ptr, int64_t nclusters, int32_t max_rows,
                                     int32_t max_cols, int32_t max_entries, void* blobs_out) {
    if (n <= 0 || !rowptr || !colind || !val || !order || !cluster_ptr || nclusters <= 0 || max_rows <= 0 || max_cols <= 0 ||
        max_cols > 128 || max_entries <= 0 || !blobs_out)
        return HFB_E_BADARG;
    const SpmmBlobLayout L = blob_layout(max_rows, max_cols, max_entries);
    std::vector<int32_t> stamp((size_t)n, -1), local((size_t)n, 0);
    std::vector<int32_t> touched;                 // distinct columns of the current cluster, first-touch order
    std::vector<unsigned char> half((size_t)n, 0);  // bit 0: touched by cluster rows 0..7, bit 1: by rows >= 8
    unsigned char* out = static_cast<unsigned char*>(blobs_out);
    for (int64_t c = 0; c < nclusters; ++c) {
        unsigned char* blob = out + (size_t)c * L.stride;
        memset(blob, 0, (size_t)L.stride);
        int32_t* hdr = reinterpret_cast<int32_t*>(blob);
        int32_t* rowoff = reinterpret_cast<int32_t*>(blob + L.off_rowoff);
        int32_t* outrow = reinterpret_cast<int32_t*>(blob + L.off_outrow);
        int32_t* cols = reinterpret_cast<int32_t*>(blob + L.off_cols);
        unsigned char* ent = blob + L.off_ent;
        const int32_t s0 = cluster_ptr[c], s1 = cluster_ptr[c + 1];
        const int32_t nrow = s1 - s0;
        if (nrow <= 0 || nrow > max_rows) return HFB_E_BADARG;
        // pass 1: the distinct columns and which 8-row half of the cluster touches them
        touched.clear();
        for (int32_t r = 0; r < nrow; ++r) {
            const int32_t row = order[s0 + r];
            if (row < 0 || row >= n) return HFB_E_BADARG;
            for (int32_t j = rowptr[row]; j < rowptr[row + 1]; ++j) {
                const int32_t col = colind[j];
                if (col < 0 || col >= n) return HFB_E_BADARG;
                if (stamp[col] != (int32_t)c) {
                    if ((int32_t)touched.size() >= max_cols) return HFB_E_UNSUPPORTED;
                    stamp[col] = (int32_t)c;
                    half[col] = 0;
                    touched.push_back(col);
                }
                half[col] |= (r < 8) ? 1 : 2;
            }
        }
        // local numbering: columns only the upper half touches, then shared ones, then lower-half only (first-touch order
        // within a class).  Each half's nonzeros then fill a contiguous range of local columns, so the DMMA kernel skips
        // the 8 x 4 blocks of the dense cluster matrix outside that range; the other kernels do not care about the order.
        int32_t ncol = 0;
        for (unsigned char want : {(unsigned char)1, (unsigned char)3, (unsigned char)2})
            for (int32_t col : touched)
                if (half[col] == want) {
                    local[col] = ncol;
                    cols[ncol++] = col;
                }
        // pass 2: the entries
        int32_t nent = 0;
        for (int32_t r = 0; r < nrow; ++r) {
            const int32_t row = order[s0 + r];
            rowoff[r] = nent;
            outrow[r] = row;
            for (int32_t j = rowptr[row]; j < rowptr[row + 1]; ++j) {
                if (nent >= max_entries) return HFB_E_UNSUPPORTED;
                memcpy(ent + 16 * (size_t)nent, &val[j], 8);
                const int32_t l = local[colind[j]];
                memcpy(ent + 16 * (size_t)nent + 8, &l, 4);
                memcpy(ent + 16 * (size_t)nent + 12, &r, 4);
                ++nent;
            }
        }
        rowoff[nrow] = nent;
        hdr[0] = nrow;
        hdr[1] = ncol;
        hdr[2] = nent;
    }
    return 0;
}

extern "C" int hfb_csr_spmm_dmma(int64_t nclusters, int64_t m, const void* blobs, int32_t max_rows, int32_t max_cols,
                                 int32_t max_entries, const double* B, int64_t ldb, double* C, int64_t ldc, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nclusters <= 0 || m <= 0 || !blobs || !B || !C || B == C || max_rows <= 0 || max_cols <= 0 || max_entries <= 0)
        return HFB_E_BADARG;
    if (ldb < m + (m & 1) || ldc < m) return HFB_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(B) & 15) || (reinterpret_cast<uintptr_t>(C) & 15) || (reinterpret_cast<uintptr_t>(blobs) & 15) ||
        (ldb & 1) || (ldc & 1))
        return HFB_E_ALIGN;
    if (nclusters > 0x7fffffffLL || m > 0x3fffffffLL) return HFB_E_UNSUPPORTED;
    if (max_rows > 16 || max_cols > 48) return HFB_E_UNSUPPORTED;   // A fragments must fit in registers
    const SpmmBlobLayout L = blob_layout(max_rows, max_cols, max_entries);
    // panel width for clusters of <= 8 rows: HFB_SPMM_DMMA_NT = 1 (64 columns, default) / 2 (128 columns)
    const char* nt_env = getenv("HFB_SPMM_DMMA_NT");
    const int nt = nt_env ? atoi(nt_env) : 1;
#define HFB_DM_LAUNCH(RH_, KS_, NT_) launch_dmma<RH_, KS_, NT_>(nclusters, (int)m, L, blobs, B, ldb, C, ldc, stream)
#define HFB_DM_PICK(RH_, KS_) (nt == 2 ? HFB_DM_LAUNCH(RH_, KS_, 2) : HFB_DM_LAUNCH(RH_, KS_, 1))
#define HFB_DM_ONE(RH_, KS_) HFB_DM_LAUNCH(RH_, KS_, 1)
    if (max_rows <= 8) {
        if (max_cols <= 16) return HFB_DM_PICK(1, 4);
        if (max_cols <= 24) return HFB_DM_PICK(1, 6);
        if (max_cols <= 32) return HFB_DM_PICK(1, 8);
        return HFB_DM_PICK(1, 12);
    }
    if (max_cols <= 24) return HFB_DM_ONE(2, 6);      // two row halves: 64-column panels only (two n-tiles per warp)
    if (max_cols <= 32) return HFB_DM_ONE(2, 8);
    return HFB_DM_ONE(2, 12);
#undef HFB_DM_ONE
#undef HFB_DM_PICK
#undef HFB_DM_LAUNCH
}

// ---------------------------------------------------------------------------------------------------- fragment blobs (host)
extern "C" int64_t hfb_csr_frag_blob_stride(int32_t max_rows, int32_t max_cols) {
    int rh, maxks;
    if (!frag_shape(max_rows, max_cols, rh, maxks)) return HFB_E_UNSUPPORTED;
    return frag_layout(rh, maxks).stride;
}

// HOST function: packs the clusters of (order, cluster_ptr) into the fragment-blob format of csr_spmm_dmma_frag_kernel.
extern "C" int hfb_csr_pack_clusters_frag(int64_t n, const int32_t* rowptr, const int32_t* colind, const double* val,
                                          const int32_t* order, const int32_t* cluster_ptr, int64_t nclusters, int32_t max_rows,
                                          int32_t max_cols, void* blobs_out) {
    if (n <= 0 || !rowptr || !colind || !val || !order || !cluster_ptr || nclusters <= 0 || !blobs_out) return HFB_E_BADARG;
    int rh, maxks;
    if (!frag_shape(max_rows, max_cols, rh, maxks)) return HFB_E_UNSUPPORTED;
    const FragBlobLayout F = frag_layout(rh, maxks);
    std::vector<int32_t> stamp((size_t)n, -1), local((size_t)n, 0), touched;
    std::vector<unsigned char> half((size_t)n, 0);
    unsigned char* out = static_cast<unsigned char*>(blobs_out);
    for (int64_t c = 0; c < nclusters; ++c) {
        unsigned char* blob = out + (size_t)c * F.stride;
        memset(blob, 0, (size_t)F.stride);
        int32_t* hdr = reinterpret_cast<int32_t*>(blob);
        int32_t* outrow = reinterpret_cast<int32_t*>(blob + F.off_outrow);
        int32_t* cols = reinterpret_cast<int32_t*>(blob + F.off_cols);
        double* afrag = reinterpret_cast<double*>(blob + F.off_afrag);
        const int32_t s0 = cluster_ptr[c], nrow = cluster_ptr[c + 1] - s0;
        if (nrow <= 0 || nrow > max_rows) return HFB_E_BADARG;
        touched.clear();
        for (int32_t r = 0; r < nrow; ++r) {
            const int32_t row = order[s0 + r];
            if (row < 0 || row >= n) return HFB_E_BADARG;
            for (int32_t j = rowptr[row]; j < rowptr[row + 1]; ++j) {
                const int32_t col = colind[j];
                if (col < 0 || col >= n) return HFB_E_BADARG;
                if (stamp[col] != (int32_t)c) {
                    if ((int32_t)touched.size() >= max_cols) return HFB_E_UNSUPPORTED;
                    stamp[col] = (int32_t)c;
                    half[col] = 0;
                    touched.push_back(col);
                }
                half[col] |= (r < 8) ? 1 : 2;
            }
        }
        // local columns: upper-half only, shared, lower-half only (each half's nonzero blocks are then contiguous in k)
        int32_t ncol = 0;
        for (unsigned char want : {(unsigned char)1, (unsigned char)3, (unsigned char)2})
            for (int32_t col : touched)
                if (half[col] == want) {
                    local[col] = ncol;
                    cols[ncol++] = col;
                }
        uint32_t nz[2] = {0, 0};
        for (int32_t r = 0; r < nrow; ++r) {
            const int32_t row = order[s0 + r];
            outrow[r] = row;
            const int hh = r >> 3, g = r & 7;
            for (int32_t j = rowptr[row]; j < rowptr[row + 1]; ++j) {
                const int32_t l = local[colind[j]];
                const int ks = l >> 2, t = l & 3;
                afrag[((size_t)ks * rh + hh) * 32 + (g << 2 | t)] += val[j];      // += : duplicate entries sum, as in CSR
                if (val[j] != 0.0) nz[hh] |= 1u << ks;
            }
        }
        hdr[0] = nrow;
        hdr[1] = ncol;
        hdr[2] = (int32_t)nz[0];
        hdr[3] = (int32_t)nz[1];
    }
    return 0;
}

static int dmma_frag_dispatch(int64_t nclusters, int64_t m, const void* blobs, int32_t max_rows, int32_t max_cols,
                              int32_t chunk_cols, const double* B, int64_t ldb, double* C, int64_t ldc, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nclusters <= 0 || m <= 0 || !blobs || !B || !C || B == C || chunk_cols < 0) return HFB_E_BADARG;
    if (ldb < m + (m & 1) || ldc < m) return HFB_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(B) & 15) || (reinterpret_cast<uintptr_t>(C) & 15) || (reinterpret_cast<uintptr_t>(blobs) & 15) ||
        (ldb & 1) || (ldc & 1))
        return HFB_E_ALIGN;
    if (nclusters > 0x7fffffffLL || m > 0x3fffffffLL) return HFB_E_UNSUPPORTED;
    int rh, maxks;
    if (!frag_shape(max_rows, max_cols, rh, maxks)) return HFB_E_UNSUPPORTED;
    const FragBlobLayout F = frag_layout(rh, maxks);
    // chunk width: whole rows when they fit (<= 320 columns and <= ~100 KB of staged rows), else equal chunks
    int W = chunk_cols > 0 ? (chunk_cols + 15) / 16 * 16 : 64 * SLAB_NG;
    if (W > 64 * SLAB_NG) W = 64 * SLAB_NG;
    for (int nchunk = (int)((m + W - 1) / W);; ++nchunk) {
        W = (int)(((m + nchunk - 1) / nchunk + 15) / 16 * 16);
        if (8u * (size_t)(4 * maxks) * (W + 4) <= 100u * 1024u || W <= 16) break;
    }
#define HFB_DM_FRAG(RH_, KS_) launch_dmma_frag<RH_, KS_>(nclusters, (int)m, W, F, blobs, B, ldb, C, ldc, stream)
    if (rh == 1) {
        if (maxks == 4) return HFB_DM_FRAG(1, 4);
        if (maxks == 6) return HFB_DM_FRAG(1, 6);
        if (maxks == 8) return HFB_DM_FRAG(1, 8);
        return HFB_DM_FRAG(1, 12);
    }
    if (maxks <= 6) return HFB_DM_FRAG(2, 6);
    if (maxks == 8) return HFB_DM_FRAG(2, 8);
    return HFB_DM_FRAG(2, 12);
#undef HFB_DM_FRAG
}

extern "C" int hfb_csr_spmm_dmma_frag(int64_t nclusters, int64_t m, const void* blobs, int32_t max_rows, int32_t max_cols,
                                      int32_t chunk_cols, const double* B, int64_t ldb, double* C, int64_t ldc, void* stream) {
    return dmma_frag_dispatch(nclusters, m, blobs, max_rows, max_cols, chunk_cols, B, ldb, C, ldc, stream);
}

/* Ring-pipelined form of hfb_csr_spmm_dmma_frag (same records, same results bit for bit): clusters of two row halves
 * (8 < max_rows <= 16), blocks of up to 384 columns. */
extern "C" int hfb_csr_spmm_dmma_ring(int64_t nclusters, int64_t m, const void* blobs, int32_t max_rows, int32_t max_cols,
                                      const double* B, int64_t ldb, double* C, int64_t ldc, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nclusters <= 0 || m <= 0 || !blobs || !B || !C || B == C) return HFB_E_BADARG;
    if (ldb < m + (m & 1) || ldc < m) return HFB_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(B) & 15) || (reinterpret_cast<uintptr_t>(C) & 15) || (reinterpret_cast<uintptr_t>(blobs) & 15) ||
        (ldb & 1) || (ldc & 1))
        return HFB_E_ALIGN;
    if (nclusters > 0x7fffffffLL || m > 16 * RING_CG * RING_NG) return HFB_E_UNSUPPORTED;
    int rh, maxks;
    if (!frag_shape(max_rows, max_cols, rh, maxks) || rh != 2) return HFB_E_UNSUPPORTED;
    const FragBlobLayout F = frag_layout(rh, maxks);
    if (maxks <= 6) return launch_ring<6>(nclusters, (int)m, F, blobs, B, ldb, C, ldc, stream);
    if (maxks == 8) return launch_ring<8>(nclusters, (int)m, F, blobs, B, ldb, C, ldc, stream);
    return launch_ring<12>(nclusters, (int)m, F, blobs, B, ldb, C, ldc, stream);
}
