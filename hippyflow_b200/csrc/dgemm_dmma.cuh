// FP64 tensor-core GEMM for tall-skinny operands: DMMA.8x8x4 fed by TMA through an mbarrier ring.
//
//   C[M x N] = alpha * op(A) * op(B)      (row-major C, deterministic split-K)
//
// CTA = 8 consumer warps + a 4-warp TMA producer group (384 threads, 1 CTA/SM; setmaxnreg moves the
// producers' registers to the consumers: 40 / 232).  CTA tile = 128 x (8*NT),
// K step 16.  Each consumer warp owns a 16 x (8*NT) strip: 2 x NT m8n8 accumulator tiles in registers
// (8*NT regs), so one B fragment load feeds 2 DMMAs and one A fragment feeds NT DMMAs.  Measured on
// B200 (tools/microbench_fp64.cu): DMMA.8x8x4 peaks at 37.07 TFLOP/s with 8 warps/SM; 4 warps reach 35.3.
//
// Shared-memory layout (per stage), all written by TMA with the 128-byte swizzle:
//   "KC" operand (K contiguous in global: A of NN/NT, B of NT): [rows][16 k]      -- one 128 B line per row
//   "XC" operand (M or N contiguous in global: A of TN, B of NN/TN): panels of 16 x: [panel][16 k-rows][16 x]
// 128B swizzle: 16-byte chunk c of line r lives at chunk c ^ (r & 7).
//
// Fragment mapping (g = lane>>2, t = lane&3).  A k-group is 8 consecutive k; its two k4 slices use the
// interleaved k sets {0,2,4,6} (s=0) and {1,3,5,7} (s=1): MMA k-index t <-> actual k = 2t + s.  With this
//   KC: one LDS.128 at (row rho(g), chunk 4q+t) yields the s=0 and s=1 fragments of one 8-row tile, where
//       rho(g) = 4*(g&1) + (g>>1) makes the eight lanes of a quarter-warp hit eight distinct chunks;
//   XC: one LDS.128 at (k-row 8q+2t+s, chunk g) yields the fragments of TWO 8-wide tiles (x = 2g and 2g+1)
//       for slice s; chunk g ^ (2t+s) is again distinct over a quarter-warp.
// Both are bank-conflict free, and both sides of a DMMA agree on which actual k sits in which k-index.
// The permutations are undone in the epilogue (see row_of / col_of).
#pragma once
#include "hfb_common.cuh"

namespace hfb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 16;
constexpr int GEMM_CONSUMER_WARPS = 8;
constexpr int GEMM_PRODUCER_WARPS = 4;  // a full warpgroup so setmaxnreg can hand its registers to the consumers
constexpr int GEMM_THREADS = (GEMM_CONSUMER_WARPS + GEMM_PRODUCER_WARPS) * 32;
constexpr int GEMM_REGS_PRODUCER = 40;
constexpr int GEMM_MAX_PEERS = 16;       // ranks of one NVLink domain addressed by the fused reduce-scatter epilogue
constexpr int GEMM_REGS_CONSUMER = 232;  // per SMSP: 2 consumer warps x 232 + 1 producer warp x 40 = 504 <= 512

struct GemmParams {
    int M, N, K;
    int m_tiles, n_tiles, splits;
    int kb_total;          // ceil(K / 16)
    double alpha;          // applied here only when splits == 1
    double* C;             // output (splits == 1) or split workspace (splits > 1)
    long long ldc;         // leading dimension of C / workspace
    long long split_stride;  // elements between consecutive split slabs in the workspace
    int vec_store;         // 1 if C base is 16-byte aligned and ldc even (16-byte stores allowed); 2 (peer epilogue): 32-byte aligned, ldc % 4 == 0
    int symmetric;         // 1: the result is symmetric (M == N); tiles strictly below the diagonal are skipped
    int accumulate;        // 1: C += alpha * op(A) op(B)  (applied here when splits == 1, else by the split-K reduce)
    // strided-batch extension (hfb_dgemm_batched); the operands are 3-D tensor maps whose third coordinate is the sample
    int batch;             // independent outputs C_b = alpha op(A_b) op(B_b), b < batch (1 for a plain GEMM)
    int kfold;             // samples folded into the K loop: C = alpha sum_z op(A_z) op(B_z), z < kfold (1 for a plain GEMM)
    int kb_per;            // k-blocks per sample; kb_total = kb_per * kfold
    int a_batched, b_batched;  // operand has a sample axis (third coordinate = sample) or is shared (third coordinate 0)
    long long strideC;     // elements between consecutive C_b
    int b_upper;           // 1: B (K x N, NN/TN layouts) is upper triangular: k-blocks past the tile's last column are skipped
    // fused reduce-scatter over peer memory (hfb_dgemm_peer): rows [o*peer_rows, (o+1)*peer_rows) of the result are stored
    // into peer_dst[o] -- this rank's slot inside rank o's exchange buffer, a peer-mapped (NVLink) or local address -- with
    // leading dimension ldc and rows counted from the start of the block.  peer_rows is a multiple of GEMM_BM, so a CTA's
    // tile has ONE destination.  0 = off (plain store into C).
    int peer_rows;
    double* peer_dst[GEMM_MAX_PEERS];
};

template <int LAYOUT, int NT>
struct GemmCfg {
    static constexpr bool A_KC = (LAYOUT == 0 || LAYOUT == 2);
    static constexpr bool B_KC = (LAYOUT == 2);
    static constexpr int BN = NT * 8;
    static constexpr int NPB = (NT + 1) / 2;  // 16-wide B panels (XC)
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 8;
    static constexpr int B_BYTES = B_KC ? BN * GEMM_BK * 8 : NPB * GEMM_BK * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES_RAW = (200 * 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ int rho(int g) { return ((g & 1) << 2) | (g >> 1); }

template <int LAYOUT, int NT, bool PEER = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
dgemm_dmma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const GemmParams p) {
    using Cfg = GemmCfg<LAYOUT, NT>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment: the swizzle pattern is a function of the shared address bits 4..9
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // tile / split decomposition: n fastest so CTAs sharing an A tile are co-scheduled, then m, then split
    // (CTAs of one split share the same K range of B -> they hit in L2).
    int bid = blockIdx.x;
    const int n_tile = bid % p.n_tiles;
    bid /= p.n_tiles;
    const int m_tile = bid % p.m_tiles;
    bid /= p.m_tiles;
    const int split = bid % p.splits;
    const int bat = bid / p.splits;  // sample of an independent-output batch (0 for a plain GEMM)
    const int kb_base = p.kb_total / p.splits, kb_rem = p.kb_total % p.splits;
    const int kb_begin = split * kb_base + (split < kb_rem ? split : kb_rem);
    int kb_count = kb_base + (split < kb_rem ? 1 : 0);
    const int m0 = m_tile * GEMM_BM;
    const int n0 = n_tile * Cfg::BN;
    if (p.b_upper) {  // B[k][j] = 0 for k > j: columns n0 .. n0+BN-1 only need k <= n0+BN-1 (uniform per CTA)
        const int kb_last = (min(n0 + Cfg::BN, p.N) + GEMM_BK - 1) / GEMM_BK;
        kb_count = max(0, min(kb_count, kb_last - kb_begin));
    }
    if (p.symmetric && n0 + Cfg::BN <= m0) return;  // whole tile below the diagonal: filled by the mirror kernel

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], GEMM_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp >= GEMM_CONSUMER_WARPS) {
        // ------------------------------------------------------------------ TMA producer warpgroup
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GEMM_REGS_PRODUCER));
        if (warp == GEMM_CONSUMER_WARPS && lane == 0) {
            prefetch_tmap(&mapA);
            prefetch_tmap(&mapB);
            for (int it = 0; it < kb_count; ++it) {
                const int stage = it % STAGES;
                const uint32_t phase = (it / STAGES) & 1;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
                uint8_t* sB = sA + Cfg::A_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                // k-block -> (sample, k offset): with kfold > 1 the K loop runs over the samples as well
                int kbg = kb_begin + it, z = bat;
                if (p.kfold > 1) {
                    z = kbg / p.kb_per;
                    kbg -= z * p.kb_per;
                }
                const int k0 = kbg * GEMM_BK;
                const int zA = p.a_batched ? z : 0, zB = p.b_batched ? z : 0;
                if (Cfg::A_KC) {
                    tma_load_3d(sA, &mapA, &full_bar[stage], k0, m0, zA);  // box {16 k, 128 rows, 1}
                } else {
#pragma unroll
                    for (int pnl = 0; pnl < GEMM_BM / 16; ++pnl)  // box {16 m, 16 k-rows, 1}
                        tma_load_3d(sA + pnl * (GEMM_BK * 128), &mapA, &full_bar[stage], m0 + 16 * pnl, k0, zA);
                }
                if (Cfg::B_KC) {
                    tma_load_3d(sB, &mapB, &full_bar[stage], k0, n0, zB);  // box {16 k, BN rows, 1}
                } else {
#pragma unroll
                    for (int pnl = 0; pnl < Cfg::NPB; ++pnl)  // box {16 n, 16 k-rows, 1}
                        tma_load_3d(sB + pnl * (GEMM_BK * 128), &mapB, &full_bar[stage], n0 + 16 * pnl, k0, zB);
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- DMMA consumers
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GEMM_REGS_CONSUMER));
    const int g = lane >> 2, t = lane & 3;
    const int rg = rho(g);
    double acc[2][NT][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int jn = 0; jn < NT; ++jn) acc[j][jn][0] = acc[j][jn][1] = 0.0;

    const uint32_t smem_base = smem_u32(smem);
    // per-thread constant byte offsets inside a stage
    // A, KC: row 16*warp + 8*j + rho(g), chunk (4q + t) ^ rho(g)
    // A, XC: panel `warp`, k-row 8q + 2t + s, chunk g ^ (2t + s)
    uint32_t a_off[2][2];  // KC: [j][q] ; XC: [q][s]
    if (Cfg::A_KC) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q)
                a_off[j][q] = (16 * warp + 8 * j + rg) * 128 + (((4 * q + t) ^ rg) << 4);
    } else {
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int s = 0; s < 2; ++s)
                a_off[q][s] = warp * (GEMM_BK * 128) + (8 * q + 2 * t + s) * 128 + ((g ^ (2 * t + s)) << 4);
    }
    // B, XC: panel pp, k-row 8q+2t+s, chunk g ^ (2t+s):   b_off[q][s] + pp * 2048
    // B, KC: row 8*jn + rho(g), chunk (4q+t) ^ rho(g):    b_off[q][0] + jn * 1024
    uint32_t b_off[2][2];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int s = 0; s < 2; ++s)
            b_off[q][s] = Cfg::B_KC ? (uint32_t)(rg * 128 + (((4 * q + t) ^ rg) << 4))
                                    : (uint32_t)((8 * q + 2 * t + s) * 128 + ((g ^ (2 * t + s)) << 4));

    uint32_t b_half[2][2];  // XC, odd NT: last tile, k-row 8q+2t+s, 8-byte slot of column g
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int s = 0; s < 2; ++s)
            b_half[q][s] = (uint32_t)((8 * q + 2 * t + s) * 128 + (((g >> 1) ^ (2 * t + s)) << 4) + ((g & 1) << 3));

    for (int it = 0; it < kb_count; ++it) {
        const int stage = it % STAGES;
        const uint32_t phase = (it / STAGES) & 1;
        mbar_wait(&full_bar[stage], phase);
        const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
        const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            double a[2][2];  // [j][s]
            if (Cfg::A_KC) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const double2 v = lds128(sA + a_off[j][q]);
                    a[j][0] = v.x;
                    a[j][1] = v.y;
                }
            } else {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const double2 v = lds128(sA + a_off[q][s]);
                    a[0][s] = v.x;
                    a[1][s] = v.y;
                }
            }
            if (Cfg::B_KC) {
#pragma unroll
                for (int jn = 0; jn < NT; ++jn) {
                    const double2 v = lds128(sB + b_off[q][0] + jn * 1024);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        dmma884(acc[j][jn][0], acc[j][jn][1], a[j][0], v.x);
                        dmma884(acc[j][jn][0], acc[j][jn][1], a[j][1], v.y);
                    }
                }
            } else {
#pragma unroll
                for (int pp = 0; pp < NT / 2; ++pp) {
                    const double2 v0 = lds128(sB + b_off[q][0] + pp * 2048);
                    const double2 v1 = lds128(sB + b_off[q][1] + pp * 2048);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        dmma884(acc[j][2 * pp][0], acc[j][2 * pp][1], a[j][0], v0.x);
                        dmma884(acc[j][2 * pp][0], acc[j][2 * pp][1], a[j][1], v1.x);
                        dmma884(acc[j][2 * pp + 1][0], acc[j][2 * pp + 1][1], a[j][0], v0.y);
                        dmma884(acc[j][2 * pp + 1][0], acc[j][2 * pp + 1][1], a[j][1], v1.y);
                    }
                }
                if (NT & 1) {
                    // odd tile count: the last tile is the FIRST 8 columns of the last panel, direct mapping
                    // (MMA column g <-> column g), 8-byte loads; (g>>1) ^ (2t+s) is distinct over a half-warp.
                    const double w0 = lds64(sB + b_half[q][0] + (NT / 2) * 2048);
                    const double w1 = lds64(sB + b_half[q][1] + (NT / 2) * 2048);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        dmma884(acc[j][NT - 1][0], acc[j][NT - 1][1], a[j][0], w0);
                        dmma884(acc[j][NT - 1][0], acc[j][NT - 1][1], a[j][1], w1);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
    }

    // ---------------------------------------------------------------------- epilogue
    double* Cout = p.C + (long long)split * p.split_stride + (long long)bat * p.strideC;
    if (PEER) {
        // reduce-scatter by push: the tile goes straight from the accumulators into the owner's exchange buffer over NVLink
        const int owner = m0 / p.peer_rows;
        Cout = p.peer_dst[owner] - (long long)owner * p.peer_rows * p.ldc;
    }
    const double alpha = (p.splits == 1) ? p.alpha : 1.0;
    const bool accum = (p.splits == 1) && p.accumulate;
    auto put1 = [&](double* dst, double v) { *dst = accum ? (*dst + v) : v; };
    auto put2 = [&](double* dst, double v0, double v1) {
        double2 o = make_double2(v0, v1);
        if (accum) {
            const double2 c = *reinterpret_cast<const double2*>(dst);
            o.x += c.x;
            o.y += c.y;
        }
        *reinterpret_cast<double2*>(dst) = o;
    };
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int row = m0 + 16 * warp + (Cfg::A_KC ? (8 * j + rg) : (2 * g + j));
        if (row >= p.M) continue;
        double* crow = Cout + (long long)row * p.ldc;
        if (Cfg::B_KC) {
#pragma unroll
            for (int jn = 0; jn < NT; ++jn) {
                const int c0 = n0 + 8 * jn + t, c1 = c0 + 4;  // rho(2t) = t, rho(2t+1) = 4 + t
                if (c0 < p.N) put1(crow + c0, alpha * acc[j][jn][0]);
                if (c1 < p.N) put1(crow + c1, alpha * acc[j][jn][1]);
            }
        } else {
#pragma unroll
            for (int pp = 0; pp < NT / 2; ++pp) {
                // columns 16pp + 4t + {0,1,2,3} = tiles (2pp, 2pp+1) x fragment halves (0, 1)
                const int c = n0 + 16 * pp + 4 * t;
                const double v0 = alpha * acc[j][2 * pp][0], v1 = alpha * acc[j][2 * pp + 1][0];
                const double v2 = alpha * acc[j][2 * pp][1], v3 = alpha * acc[j][2 * pp + 1][1];
                if (PEER && p.vec_store == 2 && c + 3 < p.N) {
                    // one 256-bit store per lane: the four lanes of a quad write a full 128-byte line of the row, so a tile row
                    // crosses NVLink as whole lines instead of interleaved 16-byte pieces
                    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(crow + c)), "d"(v0),
                                 "d"(v1), "d"(v2), "d"(v3)
                                 : "memory");
                } else if (p.vec_store && c + 3 < p.N) {
                    put2(crow + c, v0, v1);
                    put2(crow + c + 2, v2, v3);
                } else {
                    if (c < p.N) put1(crow + c, v0);
                    if (c + 1 < p.N) put1(crow + c + 1, v1);
                    if (c + 2 < p.N) put1(crow + c + 2, v2);
                    if (c + 3 < p.N) put1(crow + c + 3, v3);
                }
            }
            if (NT & 1) {
                const int c = n0 + 8 * (NT - 1) + 2 * t;  // direct mapping: columns 2t, 2t+1 of the last tile
                const double v0 = alpha * acc[j][NT - 1][0], v1 = alpha * acc[j][NT - 1][1];
                if (p.vec_store && c + 1 < p.N) {
                    put2(crow + c, v0, v1);
                } else {
                    if (c < p.N) put1(crow + c, v0);
                    if (c + 1 < p.N) put1(crow + c + 1, v1);
                }
            }
        }
    }
}

}  // namespace hfb
