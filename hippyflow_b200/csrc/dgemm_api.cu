// Host side of hfb_dgemm: tensor-map encoding, tile/split selection, launch (see include/hfb200.h).
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/hfb200.h"
#include "dgemm_dmma.cuh"

namespace hfb {

long long g_launch_count = 0;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

// 3-D fp64 tensor map over `nbatch` row-major arrays `batch_stride` elements apart: `inner` contiguous elements per
// row, `outer` rows, leading dimension ld (elements); box = {16, box_rows, 1}; 128-byte swizzle; out-of-bounds reads
// give zeros (also for the rows past `outer` of ONE sample: a K or M tail never reads the next sample).  A plain
// (non-batched) operand is the nbatch = 1 case.
static int make_map(CUtensorMap* map, const double* base, long long inner, long long outer, long long ld,
                    int box_rows, long long nbatch = 1, long long batch_stride = 0) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return HFB_E_NODRIVER;
    if (nbatch <= 1 || batch_stride <= 0) {
        nbatch = 1;
        batch_stride = ld * outer;
    }
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)nbatch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)batch_stride * 8};
    cuuint32_t box[3] = {16, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

static const int kNT[] = {4, 9, 10, 14, 16, 17, 18};
static const int kNumNT = sizeof(kNT) / sizeof(int);

// N-tile width (in 8-column units): least padded columns, then fewest tiles.
static int choose_nt(long long N) {
    int best = kNT[0];
    long long best_pad = -1, best_tiles = 0;
    for (int i = 0; i < kNumNT; ++i) {
        const long long bn = 8LL * kNT[i];
        const long long tiles = (N + bn - 1) / bn;
        const long long pad = tiles * bn;
        if (best_pad < 0 || pad < best_pad || (pad == best_pad && tiles < best_tiles)) {
            best = kNT[i];
            best_pad = pad;
            best_tiles = tiles;
        }
    }
    return best;
}

static int sm_count() {
    static int cache[64] = {0};  // keyed by the current device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cache[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev] = n;
    }
    return cache[dev];
}

static int auto_splits(long long M, long long N, long long K, int symmetric = 0) {
    const int nt = choose_nt(N);
    long long tiles = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + 8 * nt - 1) / (8 * nt));
    if (symmetric) {  // count the tiles that are actually computed
        const long long mt = (M + GEMM_BM - 1) / GEMM_BM, ntl = (N + 8 * nt - 1) / (8 * nt);
        tiles = 0;
        for (long long i = 0; i < mt; ++i)
            for (long long j = 0; j < ntl; ++j)
                if (!(j * 8 * nt + 8 * nt <= i * GEMM_BM)) ++tiles;
    }
    const long long kb = (K + GEMM_BK - 1) / GEMM_BK;
    const int sms = sm_count();
    if (tiles >= 2LL * sms || kb < 64) return 1;
    long long smax = kb / 32;  // keep >= 32 k-blocks (512 k) per split so the pipeline fill is amortised
    if (smax > 128) smax = 128;
    if (smax < 1) smax = 1;
    int best = 1;
    double best_eff = 0.0;
    for (long long s = 1; s <= smax; ++s) {
        const long long ctas = tiles * s;
        const long long waves = (ctas + sms - 1) / sms;
        const double eff = (double)ctas / (double)(waves * sms);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = (int)s;
        }
        if (eff >= 0.97 && ctas >= 4LL * sms) return (int)s;
    }
    return best;
}

// Fixed-order sum of the split slabs: C = alpha * sum_s ws[s]  (bitwise reproducible).
__global__ void splitk_reduce_kernel(const double* __restrict__ ws, long long split_stride, int splits, long long ldw,
                                     double* __restrict__ C, long long ldc, int M, int N, double alpha, int accumulate) {
    const long long total = (long long)M * N;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / N;
        const int c = (int)(idx - r * N);
        const double* src = ws + r * ldw + c;
        double s = 0.0;
        for (int k = 0; k < splits; ++k) s += src[(long long)k * split_stride];
        C[r * ldc + c] = accumulate ? (C[r * ldc + c] + alpha * s) : alpha * s;
    }
}

// C[j][i] = C[i][j] for i < j: completes a symmetric result whose below-diagonal tiles were skipped.
__global__ void mirror_upper_kernel(double* __restrict__ C, long long ldc, int N) {
    __shared__ double tile[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;  // tile (bi, bj) of the upper triangle, bj >= bi
    if (bj < bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int i = bi * 32 + r, j = bj * 32 + tx;
        tile[r][tx] = (i < N && j < N) ? C[(long long)i * ldc + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = bj * 32 + r, i = bi * 32 + tx;  // write C[j][i] = tile[i - bi*32][j - bj*32]
        if (i < N && j < N && j > i) C[(long long)j * ldc + i] = tile[tx][r];
    }
}

// per-layout launchers: dgemm_inst.cu compiled with -DHFB_GEMM_LAYOUT={0,1,2}
int dgemm_launch_nn(int nt, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t s);
int dgemm_launch_tn(int nt, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t s);
int dgemm_launch_nt(int nt, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t s);

}  // namespace hfb

using namespace hfb;

extern "C" int hfb_version(void) { return 100; }
extern "C" int64_t hfb_launch_count(void) { return (int64_t)g_launch_count; }

extern "C" int hfb_dgemm_auto_splits(int layout, int64_t M, int64_t N, int64_t K) {
    if (layout < 0 || layout > 2 || M <= 0 || N <= 0 || K <= 0) return HFB_E_BADARG;
    return auto_splits(M, N, K);
}

extern "C" size_t hfb_dgemm_workspace_bytes(int layout, int64_t M, int64_t N, int64_t K, int splits) {
    if (layout < 0 || layout > 2 || M <= 0 || N <= 0 || K <= 0) return 0;
    if (splits == 0) splits = auto_splits(M, N, K);
    if (splits <= 1) return 0;
    const long long ldw = (N + 1) & ~1LL;
    return (size_t)splits * (size_t)M * (size_t)ldw * 8;
}

extern "C" int hfb_dgemm(int layout, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                         const double* B, int64_t ldb, double* C, int64_t ldc, void* workspace, size_t workspace_bytes,
                         int splits, void* stream_) {
    return hfb_dgemm_ex(layout, M, N, K, alpha, A, lda, B, ldb, C, ldc, workspace, workspace_bytes, splits, 0, stream_);
}

extern "C" size_t hfb_dgemm_ex_workspace_bytes(int layout, int64_t M, int64_t N, int64_t K, int splits, int flags) {
    if (layout < 0 || layout > 2 || M <= 0 || N <= 0 || K <= 0) return 0;
    if (splits == 0) splits = auto_splits(M, N, K, (flags & HFB_GEMM_SYMMETRIC) && M == N);
    if (splits <= 1) return 0;
    const long long ldw = (N + 1) & ~1LL;
    return (size_t)splits * (size_t)M * (size_t)ldw * 8;
}

// Common implementation of hfb_dgemm_ex and hfb_dgemm_batched.
//   mode 0: `batch` independent products C_b = alpha op(A_b) op(B_b) (+ C_b); a stride of 0 shares the operand.
//   mode 1: one product with the K loop folded over the samples, C = alpha sum_b op(A_b) op(B_b) (split-K over the folded
//           range, summed in fixed order).
static int gemm_impl(int layout, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda, int64_t strideA,
                     const double* B, int64_t ldb, int64_t strideB, double* C, int64_t ldc, int64_t strideC, int64_t batch,
                     int mode, void* workspace, size_t workspace_bytes, int splits, int flags, cudaStream_t stream,
                     double* const* peer_dst = nullptr, int npeers = 0, int64_t peer_rows = 0) {
    const int symmetric = ((flags & HFB_GEMM_SYMMETRIC) && M == N) ? 1 : 0;
    const int accumulate = (flags & HFB_GEMM_ACCUMULATE) ? 1 : 0;
    const int b_upper = (flags & HFB_GEMM_B_UPPER) ? 1 : 0;
    if (b_upper && (layout == HFB_NT || K != N || batch > 1)) return HFB_E_BADARG;
    if ((flags & HFB_GEMM_SYMMETRIC) && M != N) return HFB_E_BADARG;
    if (symmetric && accumulate) return HFB_E_UNSUPPORTED;  // the mirror pass would overwrite the accumulated half
    if (layout < 0 || layout > 2 || M <= 0 || N <= 0 || K <= 0 || !A || !B || !C || splits < 0 || batch < 1) return HFB_E_BADARG;
    if (mode != 0 && mode != 1) return HFB_E_BADARG;
    if (M > 0x7fffffffLL || N > 0x7fffffffLL || K > 0x7fffffffLL || batch > 0x7fffffffLL) return HFB_E_BADARG;
    if (strideA < 0 || strideB < 0 || strideC < 0) return HFB_E_BADARG;
    const long long a_inner = (layout == HFB_TN) ? M : K, a_outer = (layout == HFB_TN) ? K : M;
    const long long b_inner = (layout == HFB_NT) ? K : N, b_outer = (layout == HFB_NT) ? N : K;
    if (lda < a_inner || ldb < b_inner || ldc < N) return HFB_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (lda & 1) || (ldb & 1) ||
        (strideA & 1) || (strideB & 1))
        return HFB_E_ALIGN;
    if (reinterpret_cast<uintptr_t>(C) & 7) return HFB_E_ALIGN;
    const bool batched = batch > 1;
    if (batched) {
        if (symmetric) return HFB_E_UNSUPPORTED;
        if (mode == 0 && strideC < ldc * (M - 1) + N) return HFB_E_BADARG;  // outputs must not overlap
        if (mode == 1 && (strideA == 0 && strideB == 0)) return HFB_E_BADARG;
    }
    const int kfold = (batched && mode == 1) ? (int)batch : 1;
    const int nbat = (batched && mode == 0) ? (int)batch : 1;

    const int nt = choose_nt(N);
    const long long kb_per = (K + GEMM_BK - 1) / GEMM_BK;
    const long long kb_total = kb_per * kfold;
    if (kb_total > 0x7fffffffLL) return HFB_E_BADARG;
    if (nbat > 1) splits = 1;  // the sample axis already fills the GPU; outputs are written directly
    if (splits == 0) splits = auto_splits(M, N, kb_total * GEMM_BK, symmetric);
    if (splits > kb_total) splits = (int)kb_total;
    if (splits < 1) splits = 1;

    GemmParams p;
    p.M = (int)M;
    p.N = (int)N;
    p.K = (int)K;
    p.m_tiles = (int)((M + GEMM_BM - 1) / GEMM_BM);
    p.n_tiles = (int)((N + 8 * nt - 1) / (8 * nt));
    p.splits = splits;
    p.kb_total = (int)kb_total;
    p.alpha = alpha;
    p.batch = nbat;
    p.kfold = kfold;
    p.kb_per = (int)kb_per;
    p.a_batched = (batched && strideA > 0) ? 1 : 0;
    p.b_batched = (batched && strideB > 0) ? 1 : 0;
    p.strideC = nbat > 1 ? strideC : 0;
    const long long ldw = (N + 1) & ~1LL;
    if (splits > 1) {
        const size_t need = (size_t)splits * (size_t)M * (size_t)ldw * 8;
        if (!workspace || workspace_bytes < need) return HFB_E_WORKSPACE;
        if (reinterpret_cast<uintptr_t>(workspace) & 15) return HFB_E_ALIGN;
        p.C = (double*)workspace;
        p.ldc = ldw;
        p.split_stride = (long long)M * ldw;
    } else {
        p.C = C;
        p.ldc = ldc;
        p.split_stride = 0;
    }
    p.vec_store = ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && (p.ldc & 1) == 0 && (p.strideC & 1) == 0) ? 1 : 0;
    p.symmetric = symmetric;
    p.accumulate = accumulate;
    p.b_upper = b_upper;
    p.peer_rows = 0;
    memset(p.peer_dst, 0, sizeof(p.peer_dst));
    if (peer_dst) {
        // fused reduce-scatter epilogue: one destination per 128-row tile, direct stores only
        if (npeers < 1 || npeers > GEMM_MAX_PEERS || peer_rows <= 0 || (peer_rows % GEMM_BM) || peer_rows > 0x7fffffffLL ||
            (long long)npeers * peer_rows < M)
            return HFB_E_BADARG;
        if (splits != 1 || batched || symmetric || accumulate) return HFB_E_UNSUPPORTED;
        if (ldc & 1) return HFB_E_ALIGN;
        for (int o = 0; o < npeers; ++o) {
            if (!peer_dst[o]) return HFB_E_BADARG;
            if (reinterpret_cast<uintptr_t>(peer_dst[o]) & 15) return HFB_E_ALIGN;
            p.peer_dst[o] = peer_dst[o];
        }
        p.peer_rows = (int)peer_rows;
        bool wide = (ldc & 3) == 0;
        for (int o = 0; o < npeers; ++o) wide = wide && (reinterpret_cast<uintptr_t>(peer_dst[o]) & 31) == 0;
        p.vec_store = wide ? 2 : 1;
    }
    if ((long long)p.m_tiles * p.n_tiles * p.splits * nbat > 0x7fffffffLL) return HFB_E_BADARG;

    CUtensorMap mapA, mapB;
    int rc;
    const long long nA = p.a_batched ? batch : 1, nB = p.b_batched ? batch : 1;
    if (layout == HFB_TN) rc = make_map(&mapA, A, M, K, lda, 16, nA, strideA);
    else rc = make_map(&mapA, A, K, M, lda, GEMM_BM, nA, strideA);
    if (rc) return rc;
    if (layout == HFB_NT) rc = make_map(&mapB, B, K, N, ldb, 8 * nt, nB, strideB);
    else rc = make_map(&mapB, B, N, K, ldb, 16, nB, strideB);
    if (rc) return rc;

    if (layout == HFB_NN) rc = dgemm_launch_nn(nt, mapA, mapB, p, stream);
    else if (layout == HFB_TN) rc = dgemm_launch_tn(nt, mapA, mapB, p, stream);
    else rc = dgemm_launch_nt(nt, mapA, mapB, p, stream);
    if (rc) return rc;

    if (splits > 1) {
        const long long total = M * N;
        long long blocks = (total + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        splitk_reduce_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const double*)workspace, p.split_stride, splits, ldw,
                                                                  C, ldc, (int)M, (int)N, alpha, accumulate);
        ++g_launch_count;
        rc = (int)cudaGetLastError();
    }
    if (rc == 0 && symmetric) {
        const unsigned t = (unsigned)((N + 31) / 32);
        mirror_upper_kernel<<<dim3(t, t), 256, 0, stream>>>(C, ldc, (int)N);
        ++g_launch_count;
        rc = (int)cudaGetLastError();
    }
    return rc;
}

extern "C" int hfb_dgemm_ex(int layout, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                            const double* B, int64_t ldb, double* C, int64_t ldc, void* workspace,
                            size_t workspace_bytes, int splits, int flags, void* stream_) {
    return gemm_impl(layout, M, N, K, alpha, A, lda, 0, B, ldb, 0, C, ldc, 0, 1, 0, workspace, workspace_bytes, splits, flags,
                     (cudaStream_t)stream_);
}

extern "C" int hfb_dgemm_peer(int layout, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                              const double* B, int64_t ldb, double* const* slot_ptrs, int nranks, int64_t block_rows,
                              int64_t ld_slot, void* stream_) {
    if (!slot_ptrs || nranks < 1) return HFB_E_BADARG;
    // C is only a placeholder for the argument checks: every tile is redirected to slot_ptrs[row / block_rows]
    return gemm_impl(layout, M, N, K, alpha, A, lda, 0, B, ldb, 0, slot_ptrs[0], ld_slot, 0, 1, 0, nullptr, 0, 1, 0,
                     (cudaStream_t)stream_, slot_ptrs, nranks, block_rows);
}

extern "C" size_t hfb_dgemm_batched_workspace_bytes(int layout, int64_t M, int64_t N, int64_t K, int64_t batch, int mode) {
    if (layout < 0 || layout > 2 || M <= 0 || N <= 0 || K <= 0 || batch < 1) return 0;
    if (mode != 1 || batch == 1) return mode == 1 ? hfb_dgemm_ex_workspace_bytes(layout, M, N, K, 0, 0) : 0;
    const long long kb_total = ((K + GEMM_BK - 1) / GEMM_BK) * batch;
    if (kb_total > 0x7fffffffLL) return 0;
    const int splits = auto_splits(M, N, kb_total * GEMM_BK, 0);
    if (splits <= 1) return 0;
    const long long ldw = (N + 1) & ~1LL;
    return (size_t)splits * (size_t)M * (size_t)ldw * 8;
}

extern "C" int hfb_dgemm_batched(int layout, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                                 int64_t strideA, const double* B, int64_t ldb, int64_t strideB, double* C, int64_t ldc,
                                 int64_t strideC, int64_t batch, int mode, int flags, void* workspace, size_t workspace_bytes,
                                 void* stream_) {
    if (flags & HFB_GEMM_SYMMETRIC) return HFB_E_UNSUPPORTED;
    return gemm_impl(layout, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, batch, mode, workspace,
                     workspace_bytes, 0, flags, (cudaStream_t)stream_);
}
