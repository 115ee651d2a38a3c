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

// 2-D fp64 tensor map over a row-major array: `inner` contiguous elements per row, `outer` rows, leading
// dimension ld (elements); box = {16, box_rows}; 128-byte swizzle; out-of-bounds reads give zeros.
static int make_map(CUtensorMap* map, const double* base, long long inner, long long outer, long long ld,
                    int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return HFB_E_NODRIVER;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
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
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
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

extern "C" int hfb_dgemm_ex(int layout, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                            const double* B, int64_t ldb, double* C, int64_t ldc, void* workspace,
                            size_t workspace_bytes, int splits, int flags, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int symmetric = ((flags & HFB_GEMM_SYMMETRIC) && M == N) ? 1 : 0;
    const int accumulate = (flags & HFB_GEMM_ACCUMULATE) ? 1 : 0;
    if ((flags & HFB_GEMM_SYMMETRIC) && M != N) return HFB_E_BADARG;
    if (symmetric && accumulate) return HFB_E_UNSUPPORTED;  // the mirror pass would overwrite the accumulated half
    if (layout < 0 || layout > 2 || M <= 0 || N <= 0 || K <= 0 || !A || !B || !C || splits < 0) return HFB_E_BADARG;
    if (M > 0x7fffffffLL || N > 0x7fffffffLL || K > 0x7fffffffLL) return HFB_E_BADARG;
    const long long a_inner = (layout == HFB_TN) ? M : K, a_outer = (layout == HFB_TN) ? K : M;
    const long long b_inner = (layout == HFB_NT) ? K : N, b_outer = (layout == HFB_NT) ? N : K;
    if (lda < a_inner || ldb < b_inner || ldc < N) return HFB_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (lda & 1) || (ldb & 1))
        return HFB_E_ALIGN;
    if (reinterpret_cast<uintptr_t>(C) & 7) return HFB_E_ALIGN;

    const int nt = choose_nt(N);
    const long long kb_total = (K + GEMM_BK - 1) / GEMM_BK;
    if (splits == 0) splits = auto_splits(M, N, K, symmetric);
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
    p.vec_store = ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && (p.ldc & 1) == 0) ? 1 : 0;
    p.symmetric = symmetric;
    p.accumulate = accumulate;
    if ((long long)p.m_tiles * p.n_tiles * p.splits > 0x7fffffffLL) return HFB_E_BADARG;

    CUtensorMap mapA, mapB;
    int rc;
    if (layout == HFB_TN) rc = make_map(&mapA, A, M, K, lda, 16);
    else rc = make_map(&mapA, A, K, M, lda, GEMM_BM);
    if (rc) return rc;
    if (layout == HFB_NT) rc = make_map(&mapB, B, K, N, ldb, 8 * nt);
    else rc = make_map(&mapB, B, N, K, ldb, 16);
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
