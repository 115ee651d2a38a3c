// CSR SpMM, persistent TMA-fed variant (v4):  C[n x m] = Mat * B[n x m], dense row-major.
//
// The matrix is pre-packed on the HOST (hfb_csr_pack_clusters) into one fixed-stride "blob" per row cluster
// (<= max_rows mesh-neighbouring rows touching <= max_cols distinct columns, from hfb_csr_cluster_rows_capped):
//     int32 header[4] = {nrow, ncol, nent, 0}
//     int32 rowoff[max_rows + 1]   entry offsets of the cluster's rows (relative to the cluster's first entry)
//     int32 outrow[max_rows]       global row index of each cluster row (where its result goes in C)
//     int32 cols[max_cols]         the distinct global columns (= rows of B) the cluster touches
//     {double v; int32 l; int32 r} entries[nent]   value + cluster-LOCAL column index + cluster-LOCAL row index
// One persistent CTA per SM walks work items (cluster, column chunk).  A producer warp issues, per item, one bulk
// copy (cp.async.bulk, the TMA engine's linear mode; SASS UBLKCP) of the blob and one per distinct B row
// (cw*8 contiguous bytes) into a 3-stage shared-memory ring guarded by full/empty mbarriers; no thread of the
// consumer warps computes a global load address.  Consumers flatten the (row, column pair) space of the stage over
// their lanes: per matrix entry one LDS.128 of the entry (broadcast), one conflict-free LDS.128 of B, two DFMA.
// Every B row a cluster touches is fetched once per chunk (about 1.8 fetches per matrix row on P1 meshes instead
// of 7), and the ring keeps ~140 KB of loads in flight per SM, which is what the HBM latency needs.
#include "../../include/hfb200.h"
#include "hfb_common.cuh"
#include "spmm_blob.cuh"
#include <cstdlib>
#include <cstring>
#include <vector>

namespace hfb {

constexpr int V4_CONSUMER_WARPS = 16;
constexpr int V4_THREADS = (V4_CONSUMER_WARPS + 1) * 32;
constexpr int V4_MAX_STAGES = 8;
constexpr int V4_SMEM_BUDGET = 232448 - 128;  // opt-in dynamic shared memory per CTA minus the barrier block

struct SpmmV4Params {
    int m;          // columns of B / C
    int width;      // m rounded up to even (the padding column of an odd m is read, never written)
    int cw;         // chunk width in doubles (even)
    int nchunk;
    int nclusters;
    int nstages;
    int stage_bytes;
    int max_cols;
    SpmmBlobLayout L;
};

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(V4_THREADS, 1)
    csr_spmm_tma_kernel(const SpmmV4Params p, const unsigned char* __restrict__ blobs, const double* __restrict__ B,
                        long long ldb, double* __restrict__ C, long long ldc) {
    extern __shared__ __align__(128) unsigned char smem_v4[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_v4);
    uint64_t* empty = full + V4_MAX_STAGES;
    unsigned char* stages = smem_v4 + 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nstages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, V4_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const long long nitems = (long long)p.nclusters * p.nchunk;

    if (warp == V4_CONSUMER_WARPS) {
        // ------------------------------------------------------------------ producer warp
        int it = 0;
        for (long long item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            const int s = it % p.nstages;
            const uint32_t ph = (uint32_t)(it / p.nstages) & 1u;
            const int cluster = (int)(item / p.nchunk), chunk = (int)(item - (long long)cluster * p.nchunk);
            const unsigned char* blob = blobs + (size_t)cluster * p.L.stride;
            const int* hdr = reinterpret_cast<const int*>(blob);
            const int* gcols = reinterpret_cast<const int*>(blob + p.L.off_cols);
            // metadata of the NEXT item is fetched before the ring slot is waited for, so its latency is hidden
            const int ncol = __ldg(hdr + 1), nent = __ldg(hdr + 2);
            int mycol[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) mycol[q] = (lane + 32 * q < ncol) ? __ldg(gcols + lane + 32 * q) : 0;
            mbar_wait(empty + s, ph ^ 1u);
            unsigned char* st = stages + (size_t)s * p.stage_bytes;
            const int c_beg = chunk * p.cw;
            const int wcur = min(p.cw, p.width - c_beg);
            const uint32_t rowbytes = (uint32_t)wcur * 8u;
            const uint32_t blob_bytes = (uint32_t)(p.L.off_ent + 16 * nent);
            if (lane == 0) {
                mbar_arrive_expect_tx(full + s, blob_bytes + (uint32_t)ncol * rowbytes);
                bulk_g2s(st, blob, blob_bytes, full + s);
            }
            __syncwarp();
            double* sB = reinterpret_cast<double*>(st + p.L.stride);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = lane + 32 * q;
                if (j < ncol) bulk_g2s(sB + (size_t)j * p.cw, B + (long long)mycol[q] * ldb + c_beg, rowbytes, full + s);
            }
        }
    } else {
        // ------------------------------------------------------------------ consumer warps
        int it = 0;
        for (long long item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            const int s = it % p.nstages;
            const uint32_t ph = (uint32_t)(it / p.nstages) & 1u;
            const int cluster = (int)(item / p.nchunk), chunk = (int)(item - (long long)cluster * p.nchunk);
            const int c_beg = chunk * p.cw;
            const int wcur = min(p.cw, p.width - c_beg);
            const int U = wcur >> 1;  // column pairs of this chunk
            const uint32_t magic = (uint32_t)((0x100000000ULL + (unsigned)U - 1) / (unsigned)U);
            mbar_wait(full + s, ph);
            const unsigned char* st = stages + (size_t)s * p.stage_bytes;
            const int nrow = *reinterpret_cast<const int*>(st);
            const int* sRowoff = reinterpret_cast<const int*>(st + p.L.off_rowoff);
            const int* sOutrow = reinterpret_cast<const int*>(st + p.L.off_outrow);
            const uint32_t ent_addr = smem_u32(st + p.L.off_ent);
            const uint32_t sB_addr = smem_u32(st + p.L.stride);
            const uint32_t pitch = (uint32_t)p.cw * 8u;
            const int total = nrow * U;
            // the warp -> item-slice map rotates from stage to stage so that the remainder of total / (32 * warps)
            // does not always land on the same warps (warps run up to nstages - 1 stages apart)
            const int wslot = (warp + it * 5) % V4_CONSUMER_WARPS;
            for (int f = wslot * 32 + lane; f < total; f += V4_CONSUMER_WARPS * 32) {
                const int r = (U == 1) ? f : (int)__umulhi((unsigned)f, magic);  // f / U (exact for f, U < 2^16)
                const int u = f - r * U;
                const int beg = sRowoff[r], end = sRowoff[r + 1];
                const uint32_t bcol = sB_addr + (uint32_t)u * 16u;
                double2 acc = make_double2(0.0, 0.0);
#pragma unroll 4
                for (int j = beg; j < end; ++j) {
                    int vlo, vhi, l, pad_;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(vlo), "=r"(vhi), "=r"(l), "=r"(pad_)
                                 : "r"(ent_addr + (uint32_t)j * 16u));
                    const double v = __hiloint2double(vhi, vlo);
                    const double2 b = lds128(bcol + (uint32_t)l * pitch);
                    acc.x = fma(v, b.x, acc.x);
                    acc.y = fma(v, b.y, acc.y);
                }
                const int c = c_beg + 2 * u;
                double* cp = C + (long long)sOutrow[r] * ldc + c;
                if (c + 1 < p.m) {
                    *reinterpret_cast<double2*>(cp) = acc;
                } else if (c < p.m) {
                    cp[0] = acc.x;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
    }
}

static int num_sms_v4() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            n <= 0)
            n = 148;
    }
    return n;
}

// chunking: the fewest column chunks such that >= 3 ring stages fit in shared memory (2 as a last resort)
static bool plan_chunks(int m, int max_cols, const SpmmBlobLayout& L, SpmmV4Params& p) {
    const int width = m + (m & 1);
    for (int want = 3; want >= 2; --want) {
        for (int nchunk = 1; nchunk <= 64; ++nchunk) {
            int cw = (width + nchunk - 1) / nchunk;
            cw += cw & 1;
            const int stage = round_up(L.stride + max_cols * cw * 8, 128);
            if ((long long)stage * want <= V4_SMEM_BUDGET) {
                p.m = m;
                p.width = width;
                p.cw = cw;
                p.nchunk = (width + cw - 1) / cw;
                p.stage_bytes = stage;
                int ns = V4_SMEM_BUDGET / stage;
                p.nstages = ns > V4_MAX_STAGES ? V4_MAX_STAGES : ns;
                p.max_cols = max_cols;
                p.L = L;
                return true;
            }
        }
    }
    return false;
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

extern "C" int hfb_csr_spmm_tma(int64_t nclusters, int64_t m, const void* blobs, int32_t max_rows, int32_t max_cols,
                                int32_t max_entries, const double* B, int64_t ldb, double* C, int64_t ldc, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nclusters <= 0 || m <= 0 || !blobs || !B || !C || B == C || max_rows <= 0 || max_cols <= 0 || max_cols > 128 ||
        max_entries <= 0)
        return HFB_E_BADARG;
    const int64_t width = m + (m & 1);
    if (ldb < width || ldc < m) return HFB_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(B) & 15) || (reinterpret_cast<uintptr_t>(C) & 15) || (reinterpret_cast<uintptr_t>(blobs) & 15) ||
        (ldb & 1) || (ldc & 1))
        return HFB_E_ALIGN;
    if (nclusters > 0x3fffffffLL || m > 0x3fffffffLL) return HFB_E_UNSUPPORTED;
    SpmmV4Params p;
    if (!plan_chunks((int)m, max_cols, blob_layout(max_rows, max_cols, max_entries), p)) return HFB_E_UNSUPPORTED;
    p.nclusters = (int)nclusters;
    const size_t smem = 128 + (size_t)p.nstages * p.stage_bytes;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(csr_spmm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    const long long nitems = (long long)p.nclusters * p.nchunk;
    long long grid = num_sms_v4();
    if (grid > nitems) grid = nitems;
    csr_spmm_tma_kernel<<<(unsigned)grid, V4_THREADS, smem, stream>>>(p, static_cast<const unsigned char*>(blobs), B, ldb, C, ldc);
    ++g_launch_count;
    return (int)cudaGetLastError();
}
