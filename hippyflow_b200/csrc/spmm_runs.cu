// CSR SpMM, run-staged FMA variant ("runs"):  C[n x m] = Mat * B[n x m], dense row-major.
//
// What the ring-pipelined DMMA kernel (spmm_dmma.cu) measured (profiles/r02_spmm_ring.md): DRAM traffic is 1.04 x the
// algorithmic bytes, yet the kernel stops at 75 % of the HBM copy rate because (a) the dense-block formulation pays four
// times the FP64 issue slots the matrix entries need (DMMA.8x8x4 issues at the DFMA rate, 75 % of the 8 x 4 blocks of a
// cluster are non-zero but only 19 % of their entries) and (b) one TMA request per staged B row (33 per cluster) retires at
// ~65 cycles per request whatever its length.  This kernel removes both:
//   * the distinct columns of a mesh-neighbour cluster, sorted, fall into ~6 RUNS of consecutive B rows (cfg2: 6.0 per
//     cluster, at most 14).  Consecutive rows of a row-major block are contiguous in memory, so a run is ONE linear TMA copy
//     (cp.async.bulk, ~11 KB) at the block's own pitch: 7 requests per cluster instead of 33;
//   * the multiply runs on the FP64 FMA pipe over the matrix entries themselves: a consumer warp owns a row of the cluster,
//     its lanes own column pairs (lane + 32 j), and per entry of the row it issues one broadcast load of (staged row,
//     value) and NP conflict-free LDS.128 of the staged B row (any pitch is conflict-free when consecutive lanes read
//     consecutive 16-byte words, so no padded staging pitch is needed -- which is what lets a run land in shared memory
//     with a single copy).
// What bounds it (clock64 sections, profiles/r02_spmm_runs.md): the shared-memory / LSU pipe at 128 B per clock per SM.
// Per cluster at cfg2, m = 266: 97 entries x 2.1 KB = 207 KB of B loads + 67 KB written by TMA + 30 KB of result stores
// = ~2400 pipe cycles against 2790 cycles per cluster at the HBM copy rate -- the two limits nearly coincide.
//
// Structure as the ring kernel: one resident CTA per SM walks clusters blockIdx.x + i * gridDim.x; a producer warp keeps a
// ring of 2-8 slots (record + staged rows) full, completion on FULL mbarriers, run tables fetched three clusters ahead;
// 16 consumer warps; a slot is released with one arrive per warp on EMPTY once the warp's sums are in registers.
// Summation order: the entries of a row in ascending column order, fused multiply-add from 0.0 -- the order of a
// sequential CSR product; results are bitwise reproducible and independent of the ring depth / group count.
// Measured and dropped (profiles/r02_spmm_runs.md): merged column lists of row pairs (fewer loads, longer chains: slower),
// two or three accumulator sets per column pair (slower), L2 prefetch of records further ahead (no effect), two consumer
// groups of 8 warps on alternate clusters (+-3 %).
#include "../../include/hfb200.h"
#include "hfb_common.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace hfb {

constexpr int RUNS_CONSUMERS = 16;
constexpr int RUNS_THREADS = (RUNS_CONSUMERS + 1) * 32;
constexpr int RUNS_MAX_SLOTS = 8;
constexpr int RUNS_MAX_RUNS = 32;     // one producer lane per run
constexpr int RUNS_MAX_NP = 6;        // column pairs per lane: m <= 64 * 6 = 384

// Per-cluster record, written on the host by hfb_csr_pack_clusters_runs:
//     int32 header[4] = {nrow, nrun, nbrow (= staged B rows = distinct columns), nent}
//     int32 outrow[max_rows]        global row of each cluster row (rows sorted ascending)
//     int32 goff[max_rows + 1]      entry offsets of the cluster rows
//     {int32 start; int32 len | off << 16} runs[max_runs]     first B row, length, first staged row of the run
//     {int32 staged_row; int32 0; double v} entries[nent]     the entries of each row, ascending column
struct RunsLayout {
    int off_outrow, off_goff, off_runs, off_ent, stride;
};
static inline int rup(int x, int a) { return (x + a - 1) / a * a; }
static RunsLayout runs_layout(int max_rows, int max_runs, int max_entries) {
    RunsLayout L;
    L.off_outrow = 16;
    L.off_goff = L.off_outrow + 4 * rup(max_rows, 4);
    L.off_runs = L.off_goff + 4 * rup(max_rows + 1, 4);
    L.off_ent = rup(L.off_runs + 8 * max_runs, 16);
    L.stride = rup(L.off_ent + 16 * max_entries, 128);
    return L;
}

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int NP>
__global__ void __launch_bounds__(RUNS_THREADS, 1)
    csr_spmm_runs_kernel(int m, int nclusters, int nslots, int slot_bytes, RunsLayout L,
                         const unsigned char* __restrict__ blobs, const double* __restrict__ B, long long ldb,
                         double* __restrict__ C, long long ldc, unsigned long long* prof) {
    extern __shared__ __align__(128) unsigned char smem_runs[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_runs);
    uint64_t* empty = full + RUNS_MAX_SLOTS;
    unsigned char* slots = smem_runs + 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < nslots; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], RUNS_CONSUMERS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const int width = m + (m & 1);
    const int npairs = width >> 1;

    if (warp == RUNS_CONSUMERS) {
        // ------------------------------------------------------------------ producer
        int c = blockIdx.x;
        // The header and the run table of a cluster are requested three clusters ahead by two INDEPENDENT loads (the table
        // slot of a lane past nrun is zero-filled padding): a load predicated on the header would stall this warp for a
        // DRAM round trip per cluster (measured: a width-independent 2270 cycles per cluster).
        const int run_lane = min(lane, (L.off_ent - L.off_runs) / 8 - 1);
        auto fetch = [&](int cl, int2& run, int& nrun, int& nbrow) {
            run = make_int2(0, 0);
            nrun = nbrow = 0;
            if (cl < nclusters) {
                const unsigned char* bl = blobs + (size_t)cl * L.stride;
                const int4 h = __ldg(reinterpret_cast<const int4*>(bl));
                run = __ldg(reinterpret_cast<const int2*>(bl + L.off_runs) + run_lane);
                nrun = h.y;
                nbrow = h.z;
            }
        };
        int2 runA, runB, runC;
        int nrA, nrB, nrC, nbA, nbB, nbC;
        fetch(c, runA, nrA, nbA);
        fetch(c + (int)gridDim.x, runB, nrB, nbB);
        fetch(c + 2 * (int)gridDim.x, runC, nrC, nbC);
        int it = 0;
        auto one_cluster = [&](int2& runX, int& nrX, int& nbX) {
            const int s = it % nslots;
            const uint32_t ph = (uint32_t)(it / nslots) & 1u;
            const int2 run = runX;
            const int nrun = nrX, nbrow = nbX;
            const unsigned char* blob = blobs + (size_t)c * L.stride;
            fetch(c + 3 * (int)gridDim.x, runX, nrX, nbX);
            const long long tp0 = prof ? clock64() : 0;
            mbar_wait(&empty[s], ph ^ 1u);
            if (prof && lane == 0) atomicAdd(prof + 0, (unsigned long long)(clock64() - tp0));
            unsigned char* st = slots + (size_t)s * slot_bytes;
            double* sB = reinterpret_cast<double*>(st + L.stride);
            if (lane == 0) {
                // a run of len rows is copied as (len - 1) full pitches + one row width: never past the last row's columns
                const uint32_t bytes = (uint32_t)L.stride + 8u * (uint32_t)((long long)(nbrow - nrun) * ldb + (long long)nrun * width);
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_copy_g2s(st, blob, (uint32_t)L.stride, &full[s]);
            }
            __syncwarp();
            if (lane < nrun) {
                const int len = run.y & 0xffff, off = (int)((unsigned)run.y >> 16);
                bulk_copy_g2s(sB + (long long)off * ldb, B + (long long)run.x * ldb,
                              8u * (uint32_t)((long long)(len - 1) * ldb + width), &full[s]);
            }
        };
        while (true) {
            if (c >= nclusters) break;
            one_cluster(runA, nrA, nbA);
            c += gridDim.x;
            ++it;
            if (c >= nclusters) break;
            one_cluster(runB, nrB, nbB);
            c += gridDim.x;
            ++it;
            if (c >= nclusters) break;
            one_cluster(runC, nrC, nbC);
            c += gridDim.x;
            ++it;
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers: warp w owns cluster row w
    bool live[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) live[j] = lane + 32 * j < npairs;
    int it = 0;
    for (long long c = blockIdx.x; c < nclusters; c += gridDim.x, ++it) {
        const int s = it % nslots;
        const uint32_t ph = (uint32_t)(it / nslots) & 1u;
        const long long tc0 = prof ? clock64() : 0;
        mbar_wait(&full[s], ph);
        const long long tc1 = prof ? clock64() : 0;
        const unsigned char* st = slots + (size_t)s * slot_bytes;
        const int nrow = reinterpret_cast<const int*>(st)[0];
        if (warp >= nrow) {                                 // a short cluster: nothing for this warp, release right away
            if (lane == 0) mbar_arrive(&empty[s]);
            continue;
        }
        const int* goff = reinterpret_cast<const int*>(st + L.off_goff);
        const int e0 = goff[warp], e1 = goff[warp + 1];
        const int orow = reinterpret_cast<const int*>(st + L.off_outrow)[warp];
        double2 acc[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) acc[j] = make_double2(0.0, 0.0);
        const double* sB = reinterpret_cast<const double*>(st + L.stride) + 2 * lane;
        const int4* ent = reinterpret_cast<const int4*>(st + L.off_ent) + e0;
#pragma unroll 4
        for (int e = e0; e < e1; ++e, ++ent) {
            const int4 sv = *ent;                           // one broadcast load: {staged row, 0, value}
            const double v = __hiloint2double(sv.w, sv.z);
            const double2* bp = reinterpret_cast<const double2*>(sB + (long long)sv.x * ldb);
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                if (live[j]) {
                    const double2 b = bp[32 * j];
                    acc[j].x = fma(v, b.x, acc[j].x);
                    acc[j].y = fma(v, b.y, acc[j].y);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);              // everything this warp needs from the slot is in registers now
        double* cp = C + (long long)orow * ldc;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int col = 2 * (lane + 32 * j);
            if (!live[j]) continue;
            if (col + 1 < m) {
                *reinterpret_cast<double2*>(cp + col) = acc[j];
            } else if (col < m) {
                cp[col] = acc[j].x;
            }
        }
        if (prof && lane == 0 && warp == 0) {
            atomicAdd(prof + 1, (unsigned long long)(tc1 - tc0));            // consumer warp 0: waiting for data
            atomicAdd(prof + 2, (unsigned long long)(clock64() - tc1));      // header + entries + stores
            atomicAdd(prof + 3, 1ull);
        }
    }
}

static int runs_slots(const RunsLayout& L, int max_brow, int64_t ldb, int& slot_bytes) {
    const long long sb = (long long)L.stride + 8LL * max_brow * ldb;
    if (sb > 227 * 1024) return 0;
    slot_bytes = rup((int)sb, 128);
    int nslots = (227 * 1024 - 128) / slot_bytes;
    const int max_slots = getenv("HFB_RUNS_SLOTS") ? atoi(getenv("HFB_RUNS_SLOTS")) : RUNS_MAX_SLOTS;   // tuning aid
    if (nslots > max_slots) nslots = max_slots;
    if (nslots > RUNS_MAX_SLOTS) nslots = RUNS_MAX_SLOTS;
    return nslots < 2 ? 0 : nslots;
}

template <int NP>
static int launch_runs(int64_t nclusters, int m, const RunsLayout& L, int nslots, int slot_bytes, const void* blobs,
                       const double* B, int64_t ldb, double* C, int64_t ldc, cudaStream_t stream) {
    const size_t smem = 128 + (size_t)nslots * slot_bytes;
    static size_t configured[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (smem > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(csr_spmm_runs_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = smem;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long gx = sms;
    if (gx > nclusters) gx = nclusters;
    static unsigned long long* prof_dev = nullptr;
    unsigned long long* prof_buf = nullptr;
    if (getenv("HFB_RUNS_PROF")) {   // debug aid: cycles the producer waits for a slot / a consumer warp waits for data / works
        if (!prof_dev) cudaMalloc(&prof_dev, 4 * sizeof(unsigned long long));
        cudaMemsetAsync(prof_dev, 0, 4 * sizeof(unsigned long long), stream);
        prof_buf = prof_dev;
    }
    csr_spmm_runs_kernel<NP><<<(unsigned)gx, RUNS_THREADS, smem, stream>>>(
        m, (int)nclusters, nslots, slot_bytes, L, static_cast<const unsigned char*>(blobs), B, ldb, C, ldc, prof_buf);
    ++g_launch_count;
    if (prof_buf) {
        unsigned long long h[4];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
        const double ns = (double)(h[3] ? h[3] : 1);
        fprintf(stderr, "[runs prof] m=%d slots=%d: producer wait-empty %.0f cyc/cluster, consumer wait-full %.0f, consumer work %.0f "
                "(%llu samples)\n", m, nslots, (double)h[0] / (double)nclusters, (double)h[1] / ns, (double)h[2] / ns, h[3]);
    }
    return (int)cudaGetLastError();
}

static int launch_runs_np(int np, int64_t nclusters, int m, const RunsLayout& L, int nslots, int slot_bytes, const void* blobs,
                          const double* B, int64_t ldb, double* C, int64_t ldc, cudaStream_t stream) {
    switch (np) {
        case 1: return launch_runs<1>(nclusters, m, L, nslots, slot_bytes, blobs, B, ldb, C, ldc, stream);
        case 2: return launch_runs<2>(nclusters, m, L, nslots, slot_bytes, blobs, B, ldb, C, ldc, stream);
        case 3: return launch_runs<3>(nclusters, m, L, nslots, slot_bytes, blobs, B, ldb, C, ldc, stream);
        case 4: return launch_runs<4>(nclusters, m, L, nslots, slot_bytes, blobs, B, ldb, C, ldc, stream);
        case 5: return launch_runs<5>(nclusters, m, L, nslots, slot_bytes, blobs, B, ldb, C, ldc, stream);
        case 6: return launch_runs<6>(nclusters, m, L, nslots, slot_bytes, blobs, B, ldb, C, ldc, stream);
    }
    return HFB_E_UNSUPPORTED;
}

// Shared by the measure and the pack pass: the sorted rows, the sorted distinct columns and the runs of one cluster.
struct RunsScratch {
    std::vector<int32_t> rows, cols;
    std::vector<int32_t> stamp, rank;
};

static int runs_cluster_columns(int64_t n, const int32_t* rowptr, const int32_t* colind, const int32_t* order, int32_t s0,
                                int32_t s1, int64_t c, RunsScratch& S) {
    S.rows.assign(order + s0, order + s1);
    std::sort(S.rows.begin(), S.rows.end());
    S.cols.clear();
    for (int32_t row : S.rows) {
        if (row < 0 || row >= n) return HFB_E_BADARG;
        for (int32_t j = rowptr[row]; j < rowptr[row + 1]; ++j) {
            const int32_t col = colind[j];
            if (col < 0 || col >= n) return HFB_E_BADARG;
            if (S.stamp[col] != (int32_t)c) {
                S.stamp[col] = (int32_t)c;
                S.cols.push_back(col);
            }
        }
    }
    std::sort(S.cols.begin(), S.cols.end());
    for (size_t i = 0; i < S.cols.size(); ++i) S.rank[S.cols[i]] = (int32_t)i;   // staged row of a column = its rank
    return 0;
}

}  // namespace hfb

using namespace hfb;

/* HOST: caps of the run records of a cluster plan.  caps_out[4] = {max_rows, max_runs, max_brow, max_entries}. */
extern "C" int hfb_csr_runs_measure(int64_t n, const int32_t* rowptr, const int32_t* colind, const int32_t* order,
                                    const int32_t* cluster_ptr, int64_t nclusters, int32_t* caps_out) {
    if (n <= 0 || !rowptr || !colind || !order || !cluster_ptr || nclusters <= 0 || !caps_out) return HFB_E_BADARG;
    RunsScratch S;
    S.stamp.assign((size_t)n, -1);
    S.rank.assign((size_t)n, 0);
    int32_t max_rows = 0, max_runs = 0, max_brow = 0, max_ent = 0;
    for (int64_t c = 0; c < nclusters; ++c) {
        const int32_t s0 = cluster_ptr[c], s1 = cluster_ptr[c + 1];
        if (s1 <= s0) return HFB_E_BADARG;
        int rc = runs_cluster_columns(n, rowptr, colind, order, s0, s1, c, S);
        if (rc) return rc;
        int32_t nrun = 0;
        for (size_t i = 0; i < S.cols.size(); ++i)
            if (i == 0 || S.cols[i] != S.cols[i - 1] + 1) ++nrun;
        int32_t nent = 0;
        for (int32_t row : S.rows) nent += rowptr[row + 1] - rowptr[row];
        max_rows = std::max(max_rows, s1 - s0);
        max_runs = std::max(max_runs, nrun);
        max_brow = std::max(max_brow, (int32_t)S.cols.size());
        max_ent = std::max(max_ent, nent);
    }
    caps_out[0] = max_rows;
    caps_out[1] = max_runs;
    caps_out[2] = max_brow;
    caps_out[3] = max_ent;
    return 0;
}

extern "C" int64_t hfb_csr_runs_blob_stride(int32_t max_rows, int32_t max_runs, int32_t max_entries) {
    if (max_rows <= 0 || max_rows > RUNS_CONSUMERS || max_runs <= 0 || max_runs > RUNS_MAX_RUNS || max_entries <= 0)
        return HFB_E_UNSUPPORTED;
    return runs_layout(max_rows, max_runs, max_entries).stride;
}

/* HOST: packs the clusters of (order, cluster_ptr) into run records (layout above). */
extern "C" int hfb_csr_pack_clusters_runs(int64_t n, const int32_t* rowptr, const int32_t* colind, const double* val,
                                          const int32_t* order, const int32_t* cluster_ptr, int64_t nclusters, int32_t max_rows,
                                          int32_t max_runs, int32_t max_entries, void* blobs_out) {
    if (n <= 0 || !rowptr || !colind || !val || !order || !cluster_ptr || nclusters <= 0 || !blobs_out) return HFB_E_BADARG;
    if (hfb_csr_runs_blob_stride(max_rows, max_runs, max_entries) < 0) return HFB_E_UNSUPPORTED;
    const RunsLayout L = runs_layout(max_rows, max_runs, max_entries);
    RunsScratch S;
    S.stamp.assign((size_t)n, -1);
    S.rank.assign((size_t)n, 0);
    std::vector<std::pair<int32_t, double>> rowent;
    unsigned char* out = static_cast<unsigned char*>(blobs_out);
    for (int64_t c = 0; c < nclusters; ++c) {
        unsigned char* blob = out + (size_t)c * L.stride;
        memset(blob, 0, (size_t)L.stride);
        const int32_t s0 = cluster_ptr[c], s1 = cluster_ptr[c + 1];
        const int32_t nrow = s1 - s0;
        if (nrow <= 0 || nrow > max_rows) return HFB_E_BADARG;
        int rc = runs_cluster_columns(n, rowptr, colind, order, s0, s1, c, S);
        if (rc) return rc;
        if (S.cols.size() > 0xffffu) return HFB_E_UNSUPPORTED;
        int32_t* hdr = reinterpret_cast<int32_t*>(blob);
        int32_t* outrow = reinterpret_cast<int32_t*>(blob + L.off_outrow);
        int32_t* goff = reinterpret_cast<int32_t*>(blob + L.off_goff);
        int32_t* runs = reinterpret_cast<int32_t*>(blob + L.off_runs);
        int32_t nrun = 0;
        for (size_t i = 0; i < S.cols.size();) {
            size_t j = i + 1;
            while (j < S.cols.size() && S.cols[j] == S.cols[j - 1] + 1) ++j;
            if (nrun >= max_runs) return HFB_E_UNSUPPORTED;
            runs[2 * nrun] = S.cols[i];
            runs[2 * nrun + 1] = (int32_t)((uint32_t)(j - i) | ((uint32_t)i << 16));
            ++nrun;
            i = j;
        }
        int32_t nent = 0;
        for (int32_t r = 0; r < nrow; ++r) {
            const int32_t row = S.rows[r];
            outrow[r] = row;
            goff[r] = nent;
            rowent.clear();
            for (int32_t j = rowptr[row]; j < rowptr[row + 1]; ++j) rowent.emplace_back(colind[j], val[j]);
            std::stable_sort(rowent.begin(), rowent.end(),
                             [](const std::pair<int32_t, double>& a, const std::pair<int32_t, double>& b) { return a.first < b.first; });
            for (size_t k = 0; k < rowent.size(); ++k) {
                if (k > 0 && rowent[k].first == rowent[k - 1].first) {      // duplicate entries sum, as in CSR
                    double prev;
                    unsigned char* e = blob + L.off_ent + 16 * (size_t)(nent - 1);
                    memcpy(&prev, e + 8, 8);
                    prev += rowent[k].second;
                    memcpy(e + 8, &prev, 8);
                    continue;
                }
                if (nent >= max_entries) return HFB_E_UNSUPPORTED;
                unsigned char* e = blob + L.off_ent + 16 * (size_t)nent;
                const int32_t slot = S.rank[rowent[k].first];
                memcpy(e, &slot, 4);
                memcpy(e + 8, &rowent[k].second, 8);
                ++nent;
            }
        }
        for (int32_t r = nrow; r <= max_rows; ++r) goff[r] = nent;
        hdr[0] = nrow;
        hdr[1] = nrun;
        hdr[2] = (int32_t)S.cols.size();
        hdr[3] = nent;
    }
    return 0;
}

/* Number of ring slots hfb_csr_spmm_runs would use for this shape (0: the shape is unsupported, use another kernel). */
extern "C" int32_t hfb_csr_spmm_runs_slots(int64_t m, int64_t ldb, int32_t max_rows, int32_t max_runs, int32_t max_brow,
                                           int32_t max_entries) {
    if (m <= 0 || m > 64 * RUNS_MAX_NP || ldb < m + (m & 1) || max_brow <= 0) return 0;
    if (hfb_csr_runs_blob_stride(max_rows, max_runs, max_entries) < 0) return 0;
    const RunsLayout L = runs_layout(max_rows, max_runs, max_entries);
    int slot_bytes = 0;
    return runs_slots(L, max_brow, ldb, slot_bytes);
}

extern "C" int hfb_csr_spmm_runs(int64_t nclusters, int64_t m, const void* blobs, int32_t max_rows, int32_t max_runs,
                                 int32_t max_brow, int32_t max_entries, const double* B, int64_t ldb, double* C, int64_t ldc,
                                 void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nclusters <= 0 || m <= 0 || !blobs || !B || !C || B == C || max_brow <= 0) return HFB_E_BADARG;
    if (ldb < m + (m & 1) || ldc < m) return HFB_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(B) & 15) || (reinterpret_cast<uintptr_t>(C) & 15) || (reinterpret_cast<uintptr_t>(blobs) & 15) ||
        (ldb & 1) || (ldc & 1))
        return HFB_E_ALIGN;
    if (nclusters > 0x7fffffffLL || m > 64 * RUNS_MAX_NP) return HFB_E_UNSUPPORTED;
    if (hfb_csr_runs_blob_stride(max_rows, max_runs, max_entries) < 0) return HFB_E_UNSUPPORTED;
    const RunsLayout L = runs_layout(max_rows, max_runs, max_entries);
    int slot_bytes = 0;
    const int nslots = runs_slots(L, max_brow, ldb, slot_bytes);
    if (nslots == 0) return HFB_E_UNSUPPORTED;
    const int np = (int)(((m + 1) / 2 + 31) / 32);      // column pairs per lane
    return launch_runs_np(np, nclusters, (int)m, L, nslots, slot_bytes, blobs, B, ldb, C, ldc, stream);
}
