// Sketch exchange over NVLink peer memory (see include/hfb200.h, "peer exchange"): the allreduce of the (n x m) lift
//   Y = sum_g X_g^T W_g                (collectiveOperator.py:73-80 -> collective.py:108-111, one MPI message per column)
// as  reduce-scatter by PUSH from the lift GEMM's epilogue (hfb_dgemm_peer, dgemm_dmma.cuh)  ->  fixed-order sum of the
// P slots by the owner of each row block, stored into every rank's result block (all-gather by PUSH).  Everything here is
// plain st.global on peer-mapped addresses (cudaIpc*), ordered by system-scope release/acquire flags; no NCCL on the data path.
#include <cstdio>
#include <cstring>

#include "../../include/hfb200.h"
#include "hfb_common.cuh"

namespace hfb {

constexpr int PEER_MAX = 16;

struct PeerPtrs {
    void* p[PEER_MAX];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* addr, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* addr) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Barrier over the ranks of one NVLink domain.  flags.p[r] is rank r's flag array (PEER_MAX counters, peer-mapped for
// r != me).  Thread t publishes flags[t][me] = epoch (release, system scope: every store this stream issued before --
// the GEMM's peer stores included -- is visible to whoever acquires the flag) and then waits for flags[me][t] >= epoch.
// Epochs only grow, so no flag is ever reset.  A wait longer than timeout_ns traps (a dead peer must not hang the GPU).
// mode: HFB_PEER_SIGNAL | HFB_PEER_WAIT (both = a barrier).
__global__ void peer_barrier_kernel(PeerPtrs flags, int me, int nranks, unsigned long long epoch,
                                    unsigned long long timeout_ns, int mode) {
    const int t = threadIdx.x;
    if (t >= nranks) return;
    if (mode & HFB_PEER_SIGNAL) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(flags.p[t]) + me, epoch);
    }
    if (!(mode & HFB_PEER_WAIT)) return;
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(flags.p[me]) + t;
    const unsigned long long t0 = global_timer_ns();
    unsigned int spins = 0;
    while (ld_acquire_sys(mine) < epoch) {
        if ((++spins & 1023u) == 0 && timeout_ns && global_timer_ns() - t0 > timeout_ns) {
            printf("hfb peer barrier: rank %d timed out waiting for rank %d (epoch %llu)\n", me, t, epoch);
            __trap();
        }
    }
    __threadfence_system();
}

// Owner side of the reduce-scatter fused with the all-gather: s[r][c] = sum_{k < nranks} slots[k][r][c] in FIXED order (bitwise
// reproducible and identical on every rank, because only the owner computes it), stored into the owner's rows of the result
// block of EVERY rank -- local first, then the peers over NVLink, starting with me + 1 so that the ranks do not all write to
// the same GPU at the same time.  A warp writes 512 contiguous bytes per destination.
__global__ void peer_reduce_bcast_kernel(const double* __restrict__ slots, long long slot_stride, int nranks, int me,
                                         long long rows, int cols, long long ld, PeerPtrs ydst, long long ldy) {
    const int pairs = (cols + 1) >> 1;
    const long long total = rows * pairs;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / pairs;
        const int c = 2 * (int)(idx - r * pairs);
        const long long off = r * ld + c;
        const long long yoff = r * ldy + c;
        if (c + 1 < cols) {
            // loads of four slots are issued together (the adds stay in rank order)
            double2 s = make_double2(0.0, 0.0);
            for (int k = 0; k < nranks; k += 4) {
                double2 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    v[j] = (k + j < nranks) ? *reinterpret_cast<const double2*>(slots + (long long)(k + j) * slot_stride + off)
                                            : make_double2(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (k + j < nranks) {
                        s.x = (k + j == 0) ? v[j].x : s.x + v[j].x;
                        s.y = (k + j == 0) ? v[j].y : s.y + v[j].y;
                    }
                }
            }
            for (int i = 0; i < nranks; ++i) {
                int d = me + i;
                if (d >= nranks) d -= nranks;
                *reinterpret_cast<double2*>(reinterpret_cast<double*>(ydst.p[d]) + yoff) = s;
            }
        } else {
            double s = slots[off];
            for (int k = 1; k < nranks; ++k) s += slots[(long long)k * slot_stride + off];
            for (int i = 0; i < nranks; ++i) {
                int d = me + i;
                if (d >= nranks) d -= nranks;
                reinterpret_cast<double*>(ydst.p[d])[yoff] = s;
            }
        }
    }
}

}  // namespace hfb

using namespace hfb;

// ---------------------------------------------------------------------------------------------------------- memory
extern "C" int hfb_peer_alloc(size_t bytes, void** ptr) {
    if (!ptr || bytes == 0) return HFB_E_BADARG;
    *ptr = nullptr;
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*ptr, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    return (int)e;
}

extern "C" int hfb_peer_free(void* ptr) {
    if (!ptr) return HFB_E_BADARG;
    return (int)cudaFree(ptr);
}

extern "C" int hfb_peer_get_handle(void* ptr, unsigned char* handle64) {
    if (!ptr || !handle64) return HFB_E_BADARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == HFB_PEER_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) return (int)e;
    memcpy(handle64, &h, sizeof(h));
    return 0;
}

extern "C" int hfb_peer_open(const unsigned char* handle64, void** ptr) {
    if (!handle64 || !ptr) return HFB_E_BADARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    *ptr = nullptr;
    return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

extern "C" int hfb_peer_close(void* ptr) {
    if (!ptr) return HFB_E_BADARG;
    return (int)cudaIpcCloseMemHandle(ptr);
}

// ---------------------------------------------------------------------------------------------------------- kernels
extern "C" int hfb_peer_barrier(void* const* flag_ptrs, int me, int nranks, uint64_t epoch, double timeout_s, int mode,
                                void* stream_) {
    if (!flag_ptrs || nranks < 1 || nranks > PEER_MAX || me < 0 || me >= nranks || epoch == 0) return HFB_E_BADARG;
    if (mode < 1 || mode > (HFB_PEER_SIGNAL | HFB_PEER_WAIT)) return HFB_E_BADARG;
    PeerPtrs f;
    memset(&f, 0, sizeof(f));
    for (int r = 0; r < nranks; ++r) {
        if (!flag_ptrs[r] || (reinterpret_cast<uintptr_t>(flag_ptrs[r]) & 7)) return HFB_E_BADARG;
        f.p[r] = flag_ptrs[r];
    }
    const unsigned long long tns = timeout_s > 0 ? (unsigned long long)(timeout_s * 1e9) : 0ULL;
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>(f, me, nranks, (unsigned long long)epoch, tns, mode);
    ++g_launch_count;
    return (int)cudaGetLastError();
}

extern "C" int hfb_peer_reduce_bcast(const double* slots, int64_t slot_stride, int nranks, int me, int64_t rows, int64_t cols,
                                     int64_t ld, double* const* y_ptrs, int64_t ldy, void* stream_) {
    if (!slots || !y_ptrs || nranks < 1 || nranks > PEER_MAX || me < 0 || me >= nranks || rows < 0 || cols <= 0 || ld < cols ||
        ldy < cols || slot_stride < rows * ld || cols > 0x7ffffffeLL)
        return HFB_E_BADARG;
    if ((ld & 1) || (ldy & 1) || (slot_stride & 1) || (reinterpret_cast<uintptr_t>(slots) & 15)) return HFB_E_ALIGN;
    PeerPtrs y;
    memset(&y, 0, sizeof(y));
    for (int r = 0; r < nranks; ++r) {
        if (!y_ptrs[r]) return HFB_E_BADARG;
        if (reinterpret_cast<uintptr_t>(y_ptrs[r]) & 15) return HFB_E_ALIGN;
        y.p[r] = y_ptrs[r];
    }
    if (rows == 0) return 0;
    const long long total = rows * ((cols + 1) / 2);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    peer_reduce_bcast_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(slots, slot_stride, nranks, me, rows, (int)cols,
                                                                                  ld, y, ldy);
    ++g_launch_count;
    return (int)cudaGetLastError();
}
