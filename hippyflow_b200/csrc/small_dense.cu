// Small dense factorisations that used to be host LAPACK round trips (see include/hfb200.h):
//
//   hfb_chol_inverse        one CTA: column-scaled (shifted) Cholesky factor R of an (m x m) Gram matrix and the
//                           triangular inverse S = D^-1 R^-1 that Cholesky-QR multiplies the sketch with, m <= 1024;
//   hfb_jacobi_svd_batched  one CTA per sample: one-sided (Hestenes) Jacobi SVD of a small (rows x cols) block held in
//                           shared memory -- orthonormalisation of the per-sample sketches J_i Omega and the eigen
//                           decomposition of the per-sample (l x l) Gram matrices of the batched randomized SVD.
//
// Both are latency-bound O(m^3) kernels on a few hundred KB; the point is that the GPU never waits for the host.
#include "../../include/hfb200.h"
#include "hfb_common.cuh"

#include <cfloat>

namespace hfb {

// ------------------------------------------------------------------------------------------------ block reductions
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// max over the CTA; every thread gets the result.  red: >= 33 doubles of shared memory.
__device__ double block_max(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double x = lane < nw ? red[lane] : -DBL_MAX;
        x = warp_max(x);
        if (lane == 0) red[32] = x;
    }
    __syncthreads();
    return red[32];
}
__device__ double block_min(double v, double* red) { return -block_max(-v, red); }

// ------------------------------------------------------------------------------------------------ Cholesky + inverse
constexpr int CI_THREADS = 1024;
constexpr int CI_NB = 8;  // panel height
constexpr int CI_U = 4;   // independent global read-modify-writes in flight per thread

// Upper Cholesky factor of W0 + shift*I (upper triangle read), in place in W, RIGHT-looking by block rows of CI_NB rows.
// Per panel: (1) the block row, already carrying every earlier update, is loaded into shared memory; (2) one thread factors
// its (CI_NB x CI_NB) diagonal block; (3) every thread forward-substitutes its own columns in registers (no barrier inside
// the panel); (4) ALL threads subtract the panel's rank-CI_NB contribution from the trailing matrix -- independent
// read-modify-writes issued CI_U at a time, so global-memory latency is covered by memory-level parallelism.
// Returns 1 (to every thread) when a pivot is not positive.
// (W0, W and S are written earlier in the same kernel: plain pointers, no __restrict__/const, so that no load is routed
// through the non-coherent read-only path; the price -- the compiler keeps every load behind earlier stores -- is why the
// loads of a batch are issued explicitly before its stores.)
// section profiling (debug aid, hfb_chol_inverse_profile): thread 0 accumulates clock64() deltas per section
#define CI_TICK(i)                                  \
    if (prof && threadIdx.x == 0) {                 \
        const long long now_ = clock64();           \
        prof[i] += now_ - tlast;                    \
        tlast = now_;                               \
    }

__device__ int chol_upper_blocked(int m, double* W0, double* W, long long ldw, double shift, double* P, int PW, double* rdiag,
                                  double* Dblk, int* s_fail, long long* prof, long long& tlast) {
    const int tid = threadIdx.x, T = blockDim.x;
    for (int base = tid; base < m * m; base += T * CI_U) {
        double v[CI_U];
#pragma unroll
        for (int u = 0; u < CI_U; ++u) {
            const int idx = base + u * T;
            const int i = idx / m, c = idx - i * m;
            v[u] = (idx < m * m && c >= i) ? W0[(long long)i * ldw + c] + (c == i ? shift : 0.0) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < CI_U; ++u) {
            const int idx = base + u * T;
            const int i = idx / m, c = idx - i * m;
            if (idx < m * m && c >= i) W[(long long)i * ldw + c] = v[u];
        }
    }
    if (tid == 0) *s_fail = 0;
    __syncthreads();
    CI_TICK(1)
    for (int i0 = 0; i0 < m; i0 += CI_NB) {
        const int nb = min(CI_NB, m - i0), i1 = i0 + nb, w = m - i0;
        for (int idx = tid; idx < CI_NB * (w + 8); idx += T) {
            const int r = idx / (w + 8), cc = idx - r * (w + 8);
            P[r * PW + cc] = (r < nb && cc >= r && cc < w) ? W[(long long)(i0 + r) * ldw + i0 + cc] : 0.0;
        }
        __syncthreads();
        CI_TICK(2)
        if (tid < 32) {  // diagonal block R11^T R11 = A11 (8 x 8) by warp 0: lane c owns column c in registers
            const int c = tid;
            double d[CI_NB];
#pragma unroll
            for (int r = 0; r < CI_NB; ++r) d[r] = (c < nb && r <= c) ? P[r * PW + c] : (r == c ? 1.0 : 0.0);
            int bad = 0;
#pragma unroll
            for (int j = 0; j < CI_NB; ++j) {
                const double piv = __shfl_sync(0xffffffffu, d[j], j);
                if (j < nb && (!(piv > 0.0) || !(piv < DBL_MAX))) bad = 1;
                const double rjj = sqrt(bad ? 1.0 : piv);
                if (c == j) d[j] = rjj;
                else if (c > j) d[j] = d[j] / rjj;
#pragma unroll
                for (int r = j + 1; r < CI_NB; ++r) {
                    const double vr = __shfl_sync(0xffffffffu, d[j], r);      // D[j][r]
                    if (c >= r) d[r] = fma(-vr, d[j], d[r]);
                }
            }
            if (c < CI_NB) {
#pragma unroll
                for (int r = 0; r < CI_NB; ++r) Dblk[r * CI_NB + c] = (c < nb && r <= c) ? d[r] : 0.0;
                if (c < nb) rdiag[i0 + c] = d[c];
            }
            if (bad && c == 0) *s_fail = 1;
        }
        __syncthreads();
        CI_TICK(3)
        if (*s_fail) return 1;
        for (int cc = tid; cc < w; cc += T) {
            if (cc < nb) {
#pragma unroll
                for (int r = 0; r < CI_NB; ++r) P[r * PW + cc] = (r <= cc) ? Dblk[r * CI_NB + cc] : 0.0;
            } else {  // R12[:, cc] = R11^-T A12[:, cc]
                double y[CI_NB];
#pragma unroll
                for (int r = 0; r < CI_NB; ++r) {
                    double v = P[r * PW + cc];
#pragma unroll
                    for (int q = 0; q < r; ++q) v = fma(-Dblk[q * CI_NB + r], y[q], v);
                    y[r] = (r < nb) ? v / Dblk[r * CI_NB + r] : 0.0;
                }
#pragma unroll
                for (int r = 0; r < CI_NB; ++r) P[r * PW + cc] = y[r];
            }
        }
        __syncthreads();
        CI_TICK(4)
        for (int idx = tid; idx < nb * w; idx += T) {
            const int r = idx / w, cc = idx - r * w;
            if (cc >= r) W[(long long)(i0 + r) * ldw + i0 + cc] = P[r * PW + cc];
        }
        // trailing update W[i][c] -= sum_r P[r][i - i0] P[r][c - i0] (i >= i1, c >= i) as a rank-CI_NB SYRK on the FP64 tensor
        // pipe: one warp per 8 x 8 tile, A[g][t] = P[k][a0 + g], B[t][g] = P[k][b0 + g] straight from the shared-memory panel
        // (2 LDS per DMMA instead of 16 LDS per element), C read-modify-written in place as 16-byte pairs.
        const int wt = m - i1, nt8 = (wt + 7) >> 3;
        const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5, g = lane >> 2, t = lane & 3;
        for (int tile0 = warp; tile0 < nt8 * nt8; tile0 += nwarps * CI_U) {  // CI_U tiles per trip: their loads overlap
            double c0[CI_U], c1[CI_U];
            double* cp[CI_U];
            int ca[CI_U], cb[CI_U];
            bool ok0[CI_U], ok1[CI_U];
#pragma unroll
            for (int u = 0; u < CI_U; ++u) {
                const int tile = tile0 + u * nwarps;
                const int ti = tile / nt8, tj = tile - ti * nt8;
                const int i = i1 + 8 * ti + g, c = i1 + 8 * tj + 2 * t;
                const bool live = tile < nt8 * nt8 && tj >= ti;
                ok0[u] = live && i < m && c < m && c >= i;
                ok1[u] = live && i < m && c + 1 < m && c + 1 >= i;
                cp[u] = W + (long long)i * ldw + c;
                ca[u] = live ? nb + 8 * ti + g : 0;
                cb[u] = live ? nb + 8 * tj + g : 0;
                c0[u] = ok0[u] ? cp[u][0] : 0.0;
                c1[u] = ok1[u] ? cp[u][1] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < CI_U; ++u) {
#pragma unroll
                for (int ks = 0; ks < CI_NB / 4; ++ks)
                    dmma884(c0[u], c1[u], -P[(4 * ks + t) * PW + ca[u]], P[(4 * ks + t) * PW + cb[u]]);
            }
#pragma unroll
            for (int u = 0; u < CI_U; ++u) {
                if (ok0[u]) cp[u][0] = c0[u];
                if (ok1[u]) cp[u][1] = c1[u];
            }
        }
        __syncthreads();
        CI_TICK(5)
    }
    return 0;
}

// S (upper) = R^-1 for the upper factor in W, block rows from the bottom, EAGER: the partial sums
// sum_{k solved} R[i][k] S[k][c] of every unsolved row are kept in S itself.  Per panel: load the sums of its rows, every
// thread back-substitutes its own columns in registers, then ALL threads add the panel's contribution to the rows above
// (block column of R staged in shared memory, read-modify-writes issued CI_U at a time).
__device__ void inverse_upper_blocked(int m, double* W, long long ldw, double* S, long long lds, double* P, int PW,
                                      double* Dblk, double* Rcol, long long* prof, long long& tlast) {
    const int tid = threadIdx.x, T = blockDim.x;
    for (int idx = tid; idx < m * m; idx += T) S[(long long)(idx / m) * lds + idx % m] = 0.0;
    __syncthreads();
    CI_TICK(6)
    for (int i0 = ((m - 1) / CI_NB) * CI_NB; i0 >= 0; i0 -= CI_NB) {
        const int nb = min(CI_NB, m - i0), w = m - i0;
        for (int idx = tid; idx < CI_NB * (w + 8); idx += T) {
            const int r = idx / (w + 8), cc = idx - r * (w + 8);
            P[r * PW + cc] = (r < nb && cc >= r && cc < w) ? S[(long long)(i0 + r) * lds + i0 + cc] : 0.0;
        }
        if (tid < CI_NB * CI_NB) {
            const int r = tid / CI_NB, q = tid % CI_NB;
            Dblk[tid] = (r < nb && q < nb && q >= r) ? W[(long long)(i0 + r) * ldw + i0 + q] : (r == q ? 1.0 : 0.0);
        }
        __syncthreads();
        CI_TICK(7)
        for (int cc = tid; cc < w; cc += T) {  // rows of the panel, bottom-up, for column cc: t_jj = (delta - sums) / r_jj
            double t[CI_NB];
#pragma unroll
            for (int jj = CI_NB - 1; jj >= 0; --jj) {
                double sum = P[jj * PW + cc];
#pragma unroll
                for (int q = jj + 1; q < CI_NB; ++q) sum = fma(Dblk[jj * CI_NB + q], t[q], sum);
                t[jj] = (jj < nb && jj <= cc) ? ((cc == jj ? 1.0 : 0.0) - sum) / Dblk[jj * CI_NB + jj] : 0.0;
            }
#pragma unroll
            for (int jj = 0; jj < CI_NB; ++jj) P[jj * PW + cc] = t[jj];
        }
        for (int idx = tid; idx < (i0 + 8) * CI_NB; idx += T) {  // block column R[0:i0][i0:i0+nb] for the eager update (+ zero tail)
            const int i = idx / CI_NB, q = idx - i * CI_NB;
            Rcol[idx] = (i < i0 && q < nb) ? W[(long long)i * ldw + i0 + q] : 0.0;
        }
        __syncthreads();
        CI_TICK(8)
        for (int idx = tid; idx < nb * w; idx += T) {
            const int r = idx / w, cc = idx - r * w;
            if (cc >= r) S[(long long)(i0 + r) * lds + i0 + cc] = P[r * PW + cc];
        }
        // rows above: S[i][c] += sum_q R[i][i0 + q] S[i0 + q][c], i < i0, c >= i0, again as 8 x 8 DMMA tiles:
        // A[g][t] = Rcol[i][k], B[t][g] = P[k][cc]  (P[q][cc] = 0 for cc < q, so no triangle test is needed)
        const int nti = (i0 + 7) >> 3, ntj = (w + 7) >> 3;
        const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5, g = lane >> 2, t = lane & 3;
        for (int tile0 = warp; tile0 < nti * ntj; tile0 += nwarps * CI_U) {
            double c0[CI_U], c1[CI_U];
            double* cp[CI_U];
            int ra[CI_U], cb[CI_U];
            bool ok0[CI_U], ok1[CI_U];
#pragma unroll
            for (int u = 0; u < CI_U; ++u) {
                const int tile = tile0 + u * nwarps;
                const int ti = tile / ntj, tj = tile - ti * ntj;
                const int i = 8 * ti + g, cc = 8 * tj + 2 * t;
                const bool live = tile < nti * ntj;
                ok0[u] = live && i < i0 && cc < w;
                ok1[u] = live && i < i0 && cc + 1 < w;
                cp[u] = S + (long long)i * lds + i0 + cc;
                ra[u] = live ? (8 * ti + g) * CI_NB : 0;
                cb[u] = live ? 8 * tj + g : 0;
                c0[u] = ok0[u] ? cp[u][0] : 0.0;
                c1[u] = ok1[u] ? cp[u][1] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < CI_U; ++u) {
#pragma unroll
                for (int ks = 0; ks < CI_NB / 4; ++ks)
                    dmma884(c0[u], c1[u], Rcol[ra[u] + 4 * ks + t], P[(4 * ks + t) * PW + cb[u]]);
            }
#pragma unroll
            for (int u = 0; u < CI_U; ++u) {
                if (ok0[u]) cp[u][0] = c0[u];
                if (ok1[u]) cp[u][1] = c1[u];
            }
        }
        __syncthreads();
        CI_TICK(9)
    }
}

// stat[0..7] = {fail, shift, cond, attempts, max |Gs - I|, max |d_j - 1| over live columns, dead columns, m}
__global__ void __launch_bounds__(CI_THREADS, 1)
chol_inverse_kernel(int m, const double* __restrict__ G, long long ldg, double* S, long long lds,
                    double* W0, double* W, long long ldw, double* __restrict__ stat,
                    int scale_columns, double cond_max, long long* prof) {
    extern __shared__ double sm[];
    const int PW = (m + 9) & ~1;          // + 8 zero columns: partial DMMA tiles read past the panel width
    double* P = sm;                       // CI_NB x PW
    double* dinv = P + CI_NB * PW;        // m
    double* rdiag = dinv + PW;            // m
    double* Dblk = rdiag + PW;            // CI_NB x CI_NB
    double* red = Dblk + CI_NB * CI_NB;   // 40
    double* Rcol = red + 40;              // (m + 8) x CI_NB
    const int tid = threadIdx.x, T = blockDim.x;
    const double eps = DBL_EPSILON;
    __shared__ int s_fail;
    long long tlast = prof ? clock64() : 0;

    // column scaling d_j = sqrt(G_jj); a zero column stays zero (dinv = 0, unit diagonal), as in hIPPYlib's MGS
    double dmaxdev = 0.0, ndead = 0.0;
    for (int j = tid; j < m; j += T) {
        const double g = G[(long long)j * ldg + j];
        const double d = sqrt(fmax(g, 0.0));
        const bool dead = !(d > 0.0);
        if (scale_columns) {
            dinv[j] = dead ? 0.0 : 1.0 / d;
            if (!dead) dmaxdev = fmax(dmaxdev, fabs(d - 1.0));
        } else {
            dinv[j] = 1.0;
        }
        if (dead) ndead += 1.0;
    }
    __syncthreads();
    double dev = 0.0;
    for (int idx = tid; idx < m * m; idx += T) {
        const int i = idx / m, c = idx - i * m;
        if (c < i) continue;
        double gs = 0.5 * (G[(long long)i * ldg + c] + G[(long long)c * ldg + i]) * dinv[i] * dinv[c];
        if (i == c && scale_columns && dinv[i] == 0.0) gs = 1.0;
        W0[(long long)i * ldw + c] = gs;
        dev = fmax(dev, fabs(gs - (i == c ? 1.0 : 0.0)));
    }
    dev = block_max(dev, red);
    dmaxdev = block_max(dmaxdev, red);
    double nd = ndead;
    nd = warp_sum(nd);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = nd;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < (T + 31) / 32; ++i) s += red[i];
        red[33] = s;
    }
    __syncthreads();
    ndead = red[33];
    __syncthreads();

    CI_TICK(0)
    double shift = 0.0, cond = 0.0;
    int attempts = 0, ok = 0;
    for (int attempt = 0; attempt < 8; ++attempt) {
        ++attempts;
        const int fail = chol_upper_blocked(m, W0, W, ldw, shift, P, PW, rdiag, Dblk, &s_fail, prof, tlast);
        __syncthreads();
        if (!fail) {
            double mx = 0.0, mn = DBL_MAX;
            for (int j = tid; j < m; j += T) {
                mx = fmax(mx, rdiag[j]);
                mn = fmin(mn, rdiag[j]);
            }
            mx = block_max(mx, red);
            mn = block_min(mn, red);
            cond = (mx / mn) * (mx / mn);
            if (cond < DBL_MAX && (cond < cond_max || shift > 0.0)) {
                ok = 1;
                break;
            }
        }
        shift = (shift == 0.0) ? 100.0 * m * eps : shift * 100.0;
    }
    if (ok) {
        inverse_upper_blocked(m, W, ldw, S, lds, P, PW, Dblk, Rcol, prof, tlast);
        // undo the column scaling (S = D^-1 R^-1) and clear the strictly lower part
        for (int base = tid; base < m * m; base += T * CI_U) {
            double v[CI_U];
#pragma unroll
            for (int u = 0; u < CI_U; ++u) {
                const int idx = base + u * T;
                const int i = idx / m, c = idx - i * m;
                v[u] = (idx < m * m && c >= i) ? S[(long long)i * lds + c] * dinv[i] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < CI_U; ++u) {
                const int idx = base + u * T;
                if (idx < m * m) S[(long long)(idx / m) * lds + (idx % m)] = v[u];
            }
        }
    }
    __syncthreads();
    CI_TICK(10)
    if (tid == 0) {
        stat[0] = ok ? 0.0 : 1.0;
        stat[1] = shift;
        stat[2] = cond;
        stat[3] = (double)attempts;
        stat[4] = dev;
        stat[5] = dmaxdev;
        stat[6] = ndead;
        stat[7] = (double)m;
    }
}

// ------------------------------------------------------------------------------------------------ batched Jacobi SVD
constexpr int JSVD_THREADS = 512;

// One CTA per sample.  The (rows x cols) block is held column-major in shared memory (odd pitch: conflict-free for a
// warp walking one column and for the transposing load/store).  Sweeps of plane rotations in round-robin order (cols/2
// independent pairs per round, one warp per pair) make the columns mutually orthogonal: A V = U diag(sigma).
// On exit A_b holds U (unit columns, sorted by descending sigma; an exactly zero column stays zero) or, with
// `keep_scaled`, U diag(sigma); sigma_b the singular values; info_b the number of sweeps (negative: not converged).
__global__ void __launch_bounds__(JSVD_THREADS, 1)
jacobi_svd_batched_kernel(int rows, int cols, double* __restrict__ A, long long lda, long long strideA,
                          double* __restrict__ sigma, long long ldsig, int* __restrict__ info, int max_sweeps, double tol,
                          int keep_scaled) {
    extern __shared__ double sm[];
    const int cs = rows | 1;
    double* a = sm;                               // cols x cs
    double* nrm = a + (size_t)cols * cs;          // cols
    int* order = reinterpret_cast<int*>(nrm + cols + (cols & 1));   // cols
    __shared__ int s_rot;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    double* Ab = A + (long long)blockIdx.x * strideA;

    __shared__ double s_red[40];
    double fro = 0.0;
    for (int idx = tid; idx < rows * cols; idx += T) {
        const int r = idx / cols, c = idx - r * cols;
        const double v = Ab[(long long)r * lda + c];
        a[c * cs + r] = v;
        fro = fma(v, v, fro);
    }
    if (tid == 0) s_rot = 0;
    // columns whose squared norm falls below (eps^2 rows) ||A||_F^2 are round-off relative to the block (a wide block
    // has cols - rows of them): they are left alone by the rotations and returned as zero columns
    fro = warp_sum(fro);
    if (lane == 0) s_red[warp] = fro;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < nwarps; ++i) t += s_red[i];
        s_red[32] = t * (DBL_EPSILON * DBL_EPSILON) * rows;
    }
    __syncthreads();
    const double tiny = s_red[32];

    const int n2 = (cols + 1) & ~1;  // players of the round-robin tournament (an odd column count gets a bye)
    const int npairs = n2 / 2;
    int sweeps = 0, converged = (cols < 2);
    while (!converged && sweeps < max_sweeps) {
        for (int round = 0; round < n2 - 1; ++round) {
            for (int pr = warp; pr < npairs; pr += nwarps) {
                int p, q;
                if (pr == 0) {
                    p = n2 - 1;
                    q = round;
                } else {
                    p = (round + pr) % (n2 - 1);
                    q = (round + n2 - 1 - pr) % (n2 - 1);
                }
                if (p > q) {
                    const int t = p;
                    p = q;
                    q = t;
                }
                if (q >= cols) continue;
                double* ap = a + p * cs;
                double* aq = a + q * cs;
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
                for (int r = lane; r < rows; r += 32) {
                    const double x = ap[r], y = aq[r];
                    alpha = fma(x, x, alpha);
                    beta = fma(y, y, beta);
                    gamma = fma(x, y, gamma);
                }
                alpha = warp_sum(alpha);
                beta = warp_sum(beta);
                gamma = warp_sum(gamma);
                if (alpha > tiny && beta > tiny && fabs(gamma) > tol * sqrt(alpha) * sqrt(beta)) {
                    const double zeta = (beta - alpha) / (2.0 * gamma);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                    for (int r = lane; r < rows; r += 32) {
                        const double x = ap[r], y = aq[r];
                        ap[r] = c * x - s * y;
                        aq[r] = s * x + c * y;
                    }
                    if (lane == 0) s_rot = 1;
                }
            }
            __syncthreads();
        }
        ++sweeps;
        const int rotated = s_rot;
        __syncthreads();
        if (tid == 0) s_rot = 0;
        __syncthreads();
        converged = !rotated;
    }

    for (int c = warp; c < cols; c += nwarps) {
        double s = 0.0;
        for (int r = lane; r < rows; r += 32) s = fma(a[c * cs + r], a[c * cs + r], s);
        s = warp_sum(s);
        if (lane == 0) nrm[c] = (s > tiny) ? sqrt(s) : 0.0;
    }
    __syncthreads();
    for (int c = tid; c < cols; c += T) {  // rank of column c in descending order (ties by index): a tiny counting sort
        const double v = nrm[c];
        int rank = 0;
        for (int j = 0; j < cols; ++j) rank += (nrm[j] > v || (nrm[j] == v && j < c)) ? 1 : 0;
        order[rank] = c;
        sigma[(long long)blockIdx.x * ldsig + rank] = v;
    }
    __syncthreads();
    for (int idx = tid; idx < rows * cols; idx += T) {
        const int r = idx / cols, k = idx - r * cols;
        const int c = order[k];
        const double v = nrm[c];
        const double scale = keep_scaled ? 1.0 : (v > 0.0 ? 1.0 / v : 0.0);
        Ab[(long long)r * lda + k] = a[c * cs + r] * scale;
    }
    if (tid == 0) info[blockIdx.x] = converged ? sweeps : -sweeps;
}

}  // namespace hfb

using namespace hfb;

extern "C" size_t hfb_chol_inverse_workspace_bytes(int64_t m) {
    if (m <= 0 || m > 1024) return 0;
    const size_t ldw = (size_t)((m + 1) & ~1LL);
    return 2 * (size_t)m * ldw * 8;
}

static int chol_inverse_impl(int64_t m, const double* G, int64_t ldg, double* S, int64_t lds, double* stat, int scale_columns,
                             void* workspace, size_t workspace_bytes, long long* prof, void* stream_);

extern "C" int hfb_chol_inverse(int64_t m, const double* G, int64_t ldg, double* S, int64_t lds, double* stat,
                                int scale_columns, void* workspace, size_t workspace_bytes, void* stream_) {
    return chol_inverse_impl(m, G, ldg, S, lds, stat, scale_columns, workspace, workspace_bytes, nullptr, stream_);
}

// Debug aid: same call, thread 0 additionally accumulates clock64() cycles per kernel section into prof[0..10] (DEVICE,
// zeroed by the caller): {scaling, copy, panel load, diagonal block, panel solve, trailing update, zero S, panel load,
// panel solve + column stage, eager update, final scaling}.
extern "C" int hfb_chol_inverse_profile(int64_t m, const double* G, int64_t ldg, double* S, int64_t lds, double* stat,
                                        int scale_columns, void* workspace, size_t workspace_bytes, int64_t* prof, void* stream_) {
    if (!prof) return HFB_E_BADARG;
    return chol_inverse_impl(m, G, ldg, S, lds, stat, scale_columns, workspace, workspace_bytes, (long long*)prof, stream_);
}

static int chol_inverse_impl(int64_t m, const double* G, int64_t ldg, double* S, int64_t lds, double* stat, int scale_columns,
                             void* workspace, size_t workspace_bytes, long long* prof, void* stream_) {
    if (m <= 0 || !G || !S || !stat || ldg < m || lds < m) return HFB_E_BADARG;
    if (m > 1024) return HFB_E_UNSUPPORTED;
    const size_t need = hfb_chol_inverse_workspace_bytes(m);
    if (!workspace || workspace_bytes < need) return HFB_E_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 7) return HFB_E_ALIGN;
    const long long ldw = (m + 1) & ~1LL;
    double* W0 = (double*)workspace;
    double* W = W0 + (size_t)m * ldw;
    const int PW = (int)((m + 9) & ~1LL);
    const size_t smem = ((size_t)CI_NB * PW + 2 * (size_t)PW + CI_NB * CI_NB + 40 + (size_t)CI_NB * PW) * 8;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(chol_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = true;
    }
    chol_inverse_kernel<<<1, CI_THREADS, smem, (cudaStream_t)stream_>>>((int)m, G, ldg, S, lds, W0, W, ldw, stat,
                                                                        scale_columns ? 1 : 0, 1.0e13, prof);
    ++g_launch_count;
    return (int)cudaGetLastError();
}

static const size_t kJacobiSmemMax = 231000;  // 227 KB opt-in limit minus the kernel's static shared memory
extern "C" int64_t hfb_jacobi_svd_max_elems(void) { return (int64_t)(kJacobiSmemMax / 8); }

extern "C" int hfb_jacobi_svd_batched(int64_t rows, int64_t cols, double* A, int64_t lda, int64_t strideA, int64_t batch,
                                      double* sigma, int64_t ldsig, int32_t* info, int32_t max_sweeps, int keep_scaled,
                                      void* stream_) {
    if (rows <= 0 || cols <= 0 || batch <= 0 || !A || !sigma || !info || lda < cols || ldsig < cols || max_sweeps <= 0)
        return HFB_E_BADARG;
    if (batch > 1 && strideA < lda * (rows - 1) + cols) return HFB_E_BADARG;
    const size_t cs = (size_t)(rows | 1);
    const size_t elems = (size_t)cols * cs + (size_t)cols + 2;
    if (elems * 8 + (size_t)cols * 4 > kJacobiSmemMax || batch > 0x7fffffffLL) return HFB_E_UNSUPPORTED;
    const size_t smem = elems * 8 + (size_t)cols * 4;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(jacobi_svd_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kJacobiSmemMax);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = true;
    }
    jacobi_svd_batched_kernel<<<(unsigned)batch, JSVD_THREADS, smem, (cudaStream_t)stream_>>>(
        (int)rows, (int)cols, A, lda, strideA, sigma, ldsig, info, max_sweeps, 1.0e-15, keep_scaled ? 1 : 0);
    ++g_launch_count;
    return (int)cudaGetLastError();
}
