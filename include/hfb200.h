/*
 * hfb200.h -- C ABI of the B200-native reduced-basis hot path (libhfb200.so).
 *
 * Drop-in boundary for the stored-data eigensolve/projection path of hIPPYflow
 * (reference = /root/reference, pure Python; there is no native FFI in the reference, so each
 * entry point cites the reference call site whose arithmetic it replaces).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates); the library
 *     never allocates or frees caller memory; workspace sizes come from *_workspace_bytes;
 *   - every entry point takes a cudaStream_t (passed as void*) and is asynchronous on it;
 *   - all dense matrices are float64, ROW-major with an explicit leading dimension (elements);
 *     a hIPPYlib MultiVector of k vectors of length n is the dense (n, k) array of
 *     hippyflow/utilities/mv_utilities.py:31-49 (column j = vector j);
 *   - TMA operands (dgemm A and B) need a 16-byte aligned base and an EVEN leading dimension;
 *   - return value: 0 ok, <0 invalid argument (HFB_E_*), >0 a cudaError_t / CUresult code.
 */
#ifndef HFB200_H
#define HFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HFB_OK 0
#define HFB_E_BADARG (-1)     /* null pointer, non-positive dimension, unknown enum            */
#define HFB_E_ALIGN (-2)      /* base not 16-byte aligned or odd leading dimension (TMA)       */
#define HFB_E_WORKSPACE (-3)  /* workspace too small                                           */
#define HFB_E_NODRIVER (-4)   /* cuTensorMapEncodeTiled not obtainable (no CUDA driver)        */
#define HFB_E_UNSUPPORTED (-5)

/* dgemm operand layouts: C[M x N] (row-major, ldc) = alpha * op(A) * op(B)                      */
#define HFB_NN 0 /* A: M x K row-major (lda>=K);  B: K x N row-major (ldb>=N)                   */
#define HFB_TN 1 /* A: K x M row-major (lda>=M), C = A^T B;  B: K x N row-major                 */
#define HFB_NT 2 /* A: M x K row-major;  B: N x K row-major (ldb>=K), C = A B^T                 */

/* Library / device info. hfb_version returns major*10000+minor*100+patch. */
int hfb_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t hfb_launch_count(void);

/*
 * FP64 tensor-core (DMMA.8x8x4) GEMM fed by TMA, deterministic split-K.
 * Replaces, on stored data:
 *   - W = X^T B  and  Y = X W  of the sample-averaged operator applies
 *       hp.LowRankOperator.mult            (hippyflow/modeling/PODProjector.py:360, one column per call)
 *       MeanJTJfromDataOperator.mult       (hippyflow/modeling/operatorWrappers.py:95-114, two einsums)
 *       H_matvec = MX @ (MX.T @ x)/n_data  (hippyflow/modeling/PODProjector.py:753-754)
 *   - Gram matrices u_data.T @ M @ u_data  (PODProjector.py:818), T = (AQ)^T Q (hIPPYlib doublePass),
 *   - lifts phi = u_data @ U               (PODProjector.py:826), U = Q V (hIPPYlib MvDSmatMult),
 *   - projections data @ encoder, J_i Psi  (dataGenerator.py:177), J_i^T (M Phi) (dataGenerator.py:170,339).
 * splits: 0 = choose automatically; otherwise the K range is cut into `splits` parts whose partial
 * tiles go to `workspace` and are summed in fixed order (bitwise run-to-run reproducible).
 */
size_t hfb_dgemm_workspace_bytes(int layout, int64_t M, int64_t N, int64_t K, int splits);
int hfb_dgemm(int layout, int64_t M, int64_t N, int64_t K, double alpha,
              const double* A, int64_t lda, const double* B, int64_t ldb,
              double* C, int64_t ldc, void* workspace, size_t workspace_bytes, int splits,
              void* stream);
/* Same with flags.  HFB_GEMM_SYMMETRIC: the caller asserts that the (M == N) result is symmetric (Gram matrices
 * X^T M X, Y^T B Y, W^T W); tiles strictly below the diagonal are not computed and are filled by a mirror kernel. */
#define HFB_GEMM_SYMMETRIC 1
/* HFB_GEMM_ACCUMULATE: C += alpha * op(A) op(B) (C is read; not combinable with HFB_GEMM_SYMMETRIC).  Used to sum the
 * per-chunk lifts Y += X_c^T W_c while the snapshots are still being uploaded. */
#define HFB_GEMM_ACCUMULATE 2
/* HFB_GEMM_B_UPPER: the caller asserts that B (K x N with K == N, layouts NN / TN) is upper triangular -- the triangular
 * inverse S of Cholesky-QR in Q = Y S.  k-blocks below a tile's last column are skipped (a TRMM: ~25 % less work at two
 * column tiles); entries of B below the diagonal are never read past the tile boundary, so they must be zero. */
#define HFB_GEMM_B_UPPER 4
size_t hfb_dgemm_ex_workspace_bytes(int layout, int64_t M, int64_t N, int64_t K, int splits, int flags);
int hfb_dgemm_ex(int layout, int64_t M, int64_t N, int64_t K, double alpha,
                 const double* A, int64_t lda, const double* B, int64_t ldb,
                 double* C, int64_t ldc, void* workspace, size_t workspace_bytes, int splits, int flags,
                 void* stream);
/* The split count hfb_dgemm would choose for splits=0. */
int hfb_dgemm_auto_splits(int layout, int64_t M, int64_t N, int64_t K);

/*
 * Strided-batch form of the same DMMA/TMA kernel over the sample axis of stored Jacobians J (N, dQ, dM) and of per-sample
 * blocks: the operands are 3-D tensor maps (third coordinate = sample; rows past one sample's extent read as zeros, never
 * the next sample), `batch` samples `strideA` / `strideB` / `strideC` elements apart (even; 0 = operand shared by all samples).
 *   mode HFB_BATCH_INDEPENDENT: C_b = alpha op(A_b) op(B_b) [+ C_b with HFB_GEMM_ACCUMULATE], one launch for all samples:
 *       JstarPhi_i = J_i^T (M Phi)   hp.MatMvTranspmult(J, MPhi, JstarPhi)   (hippyflow/modeling/dataGenerator.py:170,339)
 *       G_i = J_i J_i^T, V_i = J_i^T U_i, B_i^T = J_i^T Q_i of the per-sample randomized SVD
 *                                    hp.accuracyEnhancedSVD(J, Omega, r, s=1) (activeSubspaceProjector.py:816,1026, dataGenerator.py:187)
 *   mode HFB_BATCH_REDUCE:      C = alpha sum_b op(A_b) op(B_b): the K loop runs over (sample, k), deterministic split-K:
 *       E[J J^T] = (1/N) sum_i J_i J_i^T and its operator form sum_i J_i (J_i^T X)
 *                                    JJT / SummedListOperator(average=True) (activeSubspaceProjector.py:625-673, jacobian.py:169-193)
 * workspace: hfb_dgemm_batched_workspace_bytes (0 for HFB_BATCH_INDEPENDENT).  HFB_GEMM_SYMMETRIC is not supported here.
 */
#define HFB_BATCH_INDEPENDENT 0
#define HFB_BATCH_REDUCE 1
size_t hfb_dgemm_batched_workspace_bytes(int layout, int64_t M, int64_t N, int64_t K, int64_t batch, int mode);
int hfb_dgemm_batched(int layout, int64_t M, int64_t N, int64_t K, double alpha,
                      const double* A, int64_t lda, int64_t strideA,
                      const double* B, int64_t ldb, int64_t strideB,
                      double* C, int64_t ldc, int64_t strideC, int64_t batch, int mode, int flags,
                      void* workspace, size_t workspace_bytes, void* stream);

/*
 * Batched small GEMM over the sample axis:  C_i[M x N] = alpha * op(A_i) * B_i  for i < batch,
 * with element strides between consecutive samples (stride 0 = shared operand).
 * layout is HFB_NN or HFB_TN.  One CTA per (sample, tile); CUDA-core DFMA (operands are tiny).
 * Replaces: einsum('ij,kj->ki', noise_cov_inv, JX) (operatorWrappers.py:107-109) and the per-sample
 * products Phi^T (J_i V) of the reduced Jacobians (dataGenerator.py:170-177, north star (c)).
 */
int hfb_dgemm_batched_small(int layout, int64_t M, int64_t N, int64_t K, double alpha,
                            const double* A, int64_t lda, int64_t strideA,
                            const double* B, int64_t ldb, int64_t strideB,
                            double* C, int64_t ldc, int64_t strideC, int64_t batch, void* stream);

/*
 * CSR SpMM  C[n x m] = Mat * B[n x m]  (dense row-major), int32 CSR as exported by the reference
 * (PODProjector.py:695-697).  Replaces M_csr @ phi (PODProjector.py:750,769,830), hp.MatMvMult(M, decoder,
 * encoder) (KLEProjector.py:167-168, activeSubspaceProjector.py:452-453,559-560) and the B-applies inside
 * hIPPYlib's Borthogonalize.
 */
int hfb_csr_spmm(int64_t nrows, int64_t m, const int32_t* rowptr, const int32_t* colind,
                 const double* val, const double* B, int64_t ldb, double* C, int64_t ldc, void* stream);

/* Same product with the rows visited in the order `order` (device int32 permutation of 0..nrows-1, from
 * hfb_csr_cluster_rows_capped): a CTA then owns 64 mesh-neighbouring rows whose B rows overlap, so most B reads hit in L1.
 * Used for narrow blocks (m < 32) and for matrices whose rows are too dense for the cluster kernels below. */
int hfb_csr_spmm_ordered(int64_t nrows, int64_t m, const int32_t* rowptr, const int32_t* colind, const double* val,
                         const int32_t* order, const double* B, int64_t ldb, double* C, int64_t ldc, void* stream);

/* HOST function (host pointers, no CUDA call): greedy breadth-first clustering of the rows of a CSR matrix with symmetric
 * pattern into clusters of <= max_rows graph-neighbouring rows touching <= max_cols DISTINCT columns; writes a permutation
 * of 0..n-1 (order_out), the slot boundaries of the clusters (cluster_ptr_out, n+1 entries allocated by the caller) and
 * their number.  O(nnz).  One-time preprocessing per matrix (hippyflow_b200/linalg.py:CsrMatrix._build_plan). */
int hfb_csr_cluster_rows_capped(int64_t n, const int32_t* rowptr, const int32_t* colind, int32_t max_rows, int32_t max_cols,
                                int32_t* order_out, int32_t* cluster_ptr_out, int64_t* nclusters_out);

/* Cluster-dense DMMA SpMM, panel form (a tuning aid since round 2: HFB_SPMM_IMPL=dmma).  HOST preprocessing hfb_csr_pack_clusters packs the
 * clusters of hfb_csr_cluster_rows_capped into fixed-stride records (layout in hippyflow_b200/csrc/spmm_blob.cuh; stride from
 * hfb_csr_cluster_blob_stride; max_entries = largest number of matrix entries in one cluster; max_rows <= 16, max_cols <= 48
 * for this kernel); the caller uploads the buffer.  One CTA per cluster: the cluster's entries become a dense
 * [rows][distinct columns] block whose DMMA A-fragments stay in registers; the distinct B rows are staged panel by panel into
 * shared memory with cp.async (double-buffered) and multiplied as DMMA.8x8x4 tiles, so every staged B element is read from
 * shared memory once and the matrix never is.  B, C, blobs: 16-byte aligned; ldb, ldc even; ldb >= m rounded up to even. */
int64_t hfb_csr_cluster_blob_stride(int32_t max_rows, int32_t max_cols, int32_t max_entries);
int hfb_csr_pack_clusters(int64_t n, const int32_t* rowptr, const int32_t* colind, const double* val, const int32_t* order,
                          const int32_t* cluster_ptr, int64_t nclusters, int32_t max_rows, int32_t max_cols,
                          int32_t max_entries, void* blobs_out /* HOST, nclusters * stride bytes */);
int hfb_csr_spmm_dmma(int64_t nclusters, int64_t m, const void* blobs, int32_t max_rows, int32_t max_cols,
                      int32_t max_entries, const double* B, int64_t ldb, double* C, int64_t ldc, void* stream);

/* The same product from "fragment records" (default for 32 <= m <= 64 and m > 384): HOST preprocessing hfb_csr_pack_clusters_frag stores each
 * cluster's dense block already in DMMA A-fragment order together with its nonzero-block masks, distinct columns and result
 * rows (fixed stride hfb_csr_frag_blob_stride; max_rows <= 16, max_cols <= 48).  The kernel needs no record staging or
 * dense-block build: it requests the whole chunk (chunk_cols columns, 0 = whole rows up to 320 columns) of every distinct B
 * row with cp.async right after reading the column list, keeps the accumulators of all its column groups in registers with
 * the k-steps as the outer loop, and starts the DMMAs of k-steps 2i, 2i+1 when rows 8i..8i+7 have landed.  256-bit result
 * stores when C rows are 32-byte aligned.  B, C, blobs: 16-byte aligned; ldb, ldc even; ldb >= m rounded up to even.
 * (Five further variants were measured and retired -- profiles/r01_spmm_variants.md, sources under
 * tools/experiments/spmm_variants/.) */
int64_t hfb_csr_frag_blob_stride(int32_t max_rows, int32_t max_cols);
int hfb_csr_pack_clusters_frag(int64_t n, const int32_t* rowptr, const int32_t* colind, const double* val, const int32_t* order,
                               const int32_t* cluster_ptr, int64_t nclusters, int32_t max_rows, int32_t max_cols,
                               void* blobs_out /* HOST, nclusters * stride bytes */);
int hfb_csr_spmm_dmma_frag(int64_t nclusters, int64_t m, const void* blobs, int32_t max_rows, int32_t max_cols,
                           int32_t chunk_cols, const double* B, int64_t ldb, double* C, int64_t ldc, void* stream);
/* Ring-pipelined form over the same fragment records (results bitwise equal to hfb_csr_spmm_dmma_frag): one resident CTA per
 * SM, a producer warp keeps a ring of 2-8 cluster buffers (fragment record + whole B rows, one TMA linear copy per row
 * completing on mbarriers) full while 16 consumer warps run the DMMA k-steps and store -- the column-list / copy / multiply / store chain of a cluster
 * overlaps its neighbours' inside the SM.  Clusters of two row halves (8 < max_rows <= 16), m <= 384; other shapes return
 * HFB_E_UNSUPPORTED and the caller uses hfb_csr_spmm_dmma_frag. */
int hfb_csr_spmm_dmma_ring(int64_t nclusters, int64_t m, const void* blobs, int32_t max_rows, int32_t max_cols,
                           const double* B, int64_t ldb, double* C, int64_t ldc, void* stream);
/* Run-staged FMA form over the same clusters (hippyflow_b200/csrc/spmm_runs.cu; default for 65 <= m <= 384; same call sites:
 * `M_csr @ phi`, PODProjector.py:769,830; hp.MatMvMult(M, decoder, encoder), KLEProjector.py:167-168).  The sorted distinct
 * columns of a mesh-neighbour cluster fall into a few RUNS of consecutive B rows; consecutive rows of a row-major block are
 * contiguous, so each run is staged by ONE linear TMA copy at the block's own pitch (7 requests per cluster instead of 33),
 * and the product is evaluated entry by entry on the FP64 FMA pipe (a consumer warp owns a cluster row, its lanes own column
 * pairs) instead of as a dense DMMA block: a row's entries are summed in ascending column order with fused multiply-adds
 * from 0, the order of a sequential CSR product.
 * HOST preprocessing: hfb_csr_runs_measure -> caps_out[4] = {max_rows, max_runs, max_brow, max_entries} of the plan;
 * hfb_csr_runs_blob_stride (< 0: more than 16 rows or 32 runs per cluster -- use the fragment kernels);
 * hfb_csr_pack_clusters_runs writes nclusters * stride bytes the caller uploads.  hfb_csr_spmm_runs_slots = ring slots the
 * kernel would use for an (n, m) block of pitch ldb, 0 when the shape is unsupported (m > 384, or fewer than two slots of
 * max_brow * ldb doubles fit in shared memory: e.g. a narrow column view of a wide block); hfb_csr_spmm_runs then returns
 * HFB_E_UNSUPPORTED.  B, C, blobs 16-byte aligned, ldb and ldc even, ldb >= m + (m & 1) (the padding column of an odd
 * width is read), B != C. */
int hfb_csr_runs_measure(int64_t n, const int32_t* rowptr, const int32_t* colind, const int32_t* order,
                         const int32_t* cluster_ptr, int64_t nclusters, int32_t* caps_out /* HOST */);
int64_t hfb_csr_runs_blob_stride(int32_t max_rows, int32_t max_runs, int32_t max_entries);
int hfb_csr_pack_clusters_runs(int64_t n, const int32_t* rowptr, const int32_t* colind, const double* val, const int32_t* order,
                               const int32_t* cluster_ptr, int64_t nclusters, int32_t max_rows, int32_t max_runs,
                               int32_t max_entries, void* blobs_out /* HOST, nclusters * stride bytes */);
int32_t hfb_csr_spmm_runs_slots(int64_t m, int64_t ldb, int32_t max_rows, int32_t max_runs, int32_t max_brow,
                                int32_t max_entries);
int hfb_csr_spmm_runs(int64_t nclusters, int64_t m, const void* blobs, int32_t max_rows, int32_t max_runs, int32_t max_brow,
                      int32_t max_entries, const double* B, int64_t ldb, double* C, int64_t ldc, void* stream);

/*
 * Same sparse matrix applied to sample-major data: C[N x n] (row i = Mat * row i of X), i.e.
 * (M X)^T for symmetric M with X stored as u_data (N, n)  (PODProjector.py:750, 818).
 */
int hfb_csr_spmm_rows(int64_t nsamples, int64_t n, const int32_t* rowptr, const int32_t* colind,
                      const double* val, const double* X, int64_t ldx, double* C, int64_t ldc, void* stream);

/* out[j] = sum_i X[i,j] * Y[i,j]  (j < m): diag(X^T Y); weighted_l2_norm_vector (PODProjector.py:658-661)
 * is sqrt of this with Y = W X.  Deterministic two-stage reduction; workspace >= hfb_coldot_workspace_bytes. */
size_t hfb_coldot_workspace_bytes(int64_t n, int64_t m);
int hfb_coldot(int64_t n, int64_t m, const double* X, int64_t ldx, const double* Y, int64_t ldy,
               double* out, void* workspace, size_t workspace_bytes, void* stream);

/* out[i] = sum_j X[i,j] * Y[i,j]  (i < nrows): per-sample inner products / squared norms, e.g. the relative projection
 * errors of KLEProjector.test_errors (KLEProjector.py:262-270).  Deterministic. */
int hfb_rowdot(int64_t nrows, int64_t ncols, const double* X, int64_t ldx, const double* Y, int64_t ldy, double* out,
               void* stream);

/* X[i,j] *= s[j]  (column scaling, e.g. phi / weighted_l2_norm_vector, PODProjector.py:829). */
int hfb_colscale(int64_t n, int64_t m, double* X, int64_t ldx, const double* s, void* stream);

/* mean[j] = (1/N) sum_i X[i,j] over the sample axis (np.mean(u_data, axis=0), PODProjector.py:733);
 * deterministic; workspace >= hfb_colmean_workspace_bytes. */
size_t hfb_colmean_workspace_bytes(int64_t N, int64_t n);
int hfb_colsum(int64_t N, int64_t n, const double* X, int64_t ldx, double scale, double* out,
               void* workspace, size_t workspace_bytes, void* stream);
/* out[j] = scale * sum_i w[i] X[i,j]  (w^T X): the row vector u_shift^T (M Omega) of the implicit mean shift; same
 * two-stage deterministic reduction and workspace as hfb_colsum. */
int hfb_colsum_weighted(int64_t N, int64_t n, const double* X, int64_t ldx, const double* w, double scale, double* out,
                        void* workspace, size_t workspace_bytes, void* stream);
/* X[i,j] -= shift[j]  (u_data - u_shift, PODProjector.py:734). */
int hfb_subtract_row(int64_t N, int64_t n, double* X, int64_t ldx, const double* shift, void* stream);

/* Y[i,j] += a * x[i] * y[j]  (rank-one update of an (n x m) block): the mean-shift corrections
 * (X - 1 u^T)^T W = X^T W - u (1^T W) that stand in for u_data - u_shift (PODProjector.py:734) when the stored
 * snapshots are not modified. */
int hfb_rank1_update(int64_t n, int64_t m, double a, const double* x, const double* y, double* Y, int64_t ldy, void* stream);

/* Y = a*X + b*Y elementwise on an (n x m) block (axpy/scale of MultiVectors; 'avg' scaling of
 * collective.py:65-68 after the NCCL sum). */
int hfb_axpby(int64_t n, int64_t m, double a, const double* X, int64_t ldx, double b, double* Y, int64_t ldy,
              void* stream);

/* Y[:,j] = a[j]*X[:,j] + b[j]*Y[:,j]  (a or b may be NULL = 1): per-column axpy of block CG, the device form of
 * the Rsolver / Msolver applies inside doublePassG (activeSubspaceProjector.py:449, KLEProjector.py:163). */
int hfb_axpby_cols(int64_t n, int64_t m, const double* a, const double* X, int64_t ldx, const double* b, double* Y,
                   int64_t ldy, void* stream);
/* Y[i,:] = s[i] * X[i,:]  (Jacobi preconditioner of the block CG). */
int hfb_rowscale(int64_t n, int64_t m, const double* s, const double* X, int64_t ldx, double* Y, int64_t ldy,
                 void* stream);

/* Counter-based Gaussian/uniform fill keyed by (seed, global row, column): data generated on device is
 * independent of how samples are sharded. kind 0 = N(0,1), 1 = U(0,1). */
int hfb_fill_random(int64_t nrows, int64_t ncols, double* X, int64_t ldx, uint64_t seed, int64_t row_offset,
                    int kind, void* stream);

/*
 * Device Cholesky-QR factor: for the symmetric (m x m) Gram matrix G = Y^T B Y of a sketch (m <= 1024), one CTA computes
 *   d_j = sqrt(G_jj) (scale_columns != 0; a zero column keeps d_j^-1 = 0), Gs = D^-1 G D^-1, R = chol(Gs + shift I) upper,
 *   S = D^-1 R^-1 (upper triangular, strictly lower part zeroed), so that Q = Y S is B-orthonormal:
 * the role of hIPPYlib's MultiVector.Borthogonalize inside doublePassG / doublePass (call sites PODProjector.py:376,
 * activeSubspaceProjector.py:449-461, KLEProjector.py:163,177), done as Cholesky-QR.  The shift starts at 0 and is raised
 * (100 m eps, x100 per attempt, <= 8 attempts) until the factorisation succeeds with cond(R)^2 < 1e13 -- the same policy the
 * host path applied with LAPACK dpotrf/dtrtri, without the device -> host -> device round trip.
 * stat (DEVICE, 8 doubles): {fail, shift, cond = (max r_jj / min r_jj)^2, attempts, max |Gs - I|, max |d_j - 1|, zero columns, m}.
 * workspace >= hfb_chol_inverse_workspace_bytes(m) (two m x m scratch matrices).
 */
size_t hfb_chol_inverse_workspace_bytes(int64_t m);
int hfb_chol_inverse(int64_t m, const double* G, int64_t ldg, double* S, int64_t lds, double* stat, int scale_columns,
                     void* workspace, size_t workspace_bytes, void* stream);
/* Debug aid: the same call; thread 0 additionally accumulates clock64() cycles per kernel section into prof[0..10]
 * (DEVICE int64, zeroed by the caller): scaling, copy, panel load, diagonal block, panel solve, trailing update, zero S,
 * panel load, panel solve + column stage, eager update, final scaling (tools/bench_small_dense.py prints them). */
int hfb_chol_inverse_profile(int64_t m, const double* G, int64_t ldg, double* S, int64_t lds, double* stat, int scale_columns,
                             void* workspace, size_t workspace_bytes, int64_t* prof, void* stream);

/*
 * Batched one-sided Jacobi SVD of small blocks, one CTA per sample, block resident in shared memory
 * (rows * cols <= ~hfb_jacobi_svd_max_elems()): A_b (rows x cols, row-major, lda, strideA between samples) is overwritten
 * by its left singular vectors U_b (unit columns sorted by descending singular value; keep_scaled != 0: U_b diag(sigma_b)),
 * sigma (batch x cols, ldsig) receives the singular values, info[b] the sweeps used (negative = not converged).
 * Used by the batched randomized SVD of stored Jacobians -- hp.accuracyEnhancedSVD(J, Omega, r, s=1)
 * (activeSubspaceProjector.py:816,1026, dataGenerator.py:187): orthonormalisation Q_i of the sketches J_i Omega
 * (hIPPYlib: MultiVector.orthogonalize) and the eigen decomposition of the (l x l) matrices (Q_i^T J_i)(Q_i^T J_i)^T
 * (hIPPYlib: np.linalg.svd of the small factor).
 */
int64_t hfb_jacobi_svd_max_elems(void);
int hfb_jacobi_svd_batched(int64_t rows, int64_t cols, double* A, int64_t lda, int64_t strideA, int64_t batch,
                           double* sigma, int64_t ldsig, int32_t* info, int32_t max_sweeps, int keep_scaled, void* stream);

/* Measure the FP64 tensor-pipe ceiling of the current device (register-resident DMMA.8x8x4 loop, 8 warps/SM,
 * best of 5): the roofline denominator bench.py reports the GEMM against.  Synchronises the stream.
 * scratch: >= SMs*256*8 bytes of device memory; *tflops_out is a HOST double. */
int hfb_measure_dmma_peak(double* scratch, size_t scratch_bytes, double* tflops_out, void* stream);

/*
 * Peer exchange: the allreduce of the (n x m) sketch  Y = sum_g X_g^T W_g  over the GPUs of one NVLink domain, fused with
 * the lift GEMM that produces it.  Replaces  CollectiveOperator.mult / MatrixMultCollectiveOperator.matMvMult
 * (hippyflow/collectives/collectiveOperator.py:31-38,73-80) -> MultipleSamePartitioningPDEsCollective.allReduce
 * (hippyflow/collectives/collective.py:61-71,108-111: one MPI Allreduce per column, host copies per vector) for the
 * stored-data operators of PODProjector.py:360-363 / activeSubspaceProjector.py:427-431.
 *
 * Each rank owns one exchange buffer (hfb_peer_alloc: cudaMalloc'd, zeroed, shareable through CUDA IPC) laid out by the
 * caller as  [flags: 16 x uint64 | slots: nranks x (block_rows x ld) | result blocks: (nranks*block_rows) x ld each];
 * the handles (hfb_peer_get_handle, 64 bytes) are exchanged out of band (torch.distributed all_gather_object) and mapped
 * with hfb_peer_open.  One exchange of rows [0, n) split into nranks blocks of block_rows (a multiple of 128) is
 *   1. hfb_dgemm_peer         rank g's partial tile of block o is stored from the accumulators straight into slot g of
 *                             rank o's buffer (256-bit st.global on peer-mapped addresses inside the DMMA kernel; no local
 *                             copy of the partial sketch Y_g);
 *   2. hfb_peer_barrier       system-scope release/acquire flags: every rank's pushes have landed;
 *   3. hfb_peer_reduce_bcast  the owner sums its nranks slots in FIXED order (bitwise reproducible, identical on all ranks)
 *                             and stores the result into its rows of EVERY rank's result block (all-gather by push);
 *   4. hfb_peer_barrier       every owner's rows have landed: the result block is complete on every rank.
 * NVLink bytes per rank and exchange: (nranks-1)/nranks * n * ld * 8 pushed by the GEMM + the same pushed by the reduction
 * (the two halves of a bandwidth-optimal allreduce); the first half overlaps the GEMM tile by tile.
 * hfb_peer_barrier: flag_ptrs[r] = address of rank r's 16 flag words (peer-mapped for r != me); epoch must grow with every
 * call and be the same on all ranks; a wait longer than timeout_s (0 = no limit) traps the kernel.  mode = HFB_PEER_SIGNAL
 * (publish only), HFB_PEER_WAIT (wait only) or both (a barrier).
 * slot_ptrs[o] (hfb_dgemm_peer) = address of THIS rank's slot inside rank o's buffer; ld_slot = ld of the slots.
 * y_ptrs[r] (hfb_peer_reduce_bcast) = address of the owner's FIRST row inside rank r's result block (leading dimension ldy).
 */
#define HFB_PEER_HANDLE_BYTES 64
#define HFB_PEER_MAX_RANKS 16
#define HFB_PEER_SIGNAL 1
#define HFB_PEER_WAIT 2
int hfb_peer_alloc(size_t bytes, void** ptr);
int hfb_peer_free(void* ptr);
int hfb_peer_get_handle(void* ptr, unsigned char* handle64);
int hfb_peer_open(const unsigned char* handle64, void** ptr);
int hfb_peer_close(void* ptr);
int hfb_dgemm_peer(int layout, int64_t M, int64_t N, int64_t K, double alpha,
                   const double* A, int64_t lda, const double* B, int64_t ldb,
                   double* const* slot_ptrs, int nranks, int64_t block_rows, int64_t ld_slot, void* stream);
int hfb_peer_barrier(void* const* flag_ptrs, int me, int nranks, uint64_t epoch, double timeout_s, int mode, void* stream);
int hfb_peer_reduce_bcast(const double* slots, int64_t slot_stride, int nranks, int me, int64_t rows, int64_t cols, int64_t ld,
                          double* const* y_ptrs, int64_t ldy, void* stream);

/*
 * HOST helper of the upload path: copy `bytes` from pageable host memory into a pinned staging buffer with `nthreads`
 * threads and non-temporal stores (the staging buffers are only read by the DMA engine afterwards).  The reference consumes
 * u_data / J in place on the host (PODProjector.py:726, operatorWrappers.py:62-64); this is the host half of getting a plain
 * NumPy array across PCIe at link rate (the device half is cudaMemcpyAsync from the staging ring).  Blocking; no CUDA call.
 */
int hfb_host_copy(void* dst, const void* src, size_t bytes, int nthreads);

#ifdef __cplusplus
}
#endif
#endif /* HFB200_H */
