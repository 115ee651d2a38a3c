"""Host preprocessing of the cluster SpMM kernels (no GPU): hfb_csr_cluster_rows_capped + hfb_csr_pack_clusters must
describe exactly the matrix they were given.  The packed records are decoded here the way csr_spmm_dmma_kernel
decodes them (dense [distinct column][cluster row] block per cluster) and the matrix is rebuilt bit-exactly."""
import numpy as np
import pytest
import scipy.sparse as sp

from hippyflow_b200 import _lib as K, synthetic as syn


def _r4(x):
    return (x + 3) // 4 * 4


def _decode(M, max_rows, max_cols):
    M = M.tocsr()
    n = M.shape[0]
    order, cptr = K.csr_cluster_rows_capped(M.indptr, M.indices, max_rows, max_cols)
    ncl = cptr.size - 1
    assert sorted(order.tolist()) == list(range(n))                    # a permutation: every row in exactly one cluster
    counts = np.diff(M.indptr)[order]
    csum = np.concatenate([[0], np.cumsum(counts)])
    max_entries = int((csum[cptr[1:]] - csum[cptr[:-1]]).max())
    mr = int(np.diff(cptr).max())
    assert mr <= max_rows
    blobs = K.csr_pack_clusters(M.indptr, M.indices, M.data, order, cptr, mr, max_cols, max_entries)
    stride = int(K.lib().hfb_csr_cluster_blob_stride(mr, max_cols, max_entries))
    assert stride % 128 == 0 and blobs.size == ncl * stride
    off_rowoff = 16
    off_outrow = off_rowoff + 4 * _r4(mr + 1)
    off_cols = off_outrow + 4 * _r4(mr)
    off_ent = off_cols + 4 * _r4(max_cols)
    rows, cols, vals = [], [], []
    for c in range(ncl):
        blob = blobs[c * stride:(c + 1) * stride]
        nrow, ncol, nent = (int(v) for v in blob[:12].view(np.int32))
        assert 0 < nrow <= mr and ncol <= max_cols and nent <= max_entries
        rowoff = blob[off_rowoff:off_rowoff + 4 * (nrow + 1)].view(np.int32)
        outrow = blob[off_outrow:off_outrow + 4 * nrow].view(np.int32)
        gcols = blob[off_cols:off_cols + 4 * ncol].view(np.int32)
        assert np.unique(gcols).size == ncol                               # DISTINCT columns
        ent = blob[off_ent:off_ent + 16 * nent]
        v = ent.view(np.float64)[0::2]
        lr = ent.view(np.int32).reshape(-1, 4)[:, 2:]
        assert rowoff[0] == 0 and rowoff[nrow] == nent
        # the local row field agrees with the row offsets
        assert np.array_equal(lr[:, 1], np.repeat(np.arange(nrow), np.diff(rowoff)))
        D = np.zeros((max_cols, mr))
        D[lr[:, 0], lr[:, 1]] = v
        jj, rr = np.nonzero(D[:ncol, :nrow])
        rows.append(outrow[rr]); cols.append(gcols[jj]); vals.append(D[jj, rr])
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=M.shape)
    return A


@pytest.mark.parametrize("caps", [(16, 32), (8, 20), (12, 24), (32, 64), (5, 20)])
def test_packed_clusters_rebuild_mesh_matrix(caps):
    M = syn.p1_mass_matrix(23, 31)
    A = _decode(M, *caps)
    assert (A != M.tocsr()).nnz == 0                                       # bit-exact, same pattern


def test_packed_clusters_rebuild_irregular_matrix():
    rng = np.random.default_rng(5)
    n = 3000
    A0 = sp.random(n, n, density=3.0 / n, random_state=11, format="csr")
    A0 = (A0 + A0.T + sp.diags(rng.standard_normal(n))).tolil()
    for r in (0, 99, n - 1):
        A0[r, :] = 0                                                       # empty rows
    A0 = A0.tocsr()
    A0.eliminate_zeros()
    A = _decode(A0, 16, 64)
    assert (A != A0).nnz == 0


def test_pack_rejects_bad_arguments():
    M = syn.p1_mass_matrix(8, 8).tocsr()
    order, cptr = K.csr_cluster_rows_capped(M.indptr, M.indices, 16, 32)
    with pytest.raises(K.HfbError):     # column budget smaller than what the clusters touch
        K.csr_pack_clusters(M.indptr, M.indices, M.data, order, cptr, 16, 4, 200)
    with pytest.raises(K.HfbError):     # entry budget too small
        K.csr_pack_clusters(M.indptr, M.indices, M.data, order, cptr, 16, 32, 3)


@pytest.mark.parametrize("caps", [(16, 32), (8, 24), (12, 24), (16, 48), (8, 16), (5, 20)])
def test_fragment_records_rebuild_matrix(caps):
    """hfb_csr_pack_clusters_frag: the DMMA A-fragment records, decoded with the kernel's lane map and block masks, give
    back the matrix bit for bit (mesh matrix and a ragged one with empty rows)."""
    from cpu_device_shim import decode_frag_blobs
    rng = np.random.default_rng(8)
    n = 2500
    A0 = sp.random(n, n, density=2.0 / n, random_state=3, format="csr")
    A0 = (A0 + A0.T + sp.diags(rng.standard_normal(n))).tolil()
    A0[7, :] = 0
    A0 = A0.tocsr()
    A0.eliminate_zeros()
    for M in (syn.p1_mass_matrix(23, 31).tocsr(), A0):
        order, cptr = K.csr_cluster_rows_capped(M.indptr, M.indices, *caps)
        mr = int(np.diff(cptr).max())
        distinct = max(np.unique(M.indices[np.concatenate([np.arange(M.indptr[r], M.indptr[r + 1]) for r in order[a:b]])]).size
                       if M.indptr[order[a:b] + 1].sum() > M.indptr[order[a:b]].sum() else 0
                       for a, b in zip(cptr[:-1], cptr[1:]))
        blobs = K.csr_pack_clusters_frag(M.indptr, M.indices, M.data, order, cptr, mr, distinct)
        A = decode_frag_blobs(K, blobs, mr, distinct, M.shape[0])
        assert (A != M).nnz == 0


def test_fragment_records_limits():
    M = syn.p1_mass_matrix(8, 8).tocsr()
    order, cptr = K.csr_cluster_rows_capped(M.indptr, M.indices, 32, 64)
    with pytest.raises(K.HfbError):                      # more rows / columns than the A fragments can hold
        K.csr_pack_clusters_frag(M.indptr, M.indices, M.data, order, cptr, 32, 64)
    order, cptr = K.csr_cluster_rows_capped(M.indptr, M.indices, 16, 32)
    with pytest.raises(K.HfbError):                      # column budget smaller than what the clusters touch
        K.csr_pack_clusters_frag(M.indptr, M.indices, M.data, order, cptr, 16, 8)


def test_peer_exchange_block_rows():
    """Owner blocks of the fused lift + exchange (hippyflow_b200/peer.py): multiples of the 128-row GEMM tile that cover the rows."""
    from hippyflow_b200.peer import block_rows_for
    for n, P in [(263169, 8), (263169, 2), (1002001, 8), (65536, 8), (5041, 2), (130, 16), (128, 1)]:
        b = block_rows_for(n, P)
        assert b % 128 == 0 and b * P >= n and (b - 128) * P < n
    assert block_rows_for(263169, 8) == 33024


# ------------------------------------------------------------------------------------------- run records (spmm_runs.cu)
def _runs_apply(M, Bfull, m, max_rows, max_cols, plan_from=None):
    """Interpret the run records exactly the way csr_spmm_runs_kernel does: stage the runs of B rows as linear copies of
    (len - 1) * ldb + width doubles at the block's pitch, then walk each row's entry list with fused multiply-adds from 0.0
    in record order.  ``Bfull`` is the (n, ldb) storage of an (n, m) block."""
    M = M.tocsr()
    n, ldb = Bfull.shape
    width = m + (m & 1)
    P = M if plan_from is None else plan_from.tocsr()          # the matrix whose pattern is clustered
    order, cptr = K.csr_cluster_rows_capped(P.indptr, P.indices, max_rows, max_cols)
    blobs, caps = K.csr_pack_clusters_runs(M.indptr, M.indices, M.data, order, cptr)
    ncl, stride, mr = cptr.size - 1, caps["stride"], caps["max_rows"]
    assert stride % 128 == 0 and blobs.size == ncl * stride
    off_outrow = 16
    off_goff = off_outrow + 4 * _r4(mr)
    off_runs = off_goff + 4 * _r4(mr + 1)
    off_ent = (off_runs + 8 * caps["max_runs"] + 15) // 16 * 16
    flat = Bfull.reshape(-1)
    C = np.full((n, m), np.nan)
    seen_rows = []
    tot_runs = 0
    for c in range(ncl):
        blob = blobs[c * stride:(c + 1) * stride]
        nrow, nrun, nbrow, nent = (int(v) for v in blob[:16].view(np.int32))
        assert 0 < nrow <= mr and nrun <= caps["max_runs"] and nbrow <= caps["max_brow"] and nent <= caps["max_entries"]
        outrow = blob[off_outrow:off_outrow + 4 * nrow].view(np.int32)
        assert np.all(np.diff(outrow) > 0)                               # rows sorted ascending
        goff = blob[off_goff:off_goff + 4 * (mr + 1)].view(np.int32)
        runs = blob[off_runs:off_runs + 8 * nrun].view(np.int32).reshape(-1, 2)
        stage = np.full(caps["max_brow"] * ldb, np.nan)                  # shared-memory staging area (doubles)
        tx = 0
        for start, lo in runs:
            ln, off = int(lo) & 0xffff, (int(lo) >> 16) & 0xffff
            cnt = (ln - 1) * ldb + width
            assert start * ldb + cnt <= flat.size                        # the copy stays inside the block
            assert off * ldb + cnt <= stage.size
            stage[off * ldb:off * ldb + cnt] = flat[start * ldb:start * ldb + cnt]
            tx += cnt
        assert tx == (nbrow - nrun) * ldb + nrun * width                 # the producer's expect_tx byte count / 8
        tot_runs += nrun
        assert goff[0] == 0 and goff[nrow] == nent and np.all(goff[nrow:] == nent)
        for r in range(nrow):
            acc = np.zeros(width)
            last = -1
            for e in range(goff[r], goff[r + 1]):
                ent = blob[off_ent + e * 16:off_ent + (e + 1) * 16]
                slot = int(ent[:4].view(np.int32)[0])
                v = float(ent[8:].view(np.float64)[0])
                assert last < slot < nbrow                               # ascending columns, each staged
                last = slot
                b = stage[slot * ldb:slot * ldb + width]
                assert not np.isnan(b[:m]).any()                         # every staged row an entry names was copied
                acc = v * b + acc
            C[outrow[r]] = acc[:m]
            seen_rows.append(int(outrow[r]))
    assert sorted(seen_rows) == list(range(n))                           # every result row written exactly once
    return C, tot_runs / ncl


@pytest.mark.parametrize("caps", [(16, 32), (12, 24), (16, 48), (5, 20)])
@pytest.mark.parametrize("m,pad", [(10, 0), (37, 1), (74, 6)])
def test_run_records_reproduce_the_product(caps, m, pad):
    M = syn.p1_mass_matrix(23, 31).tocsr()
    n = M.shape[0]
    ldb = m + (m & 1) + pad * 2
    rng = np.random.default_rng(m)
    Bfull = rng.standard_normal((n, ldb))
    C, runs_per_cluster = _runs_apply(M, Bfull, m, caps[0], caps[1])
    np.testing.assert_allclose(C, M @ Bfull[:, :m], rtol=1e-13, atol=1e-16)
    assert runs_per_cluster < caps[1] / 2                                 # mesh clusters: a handful of runs, not one per column


def test_run_records_irregular_matrix_duplicates_and_limits():
    """Ragged non-mesh pattern with empty rows; a non-canonical CSR with duplicate entries (summed); a plan with more than
    32 runs in a cluster is refused (the caller keeps the fragment kernels for it)."""
    rng = np.random.default_rng(3)
    n = 3000
    A = sp.random(n, n, density=4.0 / n, random_state=7, format="csr")
    A = (A + A.T + sp.diags(rng.standard_normal(n))).tolil()
    for r in (0, 17, n - 1):
        A[r, :] = 0
    A = A.tocsr()
    A.eliminate_zeros()
    B = rng.standard_normal((n, 12))
    C, _ = _runs_apply(A, B, 12, 16, 32)
    np.testing.assert_allclose(C, A @ B, rtol=1e-12, atol=1e-14)
    assert np.all(C[[0, 17, n - 1]] == 0.0)
    # duplicates: the same pattern stored twice with half the values, columns unsorted within the rows
    indptr = 2 * A.indptr
    indices = np.concatenate([np.concatenate([A.indices[a:b][::-1], A.indices[a:b]]) for a, b in zip(A.indptr[:-1], A.indptr[1:])])
    data = np.concatenate([np.concatenate([0.5 * A.data[a:b][::-1], 0.5 * A.data[a:b]]) for a, b in zip(A.indptr[:-1], A.indptr[1:])])
    D = sp.csr_matrix((data, indices, indptr), shape=A.shape)
    assert not D.has_canonical_format
    C2, _ = _runs_apply(D, B, 12, 16, 32, plan_from=A)
    np.testing.assert_allclose(C2, A @ B, rtol=1e-12, atol=1e-14)
    order, cptr = K.csr_cluster_rows_capped(A.indptr, A.indices, 16, 64)
    with pytest.raises(K.HfbError):
        K.csr_pack_clusters_runs(A.indptr, A.indices, A.data, order, cptr)


from hypothesis import given, settings, strategies as st, HealthCheck


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(n=st.integers(40, 300), per_row=st.floats(0.5, 5.0), band=st.integers(1, 40), caps=st.sampled_from([(16, 32), (16, 48), (8, 24), (3, 12)]),
       m=st.integers(1, 20), pad=st.integers(0, 3), seed=st.integers(0, 10 ** 6))
def test_run_records_random_symmetric_patterns(n, per_row, band, caps, m, pad, seed):
    """Property: for any symmetric sparsity pattern (banded + random long-range couplings, empty rows allowed) the run records
    either reproduce the product exactly as the kernel would evaluate it, or the packer refuses the plan (more than 32 runs
    or 16 rows in a cluster) -- never a silently wrong record."""
    rng = np.random.default_rng(seed)
    nnz = int(per_row * n)
    i = rng.integers(0, n, nnz)
    j = np.clip(i + rng.integers(-band, band + 1, nnz), 0, n - 1)
    far = rng.random(nnz) < 0.1
    j[far] = rng.integers(0, n, int(far.sum()))
    A = sp.coo_matrix((rng.standard_normal(nnz), (i, j)), shape=(n, n)).tocsr()
    A = (A + A.T).tocsr()
    A.sum_duplicates()
    try:
        order, cptr = K.csr_cluster_rows_capped(A.indptr, A.indices, caps[0], caps[1])
    except K.HfbError:
        return                                                               # a row denser than the column budget
    ldb = m + (m & 1) + 2 * pad
    Bfull = rng.standard_normal((n, ldb))
    try:
        C, _ = _runs_apply(A, Bfull, m, caps[0], caps[1])
    except K.HfbError as e:
        assert "unsupported" in str(e)
        return
    np.testing.assert_allclose(C, A @ Bfull[:, :m], rtol=1e-12, atol=1e-13)
