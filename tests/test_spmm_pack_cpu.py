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
