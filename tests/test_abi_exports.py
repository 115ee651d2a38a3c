"""CPU: the C-ABI shared library builds, loads, and exports every symbol include/hfb200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hfb200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hfb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("hfb_dgemm", "hfb_csr_spmm", "hfb_coldot", "hfb_colsum", "hfb_dgemm_batched_small"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from hippyflow_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), "libhfb200.so does not export %s" % s
    assert sorted(_lib.EXPORTED) == declared_symbols()
    assert _lib.lib().hfb_version() >= 100


def test_argument_validation_without_gpu():
    """Invalid arguments are rejected before any CUDA call."""
    from hippyflow_b200 import _lib
    L = _lib.lib()
    assert L.hfb_dgemm(7, 4, 4, 4, 1.0, None, 4, None, 4, None, 4, None, 0, 0, None) == -1
    assert L.hfb_dgemm(0, 4, 4, 4, 1.0, None, 4, None, 4, None, 4, None, 0, 0, None) == -1
    assert L.hfb_dgemm(0, 4, 4, 4, 1.0, 16, 3, 32, 4, 64, 4, None, 0, 0, None) == -1   # lda < K
    assert L.hfb_dgemm(0, 4, 4, 4, 1.0, 16, 5, 32, 4, 64, 4, None, 0, 0, None) == -2   # odd lda
    assert L.hfb_dgemm(0, 4, 4, 4, 1.0, 24, 4, 32, 4, 64, 4, None, 0, 0, None) == -2   # misaligned A
    assert L.hfb_csr_spmm(0, 4, None, None, None, None, 4, None, 4, None) == -1
    # the cluster SpMM entry points: null pointers, caps beyond the kernels' budgets, misaligned / too narrow operands
    for fn in (L.hfb_csr_spmm_dmma_frag,):
        assert fn(10, 266, None, 16, 32, 0, 16, 272, 32, 272, None) == -1                 # no records
        assert fn(10, 266, 64, 16, 32, 0, 16, 272, 16, 272, None) == -1                   # B == C
        assert fn(10, 266, 64, 16, 32, 0, 16, 272, 32, 200, None) == -1                   # ldc < m
        assert fn(10, 267, 64, 16, 32, 0, 16, 267, 32, 272, None) == -1                   # ldb < m rounded up to even
        assert fn(10, 266, 64, 16, 32, 0, 24, 272, 32, 272, None) == -2                   # B not 16-byte aligned
        assert fn(10, 266, 64, 16, 32, 0, 16, 273, 32, 272, None) == -2                   # odd ldb
        assert fn(10, 266, 64, 32, 64, 0, 16, 272, 32, 272, None) == -5                   # clusters too large for the A fragments
    assert L.hfb_csr_spmm_dmma(10, 266, 64, 32, 64, 200, 16, 272, 32, 272, None) == -5
    assert L.hfb_csr_frag_blob_stride(16, 32) == 4352 and L.hfb_csr_frag_blob_stride(8, 24) == 1792
    assert L.hfb_csr_frag_blob_stride(17, 32) < 0 and L.hfb_csr_frag_blob_stride(16, 49) < 0
    # run-staged FMA SpMM: signature (nclusters, m, blobs, max_rows, max_runs, max_brow, max_entries, B, ldb, C, ldc, stream)
    runs = L.hfb_csr_spmm_runs
    assert runs(10, 266, None, 16, 11, 32, 112, 16, 272, 32, 272, None) == -1              # no records
    assert runs(10, 266, 64, 16, 11, 32, 112, 16, 272, 16, 272, None) == -1                # B == C
    assert runs(10, 267, 64, 16, 11, 32, 112, 16, 267, 32, 272, None) == -1                # ldb < m rounded up to even
    assert runs(10, 266, 64, 16, 11, 32, 112, 24, 272, 32, 272, None) == -2                # B not 16-byte aligned
    assert runs(10, 266, 64, 16, 11, 32, 112, 16, 273, 32, 272, None) == -2                # odd ldb
    assert runs(10, 400, 64, 16, 11, 32, 112, 16, 400, 32, 400, None) == -5                # m > 384
    assert runs(10, 266, 64, 17, 11, 32, 112, 16, 272, 32, 272, None) == -5                # more rows than consumer warps
    assert runs(10, 266, 64, 16, 33, 40, 112, 16, 272, 32, 272, None) == -5                # more runs than producer lanes
    assert runs(10, 266, 64, 16, 11, 32, 112, 16, 4096, 32, 272, None) == -5               # pitch too wide for two ring slots
    assert L.hfb_csr_runs_blob_stride(16, 11, 112) == 2048
    assert L.hfb_csr_runs_blob_stride(17, 11, 112) < 0 and L.hfb_csr_runs_blob_stride(16, 33, 112) < 0
    assert L.hfb_csr_spmm_runs_slots(266, 266, 16, 11, 32, 112) == 3
    assert L.hfb_csr_spmm_runs_slots(138, 138, 16, 11, 32, 112) == 6 and L.hfb_csr_spmm_runs_slots(74, 74, 16, 11, 32, 112) == 8
    assert L.hfb_csr_spmm_runs_slots(266, 4096, 16, 11, 32, 112) == 0 and L.hfb_csr_spmm_runs_slots(385, 386, 16, 11, 32, 112) == 0
    # round-2 entry points: strided-batch GEMM, device Cholesky-QR factor, batched Jacobi SVD
    assert L.hfb_dgemm_batched(0, 4, 4, 4, 1.0, 16, 4, 16, 32, 4, 16, 64, 4, 8, 3, 0, 0, None, 0, None) == -1      # outputs overlap
    assert L.hfb_dgemm_batched(0, 4, 4, 4, 1.0, 16, 4, 15, 32, 4, 16, 64, 4, 16, 3, 0, 0, None, 0, None) == -2     # odd batch stride
    assert L.hfb_dgemm_batched(0, 4, 4, 4, 1.0, 16, 4, 16, 32, 4, 16, 64, 4, 16, 3, 7, 0, None, 0, None) == -1     # unknown mode
    assert L.hfb_dgemm_batched(0, 4, 4, 4, 1.0, 16, 4, 16, 32, 4, 16, 64, 4, 16, 3, 0, 1, None, 0, None) == -5     # symmetric flag
    assert L.hfb_dgemm_batched(0, 4, 4, 4, 1.0, 16, 4, 0, 32, 4, 0, 64, 4, 0, 3, 1, 0, None, 0, None) == -1        # reduce over shared operands
    assert L.hfb_chol_inverse(2000, 16, 2000, 32, 2000, 64, 1, None, 0, None) == -5                                 # m > 1024
    assert L.hfb_chol_inverse(8, 16, 4, 32, 8, 64, 1, None, 0, None) == -1                                          # ldg < m
    assert L.hfb_chol_inverse(8, 16, 8, 32, 8, 64, 1, None, 0, None) == -3                                          # no workspace
    assert L.hfb_chol_inverse_workspace_bytes(266) == 2 * 266 * 266 * 8 and L.hfb_chol_inverse_workspace_bytes(5000) == 0
    assert L.hfb_jacobi_svd_batched(400, 400, 16, 400, 160000, 2, 32, 400, 64, 30, 0, None) == -5                   # does not fit shared memory
    assert L.hfb_jacobi_svd_batched(10, 4, 16, 2, 40, 2, 32, 4, 64, 30, 0, None) == -1                              # lda < cols
    # peer exchange (fused lift + reduce-scatter over NVLink): rank tables, block alignment, epochs
    import ctypes as C
    two = (C.c_void_p * 2)(C.c_void_p(4096), C.c_void_p(8192))
    bad = (C.c_void_p * 2)(C.c_void_p(4096), C.c_void_p(8200))
    assert L.hfb_dgemm_peer(1, 1000, 32, 64, 1.0, 16, 1000, 32, 32, None, 2, 512, 32, None) == -1               # no slot table
    assert L.hfb_dgemm_peer(1, 1000, 32, 64, 1.0, 16, 1000, 32, 32, two, 2, 500, 32, None) == -1                # block not a multiple of 128
    assert L.hfb_dgemm_peer(1, 1000, 32, 64, 1.0, 16, 1000, 32, 32, two, 2, 384, 32, None) == -1                # blocks do not cover M
    assert L.hfb_dgemm_peer(1, 1000, 32, 64, 1.0, 16, 1000, 32, 32, two, 17, 512, 32, None) == -1               # more ranks than the table holds
    assert L.hfb_dgemm_peer(1, 1000, 32, 64, 1.0, 16, 1000, 32, 32, bad, 2, 512, 32, None) == -2                # slot not 16-byte aligned
    assert L.hfb_dgemm_peer(1, 1000, 32, 64, 1.0, 16, 1000, 32, 32, two, 2, 512, 33, None) == -2                # odd slot ld
    assert L.hfb_dgemm_peer(0, 1000, 32, 64, 1.0, 16, 64, 32, 32, two, 2, 512, 32, None) in (-5, -4)            # TN only (or no driver here)
    assert L.hfb_peer_barrier(None, 0, 2, 1, 1.0, 3, None) == -1
    assert L.hfb_peer_barrier(two, 2, 2, 1, 1.0, 3, None) == -1                                                 # rank out of range
    assert L.hfb_peer_barrier(two, 0, 2, 0, 1.0, 3, None) == -1                                                 # epoch 0 is the initial flag value
    assert L.hfb_peer_barrier(two, 0, 2, 1, 1.0, 0, None) == -1                                                 # neither signal nor wait
    assert L.hfb_peer_reduce_bcast(16, 100, 2, 0, 10, 20, 20, two, 20, None) == -1                              # slots overlap
    assert L.hfb_peer_reduce_bcast(16, 210, 2, 0, 10, 20, 21, two, 21, None) == -2                              # odd ld
    assert L.hfb_peer_reduce_bcast(16, 200, 2, 2, 10, 20, 20, two, 20, None) == -1                              # rank out of range
    assert L.hfb_peer_reduce_bcast(16, 200, 2, 0, 10, 20, 20, bad, 20, None) == -2                              # result block not 16-byte aligned
    assert L.hfb_peer_reduce_bcast(16, 200, 2, 0, 10, 20, 20, two, 10, None) == -1                              # ldy < cols
    assert L.hfb_peer_alloc(0, None) == -1 and L.hfb_peer_free(None) == -1 and L.hfb_peer_open(None, None) == -1
    assert L.hfb_dgemm_workspace_bytes(0, 128, 16, 4096, 4) == 4 * 128 * 16 * 8
    assert L.hfb_dgemm_auto_splits(0, 4096, 266, 263169) >= 2


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (tier rule 3)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hippyflow_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                # ... nor through the CPU test double of tests/ (numbers produced under it say nothing about the kernels)
                assert "cpu_device_shim" not in src and "emulated_device" not in src, f


def test_host_copy_threads_and_tails():
    """hfb_host_copy (pageable -> pinned leg of the staged upload): every byte arrives for sizes around the thread / cache-line
    boundaries, unaligned destinations and thread counts above the chunk count."""
    import numpy as np
    import torch
    from hippyflow_b200 import _lib
    assert _lib.lib().hfb_host_copy(None, None, 10, 1) == -1
    rng = np.random.default_rng(0)
    for count, threads, offset in [(1, 4, 0), (7, 1, 0), (8, 3, 1), (131072, 4, 0), (131072 + 5, 4, 1), (300001, 7, 0),
                                   (1 << 18, 64, 0), (263169 * 3, 8, 0)]:
        src = torch.from_numpy(rng.standard_normal(count))
        dst = torch.full((count + 4,), -7.0, dtype=torch.float64)
        _lib.host_copy_(dst[offset:offset + count], src, threads)
        assert torch.equal(dst[offset:offset + count], src)
        assert float(dst[offset + count]) == -7.0 and (offset == 0 or float(dst[0]) == -7.0)      # nothing written outside
    with pytest.raises(_lib.HfbError):
        _lib.host_copy_(torch.zeros(4), torch.zeros(4, dtype=torch.float64), 1)                     # dtype mismatch
