"""ctypes binding of libhfb200.so (the C ABI in include/hfb200.h) plus thin torch-tensor wrappers.

PyTorch is plumbing here: it owns device memory and streams; every numerical operation of the hot
path goes through the C ABI below.  There is NO fallback: if the shared library is missing or a
call returns an error code the wrappers raise.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhfb200.so")

HFB_NN, HFB_TN, HFB_NT = 0, 1, 2

_ERRORS = {
    -1: "invalid argument",
    -2: "operand not 16-byte aligned or odd leading dimension",
    -3: "workspace too small",
    -4: "CUDA driver entry point cuTensorMapEncodeTiled unavailable",
    -5: "unsupported size",
}


class HfbError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libhfb200.so (built by ``__graft_entry__.build()`` / ``make -C hippyflow_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HfbError(
            "libhfb200.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the hot path)")
    L = ctypes.CDLL(LIB_PATH)
    i64, i32, dbl, vp, sz, u64 = (ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t,
                                  ctypes.c_uint64)
    sigs = {
        "hfb_version": (i32, []),
        "hfb_launch_count": (i64, []),
        "hfb_dgemm_workspace_bytes": (sz, [i32, i64, i64, i64, i32]),
        "hfb_dgemm_auto_splits": (i32, [i32, i64, i64, i64]),
        "hfb_dgemm": (i32, [i32, i64, i64, i64, dbl, vp, i64, vp, i64, vp, i64, vp, sz, i32, vp]),
        "hfb_dgemm_ex_workspace_bytes": (sz, [i32, i64, i64, i64, i32, i32]),
        "hfb_dgemm_ex": (i32, [i32, i64, i64, i64, dbl, vp, i64, vp, i64, vp, i64, vp, sz, i32, i32, vp]),
        "hfb_dgemm_batched_workspace_bytes": (sz, [i32, i64, i64, i64, i64, i32]),
        "hfb_dgemm_batched": (i32, [i32, i64, i64, i64, dbl, vp, i64, i64, vp, i64, i64, vp, i64, i64, i64, i32, i32, vp, sz, vp]),
        "hfb_dgemm_batched_small": (i32, [i32, i64, i64, i64, dbl, vp, i64, i64, vp, i64, i64, vp, i64, i64, i64, vp]),
        "hfb_csr_spmm": (i32, [i64, i64, vp, vp, vp, vp, i64, vp, i64, vp]),
        "hfb_csr_spmm_ordered": (i32, [i64, i64, vp, vp, vp, vp, vp, i64, vp, i64, vp]),
        "hfb_csr_cluster_rows_capped": (i32, [i64, vp, vp, i32, i32, vp, vp, ctypes.POINTER(ctypes.c_int64)]),
        "hfb_csr_cluster_blob_stride": (i64, [i32, i32, i32]),
        "hfb_csr_pack_clusters": (i32, [i64, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp]),
        "hfb_csr_spmm_dmma": (i32, [i64, i64, vp, i32, i32, i32, vp, i64, vp, i64, vp]),
        "hfb_csr_frag_blob_stride": (i64, [i32, i32]),
        "hfb_csr_pack_clusters_frag": (i32, [i64, vp, vp, vp, vp, vp, i64, i32, i32, vp]),
        "hfb_csr_spmm_dmma_frag": (i32, [i64, i64, vp, i32, i32, i32, vp, i64, vp, i64, vp]),
        "hfb_csr_spmm_dmma_ring": (i32, [i64, i64, vp, i32, i32, vp, i64, vp, i64, vp]),
        "hfb_csr_runs_measure": (i32, [i64, vp, vp, vp, vp, i64, vp]),
        "hfb_csr_runs_blob_stride": (i64, [i32, i32, i32]),
        "hfb_csr_pack_clusters_runs": (i32, [i64, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp]),
        "hfb_csr_spmm_runs_slots": (i32, [i64, i64, i32, i32, i32, i32]),
        "hfb_csr_spmm_runs": (i32, [i64, i64, vp, i32, i32, i32, i32, vp, i64, vp, i64, vp]),
        "hfb_csr_spmm_rows": (i32, [i64, i64, vp, vp, vp, vp, i64, vp, i64, vp]),
        "hfb_coldot_workspace_bytes": (sz, [i64, i64]),
        "hfb_coldot": (i32, [i64, i64, vp, i64, vp, i64, vp, vp, sz, vp]),
        "hfb_rowdot": (i32, [i64, i64, vp, i64, vp, i64, vp, vp]),
        "hfb_colscale": (i32, [i64, i64, vp, i64, vp, vp]),
        "hfb_colmean_workspace_bytes": (sz, [i64, i64]),
        "hfb_colsum": (i32, [i64, i64, vp, i64, dbl, vp, vp, sz, vp]),
        "hfb_colsum_weighted": (i32, [i64, i64, vp, i64, vp, dbl, vp, vp, sz, vp]),
        "hfb_subtract_row": (i32, [i64, i64, vp, i64, vp, vp]),
        "hfb_rank1_update": (i32, [i64, i64, dbl, vp, vp, vp, i64, vp]),
        "hfb_axpby": (i32, [i64, i64, dbl, vp, i64, dbl, vp, i64, vp]),
        "hfb_axpby_cols": (i32, [i64, i64, vp, vp, i64, vp, vp, i64, vp]),
        "hfb_rowscale": (i32, [i64, i64, vp, vp, i64, vp, i64, vp]),
        "hfb_measure_dmma_peak": (i32, [vp, sz, ctypes.POINTER(ctypes.c_double), vp]),
        "hfb_chol_inverse_workspace_bytes": (sz, [i64]),
        "hfb_chol_inverse": (i32, [i64, vp, i64, vp, i64, vp, i32, vp, sz, vp]),
        "hfb_chol_inverse_profile": (i32, [i64, vp, i64, vp, i64, vp, i32, vp, sz, vp, vp]),
        "hfb_jacobi_svd_max_elems": (i64, []),
        "hfb_jacobi_svd_batched": (i32, [i64, i64, vp, i64, i64, i64, vp, i64, vp, i32, i32, vp]),
        "hfb_fill_random": (i32, [i64, i64, vp, i64, u64, i64, i32, vp]),
        "hfb_host_copy": (i32, [vp, vp, sz, i32]),
        "hfb_peer_alloc": (i32, [sz, ctypes.POINTER(vp)]),
        "hfb_peer_free": (i32, [vp]),
        "hfb_peer_get_handle": (i32, [vp, ctypes.c_char_p]),
        "hfb_peer_open": (i32, [ctypes.c_char_p, ctypes.POINTER(vp)]),
        "hfb_peer_close": (i32, [vp]),
        "hfb_dgemm_peer": (i32, [i32, i64, i64, i64, dbl, vp, i64, vp, i64, ctypes.POINTER(vp), i32, i64, i64, vp]),
        "hfb_peer_barrier": (i32, [ctypes.POINTER(vp), i32, i32, u64, dbl, i32, vp]),
        "hfb_peer_reduce_bcast": (i32, [vp, i64, i32, i32, i64, i64, i64, ctypes.POINTER(vp), i64, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED = ["hfb_version", "hfb_launch_count", "hfb_dgemm_workspace_bytes", "hfb_dgemm_auto_splits", "hfb_dgemm", "hfb_dgemm_ex",
            "hfb_dgemm_ex_workspace_bytes",
            "hfb_dgemm_batched_workspace_bytes", "hfb_dgemm_batched",
            "hfb_dgemm_batched_small", "hfb_csr_spmm", "hfb_csr_spmm_ordered", 
            "hfb_csr_cluster_rows_capped", 
            "hfb_csr_cluster_blob_stride", "hfb_csr_pack_clusters", "hfb_csr_spmm_dmma",
            "hfb_csr_frag_blob_stride", "hfb_csr_pack_clusters_frag", "hfb_csr_spmm_dmma_frag", "hfb_csr_spmm_dmma_ring",
            "hfb_csr_runs_measure", "hfb_csr_runs_blob_stride", "hfb_csr_pack_clusters_runs", "hfb_csr_spmm_runs_slots",
            "hfb_csr_spmm_runs",
            "hfb_csr_spmm_rows", "hfb_coldot_workspace_bytes",
            "hfb_coldot", "hfb_rowdot", "hfb_colscale", "hfb_colmean_workspace_bytes", "hfb_colsum", "hfb_colsum_weighted", "hfb_subtract_row",
            "hfb_rank1_update", "hfb_axpby", "hfb_axpby_cols", "hfb_rowscale", "hfb_fill_random",
            "hfb_measure_dmma_peak", "hfb_chol_inverse_workspace_bytes", "hfb_chol_inverse", "hfb_chol_inverse_profile",
            "hfb_jacobi_svd_max_elems",
            "hfb_jacobi_svd_batched",
            "hfb_host_copy", "hfb_peer_alloc", "hfb_peer_free", "hfb_peer_get_handle", "hfb_peer_open", "hfb_peer_close", "hfb_dgemm_peer",
            "hfb_peer_barrier", "hfb_peer_reduce_bcast"]


def _check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise HfbError("%s: %s (code %d)" % (what, _ERRORS.get(rc, "error"), rc))
    raise HfbError("%s: CUDA error code %d" % (what, rc))


def _stream():
    """Current stream of the current device.  One process per GPU is the supported setup; ``_req`` refuses operands
    that live on another device than the current one (the launch would go to the wrong stream, and the C side keeps
    per-process caches of the SM count and kernel attributes)."""
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def is_device_tensor(t):
    """True for a tensor that lives in device memory (the only place the kernels can read)."""
    return isinstance(t, torch.Tensor) and t.is_cuda


def _req(t, name):
    if not (is_device_tensor(t) and t.dtype == torch.float64 and t.dim() == 2 and t.stride(1) == 1):
        raise HfbError("%s must be a 2-D float64 CUDA tensor with unit inner stride" % name)
    if t.is_cuda and t.device.index != torch.cuda.current_device():
        raise HfbError("%s lives on cuda:%d but the current device is cuda:%d: wrap the call in torch.cuda.device(...)"
                       % (name, t.device.index, torch.cuda.current_device()))
    return t


def _ld(t):
    # leading dimension of a 2-D row-major view (a single row may report any stride)
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


# workspace cache: one growable byte buffer per device (caller-owned memory from the C ABI's viewpoint)
_ws = {}


def workspace(nbytes, device):
    key = (device.type, device.index)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def padded_empty(rows, cols, device, pad=16):
    """(rows, cols) float64 view of a buffer whose leading dimension is a multiple of ``pad`` elements
    (TMA needs an even ld; 16 keeps every row 128-byte aligned)."""
    ld = ((cols + pad - 1) // pad) * pad
    return torch.empty((rows, ld), dtype=torch.float64, device=device)[:, :cols]


def padded_zeros(rows, cols, device, pad=16):
    ld = ((cols + pad - 1) // pad) * pad
    return torch.zeros((rows, ld), dtype=torch.float64, device=device)[:, :cols]


def to_padded(a, device, pad=16):
    """Copy a host/device (rows, cols) array into a padded-ld device buffer."""
    t = torch.as_tensor(a)
    out = padded_empty(t.shape[0], t.shape[1], device, pad)
    out.copy_(t.to(dtype=torch.float64), non_blocking=True)
    return out


def dgemm(layout, A, B, out=None, alpha=1.0, splits=0, symmetric=False, accumulate=False, b_upper=False):
    """out[M,N] = alpha * op(A) op(B) on the DMMA/TMA kernel.  layout: HFB_NN (A MxK, B KxN),
    HFB_TN (A KxM), HFB_NT (B NxK).  symmetric=True: the caller asserts a symmetric result (Gram matrix); only the
    tiles on or above the diagonal are computed, the rest is mirrored."""
    L = lib()
    _req(A, "A"), _req(B, "B")
    if layout == HFB_NN:
        M, K = A.shape
        K2, N = B.shape
    elif layout == HFB_TN:
        K, M = A.shape
        K2, N = B.shape
    elif layout == HFB_NT:
        M, K = A.shape
        N, K2 = B.shape
    else:
        raise HfbError("unknown layout")
    if K != K2:
        raise HfbError("dgemm: inner dimensions differ (%d vs %d)" % (K, K2))
    fresh_out = out is None
    if out is None:
        out = padded_empty(M, N, A.device)
    _req(out, "out")
    if tuple(out.shape) != (M, N):
        raise HfbError("dgemm: out has shape %s, expected %s" % (tuple(out.shape), (M, N)))
    flags = 1 if (symmetric and M == N) else 0
    if accumulate:
        if fresh_out or flags:
            raise HfbError("dgemm: accumulate needs an existing out and is not combinable with symmetric")
        flags |= 2
    if b_upper:                     # B is the upper-triangular factor S of Cholesky-QR: skip the zero k-blocks (TRMM)
        if K != N or layout == HFB_NT:
            raise HfbError("dgemm: b_upper needs a square K x N operand B (NN / TN layouts)")
        flags |= 4
        splits = 1
    nbytes = L.hfb_dgemm_ex_workspace_bytes(layout, M, N, K, splits, flags)
    ws = workspace(nbytes, A.device) if nbytes else None
    if TIMING is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = L.hfb_dgemm_ex(layout, M, N, K, float(alpha), A.data_ptr(), _ld(A), B.data_ptr(), _ld(B), out.data_ptr(), _ld(out),
                        ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0, int(splits),
                        flags, _stream())
    if TIMING is not None:
        e1.record()
        TIMING.append(((layout, M, N, K), e0, e1))
    _check(rc, "hfb_dgemm")
    return out


def _batch_operand(t, name):
    """(rows, cols, ld, batch stride, batch) of a 2-D (shared) or 3-D (batch, rows, cols) float64 CUDA operand."""
    if not (is_device_tensor(t) and t.dtype == torch.float64 and t.dim() in (2, 3) and t.stride(-1) == 1):
        raise HfbError("%s must be a 2-D or 3-D float64 CUDA tensor with unit inner stride" % name)
    if t.dim() == 2:
        return t.shape[0], t.shape[1], _ld(t), 0, 1
    ld = t.stride(1) if t.shape[1] > 1 else max(t.stride(1), t.shape[2])
    return t.shape[1], t.shape[2], ld, (t.stride(0) if t.shape[0] > 1 else 0), t.shape[0]


def batched_empty(batch, rows, cols, device, pad=16):
    """(batch, rows, cols) float64 view whose row pitch is a multiple of ``pad`` elements (TMA-conforming batch operand)."""
    ld = ((cols + pad - 1) // pad) * pad
    return torch.empty((batch, rows, ld), dtype=torch.float64, device=device)[:, :, :cols]


def dgemm_batched(layout, A, B, out=None, alpha=1.0, reduce=False, accumulate=False):
    """Strided-batch DMMA GEMM over the sample axis (hfb_dgemm_batched).  A, B: (batch, rows, cols) stacks, or 2-D =
    shared by every sample.  reduce=False: out (batch, M, N), out[b] = alpha op(A[b]) op(B[b]) in ONE launch;
    reduce=True: out (M, N) = alpha * sum_b op(A[b]) op(B[b]) (the K loop runs over the samples, deterministic split-K)."""
    L = lib()
    ra, ca, lda, sa, ba = _batch_operand(A, "A")
    rb, cb, ldb, sb, bb = _batch_operand(B, "B")
    batch = max(ba, bb)
    if (ba not in (1, batch)) or (bb not in (1, batch)):
        raise HfbError("dgemm_batched: batch sizes differ (%d vs %d)" % (ba, bb))
    if layout == HFB_NN:
        M, K, K2, N = ra, ca, rb, cb
    elif layout == HFB_TN:
        K, M, K2, N = ra, ca, rb, cb
    elif layout == HFB_NT:
        M, K, N, K2 = ra, ca, rb, cb
    else:
        raise HfbError("unknown layout")
    if K != K2:
        raise HfbError("dgemm_batched: inner dimensions differ (%d vs %d)" % (K, K2))
    dev = A.device
    if reduce:
        if out is None:
            if accumulate:
                raise HfbError("dgemm_batched: accumulate needs an existing out")
            out = padded_empty(M, N, dev)
        _req(out, "out")
        if tuple(out.shape) != (M, N):
            raise HfbError("dgemm_batched: out has shape %s, expected %s" % (tuple(out.shape), (M, N)))
        ldc, sc = _ld(out), 0
    else:
        if out is None:
            if accumulate:
                raise HfbError("dgemm_batched: accumulate needs an existing out")
            out = batched_empty(batch, M, N, dev)
        if not (is_device_tensor(out) and out.dtype == torch.float64 and out.dim() == 3 and out.stride(2) == 1) or \
                tuple(out.shape) != (batch, M, N):
            raise HfbError("dgemm_batched: out must be a (batch, M, N) float64 CUDA tensor with unit inner stride")
        ldc = out.stride(1) if M > 1 else max(out.stride(1), N)
        sc = out.stride(0) if batch > 1 else ldc * M
    mode = 1 if reduce else 0
    nbytes = L.hfb_dgemm_batched_workspace_bytes(layout, M, N, K, batch, mode)
    ws = workspace(nbytes, dev) if nbytes else None
    if TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = L.hfb_dgemm_batched(layout, M, N, K, float(alpha), A.data_ptr(), lda, sa, B.data_ptr(), ldb, sb, out.data_ptr(), ldc, sc,
                             batch, mode, 2 if accumulate else 0, ws.data_ptr() if ws is not None else None,
                             ws.numel() if ws is not None else 0, _stream())
    if TIMING is not None:
        e1.record()
        TIMING.append((("batched", layout, M, N, K, batch, mode), e0, e1))
    _check(rc, "hfb_dgemm_batched")
    return out


def dgemm_batched_small(layout, A, B, out, alpha=1.0):
    """out[b] = alpha * op(A[b]) @ B[b]; A/B/out are 3-D (batch, rows, cols) float64 CUDA tensors with unit
    inner stride; a batch stride of 0 (expanded tensor) shares the operand."""
    L = lib()
    batch = out.shape[0]
    if layout == HFB_TN:
        K, M = A.shape[1], A.shape[2]
    else:
        M, K = A.shape[1], A.shape[2]
    N = B.shape[2]
    for t in (A, B, out):
        if not (is_device_tensor(t) and t.dtype == torch.float64 and t.dim() == 3 and t.stride(2) == 1):
            raise HfbError("dgemm_batched_small: operands must be 3-D float64 CUDA tensors")
    if B.shape[1] != K or tuple(out.shape[1:]) != (M, N):
        raise HfbError("dgemm_batched_small: shape mismatch")
    rc = L.hfb_dgemm_batched_small(layout, M, N, K, float(alpha), A.data_ptr(), A.stride(1), A.stride(0) if A.shape[0] > 1 else 0,
                                   B.data_ptr(), B.stride(1), B.stride(0) if B.shape[0] > 1 else 0, out.data_ptr(),
                                   out.stride(1), out.stride(0), batch, _stream())
    _check(rc, "hfb_dgemm_batched_small")
    return out


def csr_spmm(rowptr, colind, val, B, out=None, order=None):
    L = lib()
    _req(B, "B")
    n, m = B.shape
    if out is None:
        out = padded_empty(n, m, B.device)
    nrows = rowptr.numel() - 1
    if order is not None:
        rc = L.hfb_csr_spmm_ordered(nrows, m, rowptr.data_ptr(), colind.data_ptr(), val.data_ptr(), order.data_ptr(),
                                    B.data_ptr(), _ld(B), out.data_ptr(), _ld(out), _stream())
    else:
        rc = L.hfb_csr_spmm(nrows, m, rowptr.data_ptr(), colind.data_ptr(), val.data_ptr(), B.data_ptr(), _ld(B),
                            out.data_ptr(), _ld(out), _stream())
    _check(rc, "hfb_csr_spmm")
    return out


def csr_cluster_rows_capped(indptr, indices, max_rows=64, max_cols=128):
    """Host preprocessing for the cluster SpMM kernels: (order, cluster_ptr) as NumPy int32 arrays."""
    import numpy as np
    L = lib()
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    n = indptr.size - 1
    order = np.empty(n, dtype=np.int32)
    cptr = np.empty(n + 1, dtype=np.int32)
    ncl = ctypes.c_int64(0)
    rc = L.hfb_csr_cluster_rows_capped(n, indptr.ctypes.data, indices.ctypes.data, int(max_rows), int(max_cols),
                                       order.ctypes.data, cptr.ctypes.data, ctypes.byref(ncl))
    _check(rc, "hfb_csr_cluster_rows_capped")
    return order, cptr[:ncl.value + 1].copy()


def csr_pack_clusters(indptr, indices, data, order, cptr, max_rows, max_cols, max_entries):
    """Host preprocessing for the panel DMMA SpMM: uint8 NumPy buffer of per-cluster records (hfb_csr_pack_clusters)."""
    import numpy as np
    L = lib()
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float64)
    order = np.ascontiguousarray(order, dtype=np.int32)
    cptr = np.ascontiguousarray(cptr, dtype=np.int32)
    ncl = cptr.size - 1
    stride = int(L.hfb_csr_cluster_blob_stride(int(max_rows), int(max_cols), int(max_entries)))
    if stride <= 0:
        raise HfbError("hfb_csr_cluster_blob_stride: invalid caps")
    blobs = np.empty(ncl * stride, dtype=np.uint8)
    rc = L.hfb_csr_pack_clusters(indptr.size - 1, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, order.ctypes.data,
                                 cptr.ctypes.data, ncl, int(max_rows), int(max_cols), int(max_entries), blobs.ctypes.data)
    _check(rc, "hfb_csr_pack_clusters")
    return blobs


def csr_spmm_dmma(plan, B, out=None):
    """C = M @ B with the cluster-dense DMMA kernel (A fragments of the per-cluster dense block in registers, B rows staged
    by cp.async); ``plan`` is the dict built by linalg.CsrMatrix._build_plan with its blobs packed."""
    L = lib()
    _req(B, "B")
    n, m = B.shape
    if out is None:
        out = padded_empty(n, m, B.device)
    rc = L.hfb_csr_spmm_dmma(plan["nclusters"], m, plan["blobs"].data_ptr(), plan["max_rows"], plan["max_cols_cap"],
                             plan["max_entries"], B.data_ptr(), _ld(B), out.data_ptr(), _ld(out), _stream())
    _check(rc, "hfb_csr_spmm_dmma")
    return out


def csr_spmm_dmma_ring(plan, B, out=None):
    """C = M @ B with the ring-pipelined fragment-record kernel (hfb_csr_spmm_dmma_ring); same records and results as
    ``csr_spmm_dmma_frag``.  Needs clusters of 9..16 rows and m <= 384."""
    L = lib()
    _req(B, "B")
    n, m = B.shape
    if out is None:
        out = padded_empty(n, m, B.device)
    rc = L.hfb_csr_spmm_dmma_ring(plan["nclusters"], m, plan["fblobs"].data_ptr(), plan["max_rows"], plan["max_cols_cap"],
                                  B.data_ptr(), _ld(B), out.data_ptr(), _ld(out), _stream())
    _check(rc, "hfb_csr_spmm_dmma_ring")
    return out


def csr_runs_measure(indptr, indices, order, cptr):
    """Caps of a cluster plan (hfb_csr_runs_measure, host): the largest cluster's rows, runs of consecutive columns, distinct
    columns (= staged B rows) and entries."""
    import numpy as np
    L = lib()
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    order = np.ascontiguousarray(order, dtype=np.int32)
    cptr = np.ascontiguousarray(cptr, dtype=np.int32)
    caps = np.zeros(4, dtype=np.int32)
    rc = L.hfb_csr_runs_measure(indptr.size - 1, indptr.ctypes.data, indices.ctypes.data, order.ctypes.data, cptr.ctypes.data,
                                cptr.size - 1, caps.ctypes.data)
    _check(rc, "hfb_csr_runs_measure")
    return {"max_rows": int(caps[0]), "max_runs": int(caps[1]), "max_brow": int(caps[2]), "max_entries": int(caps[3])}


def csr_pack_clusters_runs(indptr, indices, data, order, cptr):
    """Host preprocessing for the run-staged FMA SpMM (hfb_csr_pack_clusters_runs): returns (uint8 NumPy buffer of per-cluster
    run records, caps dict) or raises HfbError when the plan does not fit the kernel (> 16 rows or > 32 runs per cluster)."""
    import numpy as np
    L = lib()
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float64)
    order = np.ascontiguousarray(order, dtype=np.int32)
    cptr = np.ascontiguousarray(cptr, dtype=np.int32)
    ncl = cptr.size - 1
    caps = csr_runs_measure(indptr, indices, order, cptr)
    max_rows, max_runs, max_entries = caps["max_rows"], caps["max_runs"], caps["max_entries"]
    stride = int(L.hfb_csr_runs_blob_stride(max_rows, max_runs, max_entries))
    if stride <= 0:
        raise HfbError("hfb_csr_runs_blob_stride: unsupported cluster plan (rows %d, runs %d per cluster; limits 16 / 32)"
                       % (max_rows, max_runs))
    blobs = np.empty(ncl * stride, dtype=np.uint8)
    rc = L.hfb_csr_pack_clusters_runs(indptr.size - 1, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data,
                                      order.ctypes.data, cptr.ctypes.data, ncl, max_rows, max_runs, max_entries,
                                      blobs.ctypes.data)
    _check(rc, "hfb_csr_pack_clusters_runs")
    caps["stride"] = stride
    return blobs, caps


def csr_spmm_runs_slots(rplan, m, ldb):
    """Ring slots the run-staged kernel would use for an (n, m) block of pitch ldb; 0 = unsupported shape."""
    return int(lib().hfb_csr_spmm_runs_slots(int(m), int(ldb), rplan["max_rows"], rplan["max_runs"], rplan["max_brow"],
                                             rplan["max_entries"]))


def csr_spmm_runs(rplan, B, out=None):
    """C = M @ B with the run-staged FMA kernel (hfb_csr_spmm_runs); ``rplan`` = caps of csr_pack_clusters_runs plus
    ``nclusters`` and the device tensor ``blobs`` (linalg.CsrMatrix._runs_blobs)."""
    L = lib()
    _req(B, "B")
    n, m = B.shape
    if out is None:
        out = padded_empty(n, m, B.device)
    rc = L.hfb_csr_spmm_runs(rplan["nclusters"], m, rplan["blobs"].data_ptr(), rplan["max_rows"], rplan["max_runs"],
                             rplan["max_brow"], rplan["max_entries"], B.data_ptr(), _ld(B), out.data_ptr(), _ld(out), _stream())
    _check(rc, "hfb_csr_spmm_runs")
    return out


def csr_pack_clusters_frag(indptr, indices, data, order, cptr, max_rows, max_cols):
    """Host preprocessing for the fragment-blob DMMA SpMM: uint8 NumPy buffer of per-cluster records holding the dense
    cluster block in DMMA A-fragment order (hfb_csr_pack_clusters_frag)."""
    import numpy as np
    L = lib()
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float64)
    order = np.ascontiguousarray(order, dtype=np.int32)
    cptr = np.ascontiguousarray(cptr, dtype=np.int32)
    ncl = cptr.size - 1
    stride = int(L.hfb_csr_frag_blob_stride(int(max_rows), int(max_cols)))
    if stride <= 0:
        raise HfbError("hfb_csr_frag_blob_stride: cluster caps (%d, %d) exceed the DMMA kernel's (16, 48)" % (max_rows, max_cols))
    blobs = np.empty(ncl * stride, dtype=np.uint8)
    rc = L.hfb_csr_pack_clusters_frag(indptr.size - 1, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data,
                                      order.ctypes.data, cptr.ctypes.data, ncl, int(max_rows), int(max_cols), blobs.ctypes.data)
    _check(rc, "hfb_csr_pack_clusters_frag")
    return blobs


def csr_spmm_dmma_frag(plan, B, out=None, chunk_cols=0):
    """C = M @ B with the fragment-record DMMA kernel; ``plan`` carries ``fblobs`` (linalg.CsrMatrix._frag_blobs).
    ``chunk_cols`` = columns staged per CTA (0: whole rows up to 320 columns)."""
    L = lib()
    _req(B, "B")
    n, m = B.shape
    if out is None:
        out = padded_empty(n, m, B.device)
    rc = L.hfb_csr_spmm_dmma_frag(plan["nclusters"], m, plan["fblobs"].data_ptr(), plan["max_rows"], plan["max_cols_cap"],
            int(chunk_cols), B.data_ptr(), _ld(B), out.data_ptr(), _ld(out), _stream())
    _check(rc, "hfb_csr_spmm_dmma_frag")
    return out


def csr_spmm_rows(rowptr, colind, val, X, out=None):
    L = lib()
    _req(X, "X")
    N, n = X.shape
    if out is None:
        out = padded_empty(N, n, X.device)
    rc = L.hfb_csr_spmm_rows(N, n, rowptr.data_ptr(), colind.data_ptr(), val.data_ptr(), X.data_ptr(), _ld(X),
                             out.data_ptr(), _ld(out), _stream())
    _check(rc, "hfb_csr_spmm_rows")
    return out


def coldot(X, Y):
    L = lib()
    _req(X, "X"), _req(Y, "Y")
    n, m = X.shape
    out = torch.empty(m, dtype=torch.float64, device=X.device)
    nbytes = L.hfb_coldot_workspace_bytes(n, m)
    ws = workspace(nbytes, X.device)
    rc = L.hfb_coldot(n, m, X.data_ptr(), _ld(X), Y.data_ptr(), _ld(Y), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _check(rc, "hfb_coldot")
    return out


def rowdot(X, Y):
    """out[i] = <X[i, :], Y[i, :]>."""
    L = lib()
    _req(X, "X"), _req(Y, "Y")
    out = torch.empty(X.shape[0], dtype=torch.float64, device=X.device)
    rc = L.hfb_rowdot(X.shape[0], X.shape[1], X.data_ptr(), _ld(X), Y.data_ptr(), _ld(Y), out.data_ptr(), _stream())
    _check(rc, "hfb_rowdot")
    return out


def colscale_(X, s):
    L = lib()
    _req(X, "X")
    rc = L.hfb_colscale(X.shape[0], X.shape[1], X.data_ptr(), _ld(X), s.data_ptr(), _stream())
    _check(rc, "hfb_colscale")
    return X


def colsum(X, scale=1.0, weights=None):
    """scale * sum over rows (samples) of X -> vector of length X.shape[1]; with ``weights`` (device vector of length
    X.shape[0]) the rows are weighted: scale * weights^T X."""
    L = lib()
    _req(X, "X")
    N, n = X.shape
    out = torch.empty(n, dtype=torch.float64, device=X.device)
    nbytes = L.hfb_colmean_workspace_bytes(N, n)
    ws = workspace(nbytes, X.device)
    if weights is not None:
        if weights.numel() != N or not weights.is_contiguous():
            raise HfbError("colsum: weights must be a contiguous vector with one entry per row")
        rc = L.hfb_colsum_weighted(N, n, X.data_ptr(), _ld(X), weights.data_ptr(), float(scale), out.data_ptr(), ws.data_ptr(),
                                   ws.numel(), _stream())
    else:
        rc = L.hfb_colsum(N, n, X.data_ptr(), _ld(X), float(scale), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _check(rc, "hfb_colsum")
    return out


def subtract_row_(X, shift):
    L = lib()
    _req(X, "X")
    rc = L.hfb_subtract_row(X.shape[0], X.shape[1], X.data_ptr(), _ld(X), shift.data_ptr(), _stream())
    _check(rc, "hfb_subtract_row")
    return X


def rank1_update_(Y, a, x, y):
    """Y += a * outer(x, y) for device vectors x (rows) and y (columns)."""
    L = lib()
    _req(Y, "Y")
    if x.numel() != Y.shape[0] or y.numel() != Y.shape[1]:
        raise HfbError("rank1_update_: vector lengths do not match Y")
    rc = L.hfb_rank1_update(Y.shape[0], Y.shape[1], float(a), x.data_ptr(), y.data_ptr(), Y.data_ptr(), _ld(Y), _stream())
    _check(rc, "hfb_rank1_update")
    return Y


def axpby_(a, X, b, Y):
    """Y = a*X + b*Y."""
    L = lib()
    _req(X, "X"), _req(Y, "Y")
    rc = L.hfb_axpby(X.shape[0], X.shape[1], float(a), X.data_ptr(), _ld(X), float(b), Y.data_ptr(), _ld(Y), _stream())
    _check(rc, "hfb_axpby")
    return Y


def axpby_cols_(a, X, b, Y):
    """Y[:, j] = a[j] X[:, j] + b[j] Y[:, j]; a / b are device vectors or None (= 1)."""
    L = lib()
    _req(X, "X"), _req(Y, "Y")
    rc = L.hfb_axpby_cols(X.shape[0], X.shape[1], a.data_ptr() if a is not None else None, X.data_ptr(), _ld(X),
                          b.data_ptr() if b is not None else None, Y.data_ptr(), _ld(Y), _stream())
    _check(rc, "hfb_axpby_cols")
    return Y


def rowscale(s, X, out=None):
    L = lib()
    _req(X, "X")
    if out is None:
        out = padded_empty(X.shape[0], X.shape[1], X.device)
    rc = L.hfb_rowscale(X.shape[0], X.shape[1], s.data_ptr(), X.data_ptr(), _ld(X), out.data_ptr(), _ld(out), _stream())
    _check(rc, "hfb_rowscale")
    return out


def fill_random_(X, seed, row_offset=0, kind="normal"):
    L = lib()
    _req(X, "X")
    rc = L.hfb_fill_random(X.shape[0], X.shape[1], X.data_ptr(), _ld(X), int(seed), int(row_offset),
                           0 if kind == "normal" else 1, _stream())
    _check(rc, "hfb_fill_random")
    return X


def measure_dmma_peak(device):
    """FP64 DMMA ceiling of this GPU in TFLOP/s (hfb_measure_dmma_peak)."""
    L = lib()
    ws = workspace(1 << 20, device)
    out = ctypes.c_double(0.0)
    rc = L.hfb_measure_dmma_peak(ws.data_ptr(), ws.numel(), ctypes.byref(out), _stream())
    _check(rc, "hfb_measure_dmma_peak")
    return out.value


CHOL_INVERSE_MAX = 1024


def chol_inverse_profile(G, scale_columns=True):
    """Debug aid: per-section clock cycles of one hfb_chol_inverse launch (NumPy int64[11], see include/hfb200.h)."""
    L = lib()
    _req(G, "G")
    m = G.shape[0]
    S = padded_empty(m, m, G.device)
    stat = torch.empty(8, dtype=torch.float64, device=G.device)
    prof = torch.zeros(16, dtype=torch.int64, device=G.device)
    ws = torch.empty(L.hfb_chol_inverse_workspace_bytes(m), dtype=torch.uint8, device=G.device)
    rc = L.hfb_chol_inverse_profile(m, G.data_ptr(), _ld(G), S.data_ptr(), _ld(S), stat.data_ptr(), 1 if scale_columns else 0,
                                    ws.data_ptr(), ws.numel(), prof.data_ptr(), _stream())
    _check(rc, "hfb_chol_inverse_profile")
    return prof.cpu().numpy()[:11]


def chol_inverse(G, scale_columns=True):
    """Device Cholesky-QR factor of the (m x m) Gram matrix G (hfb_chol_inverse): returns (S, stat) with S = D^-1 R^-1
    (upper triangular, padded device block) and stat a DEVICE vector of 8 doubles {fail, shift, cond, attempts,
    max |Gs - I|, max |d - 1|, zero columns, m}.  Asynchronous: nothing is copied to the host."""
    L = lib()
    _req(G, "G")
    m = G.shape[0]
    if G.shape[1] != m or m > CHOL_INVERSE_MAX:
        raise HfbError("chol_inverse: G must be square with m <= %d" % CHOL_INVERSE_MAX)
    S = padded_empty(m, m, G.device)
    stat = torch.empty(8, dtype=torch.float64, device=G.device)
    nbytes = L.hfb_chol_inverse_workspace_bytes(m)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=G.device)      # own scratch: the shared workspace may be in use by a GEMM queued behind
    rc = L.hfb_chol_inverse(m, G.data_ptr(), _ld(G), S.data_ptr(), _ld(S), stat.data_ptr(), 1 if scale_columns else 0,
                            ws.data_ptr(), ws.numel(), _stream())
    _check(rc, "hfb_chol_inverse")
    return S, stat


def jacobi_svd_fits(rows, cols):
    """True when a (rows x cols) block fits the shared-memory resident batched Jacobi SVD."""
    return cols * (rows | 1) + cols + 2 + (cols + 1) // 2 <= int(lib().hfb_jacobi_svd_max_elems())


def jacobi_svd_batched_(A, keep_scaled=False, max_sweeps=40):
    """In-place batched one-sided Jacobi SVD (hfb_jacobi_svd_batched): A (batch, rows, cols) -> U (unit columns sorted by
    descending singular value; ``keep_scaled``: U diag(sigma)).  Returns (sigma (batch, cols), info (batch,) int32 device)."""
    L = lib()
    if not (is_device_tensor(A) and A.dtype == torch.float64 and A.dim() == 3 and A.stride(2) == 1):
        raise HfbError("jacobi_svd_batched_: A must be a (batch, rows, cols) float64 CUDA tensor with unit inner stride")
    batch, rows, cols = A.shape
    sigma = torch.empty((batch, cols), dtype=torch.float64, device=A.device)
    info = torch.empty(batch, dtype=torch.int32, device=A.device)
    lda = A.stride(1) if rows > 1 else max(A.stride(1), cols)
    rc = L.hfb_jacobi_svd_batched(rows, cols, A.data_ptr(), lda, A.stride(0) if batch > 1 else lda * rows, batch, sigma.data_ptr(),
                                  cols, info.data_ptr(), int(max_sweeps), 1 if keep_scaled else 0, _stream())
    _check(rc, "hfb_jacobi_svd_batched")
    return sigma, info


# optional per-call CUDA-event timing of the GEMM launches (bench.py's roofline leg)
TIMING = None


def start_timing():
    global TIMING
    TIMING = []


def stop_timing():
    """Returns {tag: (calls, total_ms)}; tags are (layout, M, N, K)."""
    global TIMING
    rec, TIMING = TIMING, None
    torch.cuda.synchronize()
    out = {}
    for tag, e0, e1 in rec or []:
        c, t = out.get(tag, (0, 0.0))
        out[tag] = (c + 1, t + e0.elapsed_time(e1))
    return out


def launch_count():
    return int(lib().hfb_launch_count())


def host_copy_(dst, src, nthreads):
    """dst[...] = src for two contiguous HOST tensors of equal size and dtype, on ``nthreads`` threads with non-temporal
    stores (hfb_host_copy): the pageable -> pinned leg of the staged upload."""
    if dst.is_cuda or src.is_cuda or dst.dtype != src.dtype or dst.numel() != src.numel() or \
            not dst.is_contiguous() or not src.is_contiguous():
        raise HfbError("host_copy_: two contiguous host tensors of equal size and dtype expected")
    _check(lib().hfb_host_copy(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src.data_ptr()),
                               dst.numel() * dst.element_size(), int(nthreads)), "hfb_host_copy")
    return dst


# ---------------------------------------------------------------------------------------------- peer exchange (NVLink)
PEER_HANDLE_BYTES = 64
PEER_MAX_RANKS = 16


def peer_alloc(nbytes):
    """Zeroed device buffer that can be shared with the other ranks of the node through CUDA IPC (address as int)."""
    p = ctypes.c_void_p()
    _check(lib().hfb_peer_alloc(int(nbytes), ctypes.byref(p)), "hfb_peer_alloc")
    return int(p.value)


def peer_free(ptr):
    _check(lib().hfb_peer_free(ctypes.c_void_p(ptr)), "hfb_peer_free")


def peer_get_handle(ptr):
    buf = ctypes.create_string_buffer(PEER_HANDLE_BYTES)
    _check(lib().hfb_peer_get_handle(ctypes.c_void_p(ptr), buf), "hfb_peer_get_handle")
    return bytes(buf.raw)


def peer_open(handle):
    if len(handle) != PEER_HANDLE_BYTES:
        raise HfbError("peer_open: a CUDA IPC handle has %d bytes" % PEER_HANDLE_BYTES)
    p = ctypes.c_void_p()
    _check(lib().hfb_peer_open(ctypes.c_char_p(handle), ctypes.byref(p)), "hfb_peer_open")
    return int(p.value)


def peer_close(ptr):
    _check(lib().hfb_peer_close(ctypes.c_void_p(ptr)), "hfb_peer_close")


def _ptr_array(ptrs):
    return (ctypes.c_void_p * len(ptrs))(*[ctypes.c_void_p(int(p)) for p in ptrs])


def dgemm_peer(A, B, slot_ptrs, block_rows, ld_slot, alpha=1.0):
    """Fused lift + reduce-scatter: rows [o * block_rows, (o+1) * block_rows) of alpha * A^T B (TN layout, A: K x M) are
    stored by the GEMM epilogue into ``slot_ptrs[o]`` (this rank's slot in rank o's exchange buffer, leading dimension
    ``ld_slot``), over NVLink for o != this rank."""
    _req(A, "A"), _req(B, "B")
    Kd, M = A.shape
    K2, N = B.shape
    if Kd != K2:
        raise HfbError("dgemm_peer: inner dimensions differ (%d vs %d)" % (Kd, K2))
    if TIMING is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = lib().hfb_dgemm_peer(HFB_TN, M, N, Kd, float(alpha), A.data_ptr(), _ld(A), B.data_ptr(), _ld(B),
                              _ptr_array(slot_ptrs), len(slot_ptrs), int(block_rows), int(ld_slot), _stream())
    if TIMING is not None:
        e1.record()
        TIMING.append(((HFB_TN, M, N, Kd), e0, e1))
    _check(rc, "hfb_dgemm_peer")


PEER_SIGNAL, PEER_WAIT = 1, 2


def peer_barrier(flag_ptrs, me, epoch, timeout_s=600.0, mode=PEER_SIGNAL | PEER_WAIT):
    _check(lib().hfb_peer_barrier(_ptr_array(flag_ptrs), int(me), len(flag_ptrs), int(epoch), float(timeout_s), int(mode),
                                  _stream()), "hfb_peer_barrier")


def peer_reduce_bcast(slots_ptr, slot_stride, me, rows, cols, ld, y_ptrs, ldy):
    _check(lib().hfb_peer_reduce_bcast(ctypes.c_void_p(slots_ptr), int(slot_stride), len(y_ptrs), int(me), int(rows), int(cols),
                                       int(ld), _ptr_array(y_ptrs), int(ldy), _stream()), "hfb_peer_reduce_bcast")


class _DeviceBlock:
    """Raw device memory described through ``__cuda_array_interface__`` so that torch can view it without a copy."""

    def __init__(self, ptr, rows, cols, ld, owner):
        self.owner = owner                  # keeps the allocation alive as long as a tensor views it
        self.__cuda_array_interface__ = {"shape": (int(rows), int(cols)), "typestr": "<f8", "data": (int(ptr), False),
                                         "strides": (int(ld) * 8, 8), "version": 2}


def tensor_from_ptr(ptr, rows, cols, ld, device, owner=None):
    """(rows, cols) float64 view with leading dimension ``ld`` of device memory at ``ptr`` (not owned by torch)."""
    t = torch.as_tensor(_DeviceBlock(ptr, rows, cols, ld, owner), device=device)
    if t.data_ptr() != int(ptr) or t.stride(0) != int(ld):
        raise HfbError("tensor_from_ptr: torch copied the block instead of viewing it")
    return t
