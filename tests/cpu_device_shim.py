"""TEST DOUBLE (test infrastructure only -- nothing under hippyflow_b200/ imports this file).

The product has no CPU path: every numerical call goes to libhfb200.so on a B200.  To exercise the HOST logic of the
projector classes in the CPU test suite (which route is taken, the order of the products, the rank-one corrections of
the mean shift, the deferred clean-up pass, chunk bookkeeping, collectives) this module swaps the thin wrappers of
``hippyflow_b200._lib`` for plain torch/NumPy fp64 evaluations of the same operations and stubs the CUDA stream /
event API with no-ops.  Numbers produced under the shim say nothing about the kernels; the `-m gpu` tests check those.
"""
import contextlib

import numpy as np
import scipy.sparse as sp
import torch


class _FakeStream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def wait_stream(self, s):
        pass

    def wait_event(self, e):
        pass

    def synchronize(self):
        pass


class _FakeEvent:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return 0.0


def _scipy_csr(rowptr, colind, val, ncols=None):
    rp, ci, v = rowptr.numpy(), colind.numpy(), val.numpy()
    n = rp.size - 1
    return sp.csr_matrix((v, ci, rp), shape=(n, ncols if ncols is not None else n))


def _decode_panel_blobs(K, plan):
    """Rebuild the matrix the way csr_spmm_dmma_kernel reads its records: (local column, local ROW) fields of every
    entry scattered into a dense per-cluster block -- the row offsets are not consulted."""
    mr, mc, me = plan["max_rows"], plan["max_cols_cap"], plan["max_entries"]
    stride = int(K.lib().hfb_csr_cluster_blob_stride(mr, mc, me))
    blobs = plan["blobs"].numpy().reshape(-1, stride)
    r4 = lambda x: (x + 3) // 4 * 4
    off_outrow = 16 + 4 * r4(mr + 1)
    off_cols = off_outrow + 4 * r4(mr)
    off_ent = off_cols + 4 * r4(mc)
    rows, cols, vals = [], [], []
    for b in blobs:
        nrow, ncol, nent = (int(x) for x in b[:12].view(np.int32))
        orow = b[off_outrow:off_outrow + 4 * mr].view(np.int32)
        cl = b[off_cols:off_cols + 4 * mc].view(np.int32)
        e = b[off_ent:off_ent + 16 * nent].reshape(-1, 16)
        v = e[:, :8].copy().view(np.float64).ravel()
        lr = e[:, 8:16].copy().view(np.int32).reshape(-1, 2)
        D = np.zeros((mc, mr))
        D[lr[:, 0], lr[:, 1]] = v
        jj, rr = np.nonzero(D[:ncol, :nrow])
        rows.append(orow[rr]); cols.append(cl[jj]); vals.append(D[jj, rr])
    n = plan["order"].numel()
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


def decode_frag_blobs(K, fblobs, max_rows, max_cols, n):
    """Rebuild the matrix from the fragment records of csr_spmm_dmma_frag_kernel, reading them the way the kernel does:
    a[ks] of lane (g = lane >> 2, t = lane & 3) of row half h is D[g + 8h][4 ks + t]; blocks whose mask bit is clear are
    taken as zero WITHOUT looking at the stored values."""
    rh = 1 if max_rows <= 8 else 2
    maxks = 4 if max_cols <= 16 else 6 if max_cols <= 24 else 8 if max_cols <= 32 else 12
    stride = int(K.lib().hfb_csr_frag_blob_stride(int(max_rows), int(max_cols)))
    off_outrow = 32
    off_cols = off_outrow + 4 * 8 * rh
    off_afrag = (off_cols + 4 * 4 * maxks + 255) // 256 * 256
    assert stride == (off_afrag + 8 * 32 * rh * maxks + 255) // 256 * 256
    blobs = np.asarray(fblobs).reshape(-1, stride)
    rows, cols, vals = [], [], []
    lane = np.arange(32)
    g, t = lane >> 2, lane & 3
    for b in blobs:
        nrow, ncol, nz0, nz1 = (int(x) for x in b[:16].view(np.int32))
        orow = b[off_outrow:off_outrow + 4 * 8 * rh].view(np.int32)
        cl = b[off_cols:off_cols + 4 * 4 * maxks].view(np.int32)
        af = b[off_afrag:off_afrag + 8 * 32 * rh * maxks].view(np.float64).reshape(maxks, rh, 32)
        D = np.zeros((8 * rh, 4 * maxks))
        for ks in range(maxks):
            for h in range(rh):
                if (nz0, nz1)[h] >> ks & 1:
                    D[g + 8 * h, 4 * ks + t] = af[ks, h]
        assert not D[nrow:].any() and not D[:, ncol:].any()
        rr, jj = np.nonzero(D)
        rows.append(orow[rr]); cols.append(cl[jj]); vals.append(D[rr, jj])
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


def decode_runs_blobs(rplan, n):
    """Rebuild the matrix from the run records of csr_spmm_runs_kernel: a staged row index maps back to a B row through
    the run table (first B row, length, first staged row), the way the producer's copies place the rows."""
    mr, stride = rplan["max_rows"], rplan["stride"]
    r4 = lambda x: (x + 3) // 4 * 4
    off_outrow = 16
    off_goff = off_outrow + 4 * r4(mr)
    off_runs = off_goff + 4 * r4(mr + 1)
    off_ent = (off_runs + 8 * rplan["max_runs"] + 15) // 16 * 16
    rows, cols, vals = [], [], []
    for b in np.asarray(rplan["blobs"]).reshape(-1, stride):
        nrow, nrun, nbrow, nent = (int(x) for x in b[:16].view(np.int32))
        orow = b[off_outrow:off_outrow + 4 * nrow].view(np.int32)
        goff = b[off_goff:off_goff + 4 * (nrow + 1)].view(np.int32)
        staged = np.full(nbrow, -1, dtype=np.int64)
        for start, lo in b[off_runs:off_runs + 8 * nrun].view(np.int32).reshape(-1, 2):
            ln, off = int(lo) & 0xffff, (int(lo) >> 16) & 0xffff
            staged[off:off + ln] = start + np.arange(ln)
        ent = b[off_ent:off_ent + 16 * nent]
        slot, v = ent.view(np.int32).reshape(-1, 4)[:, 0], ent.view(np.float64)[1::2]
        rows.append(np.repeat(orow, np.diff(goff))); cols.append(staged[slot]); vals.append(v)
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


@contextlib.contextmanager
def emulated_device():
    """Context manager: hippyflow_b200 runs on torch CPU tensors.  Yields the torch.device to pass as ``device=``."""
    from hippyflow_b200 import _lib as K
    saved_K = dict(K.__dict__)
    saved_torch = {"current_stream": torch.cuda.current_stream, "Stream": torch.cuda.Stream, "Event": torch.cuda.Event,
                   "stream": torch.cuda.stream, "synchronize": torch.cuda.synchronize,
                   "current_device": torch.cuda.current_device}
    saved_empty = torch.empty
    import os
    saved_env = os.environ.get("HFB_DEVICE_CHOL")
    os.environ["HFB_DEVICE_CHOL"] = "0"      # the test double covers the host-controlled orthonormalisation route only
    had_record = "record_stream" in torch.Tensor.__dict__
    saved_record = torch.Tensor.__dict__.get("record_stream")
    counter = {"n": 0}
    cache = {}

    def out_or_new(out, rows, cols, dev):
        return out if out is not None else K.padded_empty(rows, cols, dev)

    def dgemm(layout, A, B, out=None, alpha=1.0, splits=0, symmetric=False, accumulate=False, b_upper=False):
        K._req(A, "A"), K._req(B, "B")
        opA = A.t() if layout == K.HFB_TN else A
        opB = B.t() if layout == K.HFB_NT else B
        if opA.shape[1] != opB.shape[0]:
            raise K.HfbError("dgemm: inner dimensions differ")
        if A.data_ptr() % 16 or B.data_ptr() % 16 or K._ld(A) % 2 or K._ld(B) % 2:
            raise K.HfbError("hfb_dgemm: operand not 16-byte aligned or odd leading dimension (code -2)")
        R = alpha * (opA @ opB)
        if symmetric and R.shape[0] == R.shape[1]:
            R = torch.triu(R) + torch.triu(R, 1).t()
        if accumulate:
            if out is None or symmetric:
                raise K.HfbError("dgemm: accumulate needs an existing out")
            out.add_(R)
            counter["n"] += 1
            return out
        fresh = out_or_new(out, R.shape[0], R.shape[1], A.device)
        if tuple(fresh.shape) != tuple(R.shape):
            raise K.HfbError("dgemm: out has shape %s, expected %s" % (tuple(fresh.shape), tuple(R.shape)))
        fresh.copy_(R)
        counter["n"] += 1
        return fresh

    def dgemm_batched_small(layout, A, B, out, alpha=1.0):
        opA = A.transpose(1, 2) if layout == K.HFB_TN else A
        out.copy_(alpha * torch.matmul(opA, B))
        counter["n"] += 1
        return out

    def _spmm(Msp, B, out):
        n, m = B.shape
        res = torch.from_numpy(np.ascontiguousarray(Msp @ B.numpy()))
        out = out_or_new(out, Msp.shape[0], m, B.device)
        out.copy_(res)
        counter["n"] += 1
        return out

    def csr_spmm(rowptr, colind, val, B, out=None, order=None):
        return _spmm(_scipy_csr(rowptr, colind, val, B.shape[0]), B, out)

    def csr_spmm_dmma(plan, B, out=None):
        key = ("dmma", id(plan))
        if key not in cache:
            cache[key] = _decode_panel_blobs(K, plan)
        return _spmm(cache[key], B, out)

    def csr_spmm_dmma_frag(plan, B, out=None, chunk_cols=0):
        key = ("frag", id(plan))
        if key not in cache:
            cache[key] = decode_frag_blobs(K, plan["fblobs"].numpy(), plan["max_rows"], plan["max_cols_cap"], plan["order"].numel())
        return _spmm(cache[key], B, out)

    def csr_spmm_dmma_ring(plan, B, out=None):
        return csr_spmm_dmma_frag(plan, B, out)      # same records

    def csr_spmm_runs(rplan, B, out=None):
        key = ("runs", id(rplan))
        if key not in cache:
            cache[key] = decode_runs_blobs({**rplan, "blobs": rplan["blobs"].numpy()}, B.shape[0])
        return _spmm(cache[key], B, out)

    def csr_spmm_rows(rowptr, colind, val, X, out=None):
        Msp = _scipy_csr(rowptr, colind, val, X.shape[1])
        res = torch.from_numpy(np.ascontiguousarray((Msp @ X.numpy().T).T))
        out = out_or_new(out, X.shape[0], Msp.shape[0], X.device)
        out.copy_(res)
        counter["n"] += 1
        return out

    def coldot(X, Y):
        counter["n"] += 1
        return (X * Y).sum(0)

    def rowdot(X, Y):
        counter["n"] += 1
        return (X * Y).sum(1)

    def colscale_(X, s):
        X.mul_(s.reshape(1, -1))
        return X

    def colsum(X, scale=1.0, weights=None):
        counter["n"] += 1
        if weights is not None:
            if weights.numel() != X.shape[0]:
                raise K.HfbError("colsum: weights must be a contiguous vector with one entry per row")
            return scale * (weights.reshape(1, -1) @ X).reshape(-1)
        return scale * X.sum(0)

    def subtract_row_(X, shift):
        X.sub_(shift.reshape(1, -1))
        counter["n"] += 1
        return X

    def rank1_update_(Y, a, x, y):
        if x.numel() != Y.shape[0] or y.numel() != Y.shape[1]:
            raise K.HfbError("rank1_update_: vector lengths do not match Y")
        Y.add_(a * torch.outer(x, y))
        counter["n"] += 1
        return Y

    def axpby_(a, X, b, Y):
        Y.copy_(a * X + (b * Y if b != 0.0 else 0.0))
        return Y

    def axpby_cols_(a, X, b, Y):
        av = a.reshape(1, -1) if a is not None else 1.0
        bv = b.reshape(1, -1) if b is not None else 1.0
        Y.copy_(av * X + bv * Y)
        return Y

    def rowscale(s, X, out=None):
        out = out_or_new(out, X.shape[0], X.shape[1], X.device)
        out.copy_(s.reshape(-1, 1) * X)
        return out

    def fill_random_(X, seed, row_offset=0, kind="normal"):
        g = torch.Generator().manual_seed(int(seed) * 1000003 + int(row_offset))
        X.copy_(torch.randn(X.shape, dtype=torch.float64, generator=g) if kind == "normal"
                else torch.rand(X.shape, dtype=torch.float64, generator=g))
        return X

    def empty_nopin(*a, **k):
        k.pop("pin_memory", None)
        return saved_empty(*a, **k)

    patches = dict(dgemm=dgemm, dgemm_batched_small=dgemm_batched_small, csr_spmm=csr_spmm, csr_spmm_dmma=csr_spmm_dmma, csr_spmm_dmma_frag=csr_spmm_dmma_frag, csr_spmm_dmma_ring=csr_spmm_dmma_ring, csr_spmm_runs=csr_spmm_runs, csr_spmm_rows=csr_spmm_rows, coldot=coldot, rowdot=rowdot, colscale_=colscale_,
                   colsum=colsum, subtract_row_=subtract_row_, rank1_update_=rank1_update_, axpby_=axpby_,
                   axpby_cols_=axpby_cols_, rowscale=rowscale, fill_random_=fill_random_,
                   measure_dmma_peak=lambda device: 1.0, launch_count=lambda: counter["n"],
                   is_device_tensor=lambda t: isinstance(t, torch.Tensor))
    try:
        for k, v in patches.items():
            setattr(K, k, v)
        torch.cuda.current_stream = lambda *a, **k: _FakeStream()
        torch.cuda.Stream = _FakeStream
        torch.cuda.Event = _FakeEvent
        torch.cuda.stream = lambda s: contextlib.nullcontext()
        torch.cuda.synchronize = lambda *a, **k: None
        torch.cuda.current_device = lambda: 0
        torch.Tensor.record_stream = lambda self, s: None
        torch.empty = empty_nopin
        yield torch.device("cpu")
    finally:
        for k in patches:
            setattr(K, k, saved_K[k])
        for k, v in saved_torch.items():
            setattr(torch.cuda, k, v)
        torch.empty = saved_empty
        if saved_env is None:
            os.environ.pop("HFB_DEVICE_CHOL", None)
        else:
            os.environ["HFB_DEVICE_CHOL"] = saved_env
        if had_record:
            torch.Tensor.record_stream = saved_record
        else:
            del torch.Tensor.record_stream
