"""Double-pass randomized eigensolvers on the device: hIPPYlib's ``doublePass`` / ``doublePassG``
(called at hippyflow/modeling/PODProjector.py:376, activeSubspaceProjector.py:449-461,556-577,654,
KLEProjector.py:163,177) for operators that live in HBM.

Algorithm (s = 1 at every reference call site):
    doublePass (A, Omega, k):          Q = orth(A Omega);             T = Q^T A Q; eigh; U = Q V
    doublePassG(A, B, Binv, Omega, k): Q = B-orth(Binv A Omega);      T = Q^T A Q; eigh; U = Q V, U^T B U = I
Differences from the column-by-column hIPPYlib code, none of which changes d or span(U):
  * A is applied to all m = k + p columns at once (two DMMA GEMMs + one NCCL allreduce per pass);
  * (B-)orthonormalisation is (shifted) Cholesky-QR on the Gram matrix instead of MGS (linalg.b_orthonormalize);
    when the first pass is well conditioned its clean-up pass is folded into the small matrices
    (T = S2^T T1 S2, U = Q1 (S2 V)) and the host factorisation overlaps with the pass-2 GEMM;
  * when the operator exposes ``rayleigh`` the small matrix T = Q^T A Q is formed as a Gram matrix of the
    projected samples (one GEMM less, only an (m x m) allreduce); ``faithful=True`` forces T = (A Q)^T Q.
"""
import numpy as np

from . import _lib as K
import torch

from .linalg import b_orthonormalize, cleanup_factor, top_k_eig
from .multivector import DeviceMultiVector


def _block_apply(A, X):
    if hasattr(A, "matMvMult") and getattr(A, "overwrites", False):
        Y = DeviceMultiVector(K.padded_empty(X.tensor().shape[0], X.nvec(), X.tensor().device))   # fully overwritten
        Y.adoptable = True                      # scratch of this solver: the operator may hand back a block of its own
        A.matMvMult(X, Y)
        return Y
    Y = DeviceMultiVector(X.tensor().shape[0], X.nvec(), device=X.tensor().device)   # zeroed: operators may accumulate into Y (activeSubspaceProjector.py:214-221)
    if hasattr(A, "matMvMult"):
        A.matMvMult(X, Y)
    else:  # hippylib MatMvMult fallback: column loop over A.mult
        for j in range(X.nvec()):
            A.mult(X[j], Y[j])
    return Y


_side_streams = {}


def _fetch_async(t):
    """Start a device -> pinned-host copy of the small matrix ``t`` on a side stream; returns (host tensor, event).
    The copy waits only for the work queued so far, so kernels launched afterwards overlap with it."""
    dev = t.device
    side = _side_streams.get(dev.index)
    if side is None:
        side = _side_streams[dev.index] = torch.cuda.Stream(device=dev)
    ready = torch.cuda.Event()
    ready.record()
    host = torch.empty(tuple(t.shape), dtype=t.dtype, pin_memory=True)
    done = torch.cuda.Event()
    with torch.cuda.stream(side):
        side.wait_event(ready)
        host.copy_(t, non_blocking=True)
        done.record(side)
    t.record_stream(side)
    return host, done


def _rayleigh_matrix(A, Q, BQ, faithful):
    """T = Q^T A Q as (device matrix or None, host matrix or None); the big GEMM(s) are only queued here."""
    if hasattr(A, "rayleigh_device") and not faithful:
        return A.rayleigh_device(Q, BQ), None
    if hasattr(A, "rayleigh") and not faithful:
        return None, A.rayleigh(Q, BQ)
    AQ = _block_apply(A, Q)
    return None, AQ.dot_mv(Q)


def _rayleigh_ritz(A, Q, BQ, k, oinfo, faithful):
    """T = Q^T A Q, top-k eigenpairs, and the (m x k) DEVICE coefficient matrix C with U = Q C.

    ``oinfo`` is the info dict of ``b_orthonormalize``.  Three cases:
      * ``oinfo["pending"]`` (device Cholesky-QR, nothing read back so far): T1 is formed, the clean-up factor S2 is folded
        in on the device (T = S2^T T1 S2), and T travels to the host together with the two status vectors -- ONE
        synchronisation for the whole orthonormalisation + Rayleigh-Ritz phase.  Returns None when the status vectors say
        the optimistic path was not valid (the caller then redoes the orthonormalisation under host control).
      * ``oinfo["gram"]`` (host path with deferred clean-up): G1 is fetched on a side stream while the pass-2 GEMM runs and
        the host computes S2 = chol(G1)^-1 in its shadow.
      * neither: Q is B-orthonormal to round-off.
    The only host arithmetic is eigh(T) (hIPPYlib: np.linalg.eigh), plus the Cholesky factor of G1 in the second case."""
    dev = Q.tensor().device
    m = Q.nvec()
    pending = oinfo.pop("pending", None)
    gram = oinfo.pop("gram", None)
    if gram is not None:
        G1h, done = _fetch_async(gram)
    Td, T = _rayleigh_matrix(A, Q, BQ, faithful)
    S2d = None
    if pending is not None:
        S2d = pending["S2"]
        if pending.get("side") is not None:
            torch.cuda.current_stream(dev).wait_stream(pending["side"])   # S2 was computed on a side stream
        if Td is None:
            Td = K.to_padded(np.ascontiguousarray(T), dev)
        Td = K.dgemm(K.HFB_TN, S2d, K.dgemm(K.HFB_NN, Td, S2d))          # S2^T T S2
        stats = torch.stack([pending["stat1"], pending["stat2"]]).cpu().numpy()   # the one synchronisation of this phase
        from .linalg import pending_ok
        if not pending_ok(stats, m):
            return None
        oinfo["cond"] = [float(stats[0][2]), float(stats[1][2])]
        oinfo["passes"] = 2                                              # pass 2 folded into the small matrices
    elif gram is not None:
        done.synchronize()
        S2d = K.to_padded(cleanup_factor(G1h.numpy()), dev)
        if Td is None:
            Td = K.to_padded(np.ascontiguousarray(T), dev)
        Td = K.dgemm(K.HFB_TN, S2d, K.dgemm(K.HFB_NN, Td, S2d))          # S2^T T S2
    if Td is not None:
        T = Td.cpu().numpy()
    d, V = top_k_eig(T, k)
    Vd = K.to_padded(V, dev)
    return d, (Vd if S2d is None else K.dgemm(K.HFB_NN, S2d, Vd))


def _orthonormalize_and_ritz(A, Y, B, k, faithful, return_BQ):
    """(B-)orthonormalise the sketch Y and solve the projected problem; the optimistic device path is tried first and
    replaced by the host-controlled passes (on the untouched sketch) when its status check fails."""
    Qt, BQt, oinfo = b_orthonormalize(Y, B, return_BQ=return_BQ, defer_last=True)
    optimistic = "pending" in oinfo
    Q = DeviceMultiVector(Qt)
    BQ = DeviceMultiVector(BQt) if return_BQ else Q
    res = _rayleigh_ritz(A, Q, BQ, k, oinfo, faithful)
    if res is None:
        assert optimistic
        Qt, BQt, oinfo = b_orthonormalize(Y, B, return_BQ=return_BQ, defer_last=True, device_chol=False)
        Q = DeviceMultiVector(Qt)
        BQ = DeviceMultiVector(BQt) if return_BQ else Q
        res = _rayleigh_ritz(A, Q, BQ, k, oinfo, faithful)
        oinfo["route"] = "host (device status rejected)"
    d, C = res
    return d, Q, C, oinfo


def doublePass(A, Omega, k, s=1, faithful=False, info=None):
    """d (k,) descending, U DeviceMultiVector (n, k) with U^T U = I."""
    nvec = Omega.nvec()
    assert k <= nvec
    Q = Omega                                   # not modified: every apply writes a fresh block
    for _ in range(s):
        Q = _block_apply(A, Q)
    d, Q, C, oinfo = _orthonormalize_and_ritz(A, Q.tensor(), None, k, faithful, return_BQ=False)
    U = DeviceMultiVector(K.dgemm(K.HFB_NN, Q.tensor(), C))
    if info is not None:
        info.update(oinfo)
    return d, U


def doublePassG(A, B, Binv, Omega, k, s=1, faithful=False, info=None, Q0=None):
    """Generalised problem A u = lambda B u.  ``B`` is a linalg.CsrMatrix (or any object with
    ``matmat``); ``Binv`` an object with ``solve_block(Y) -> B^-1 Y``; it is not needed when A exposes
    ``solveB_matMvMult`` (A = B C B, so B^-1 A X = C B X without a solve).
    ``Q0``: the block B^-1 A Omega when the caller has already formed it (s = 1 only).
    Returns d (k,), U (n, k) with U^T B U = I."""
    nvec = Omega.nvec()
    assert k <= nvec
    Q = Omega                                   # not modified: every apply writes a fresh block
    for _ in range(s if Q0 is None else 0):
        if hasattr(A, "solveB_matMvMult") and getattr(A, "B", None) is B and (Binv is None or not faithful):
            Y = DeviceMultiVector(K.padded_empty(Q.tensor().shape[0], nvec, Q.tensor().device))   # fully overwritten
            Y.adoptable = True                  # scratch of this solver: the operator may hand back a block of its own
            A.solveB_matMvMult(Q, Y)
            Q = Y
        else:
            Ybar = _block_apply(A, Q)
            Q = DeviceMultiVector(Binv.solve_block(Ybar.tensor()))
    if Q0 is not None:
        Q = Q0                                  # range-finder block B^-1 A Omega supplied by the caller (pipelined upload)
    d, Q, C, oinfo = _orthonormalize_and_ritz(A, Q.tensor(), B, k, faithful, return_BQ=True)
    U = DeviceMultiVector(K.dgemm(K.HFB_NN, Q.tensor(), C))
    if info is not None:
        info.update(oinfo)
    return d, U
