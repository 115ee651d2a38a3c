"""Synthetic stored-data generators: snapshots, parameter draws, Jacobians and the
P1 mass matrix of a structured triangulation.  Host (NumPy) versions feed the parity tests and the
CPU baseline; ``*_device`` versions build the same kind of data directly in HBM for the benchmark
sizes that do not fit host RAM.  The PDE solves that produce such data in the reference
(hippyflow/modeling/PODProjector.py:343-357, dataGenerator.py:88-248) are upstream and out of scope.
"""
import numpy as np
import scipy.sparse as sp


def p1_mass_matrix(nx, ny=None):
    """P1 mass matrix on the unit square split into nx*ny*2 right triangles; (nx+1)(ny+1) dofs.
    Element matrix (area/12) [[2,1,1],[1,2,1],[1,1,2]] -> 7-point stencil, SPD.  int32 CSR, like
    the SciPy export of the reference (PODProjector.py:695-697)."""
    ny = nx if ny is None else ny
    hx, hy = 1.0 / nx, 1.0 / ny
    area = 0.5 * hx * hy
    idx = np.arange((nx + 1) * (ny + 1)).reshape(ny + 1, nx + 1)
    v00, v10, v01, v11 = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    tris = np.concatenate([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], 0)
    Me = (area / 12.0) * np.array([[2.0, 1, 1], [1, 2, 1], [1, 1, 2]])
    rows = np.repeat(tris, 3, axis=1).ravel()
    cols = np.tile(tris, (1, 3)).ravel()
    vals = np.tile(Me.ravel(), tris.shape[0])
    n = idx.size
    M = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    M.sum_duplicates()
    M.sort_indices()
    M.indices = M.indices.astype(np.int32)
    M.indptr = M.indptr.astype(np.int32)
    return M


def _mode_table(r0):
    """First r0 (a, b) frequency pairs ordered by a^2 + b^2 (ties by a)."""
    kmax = int(np.ceil(np.sqrt(r0))) + 2
    pairs = [(a * a + b * b, a, b) for a in range(1, kmax + 1) for b in range(1, kmax + 1)]
    pairs.sort()
    return np.array([(a, b) for _, a, b in pairs[:r0]])


def smooth_modes(n, r0):
    """(n, r0) smooth modes.  For n a perfect square the modes are products of sines on the grid,
    otherwise 1-D sines on [0,1]."""
    side = int(round(np.sqrt(n)))
    if side * side == n:
        x = np.linspace(0.0, 1.0, side)
        ab = _mode_table(r0)
        Sx = np.sin(np.pi * np.outer(x, np.arange(1, ab.max() + 1)))          # (side, kmax)
        Phi = (Sx[:, None, ab[:, 0] - 1] * Sx[None, :, ab[:, 1] - 1]).reshape(n, r0)
    else:
        x = np.linspace(0.0, 1.0, n)
        Phi = np.sin(np.pi * np.outer(x, np.arange(1, r0 + 1)))
    return Phi


def snapshots(n, N, r0=64, decay=1.0, eps=1e-6, seed=0):
    """(N, n) C-order snapshot array, rows = samples (the layout of ``u_data``, PODProjector.py:726):
    X = Phi0 diag(j^-decay) G + eps * noise."""
    rng = np.random.default_rng(seed)
    r0 = min(r0, n)
    Phi = smooth_modes(n, r0)
    sig = np.arange(1, r0 + 1, dtype=np.float64) ** (-decay)
    G = rng.standard_normal((N, r0))
    X = (G * sig) @ Phi.T
    X += eps * rng.standard_normal((N, n))
    return np.ascontiguousarray(X)


def jacobians(N, dQ, dM, r0=32, decay=1.0, seed=0):
    """(N, dQ, dM) C-order stored Jacobians (the layout MeanJTJfromDataOperator assumes,
    operatorWrappers.py:62-64): J_i = A0 diag(sigma) (B0 + 0.1 G_i)."""
    rng = np.random.default_rng(seed)
    r0 = min(r0, dQ, dM)
    A0 = np.linalg.qr(rng.standard_normal((dQ, r0)))[0]
    B0 = smooth_modes(dM, r0).T                                    # (r0, dM)
    B0 = B0 / np.linalg.norm(B0, axis=1, keepdims=True)
    sig = np.arange(1, r0 + 1, dtype=np.float64) ** (-decay)
    G = rng.standard_normal((N, r0, dM)) / np.sqrt(dM)
    J = np.einsum("qr,irm->iqm", A0 * sig, B0[None] + 0.1 * G)
    return np.ascontiguousarray(J)


def gaussian_omega(n, m, seed=1):
    """Gaussian test matrix (n, m), the role of hp.parRandom.normal(1., Omega)
    (PODProjector.py:367-372).  Generated on the host so the SAME Omega feeds oracle and GPU."""
    return np.random.default_rng(seed).standard_normal((n, m))


# --------------------------------------------------------------------------- device-side generators (benchmark sizes)
def p1_mass_matrix_for(n):
    """Mass matrix with exactly n = (nx+1)^2 dofs (n must be a perfect square)."""
    side = int(round(np.sqrt(n)))
    assert side * side == n, "n must be a perfect square"
    return p1_mass_matrix(side - 1)


def snapshots_device(n, N, device, r0=512, decay=1.0, eps=1e-6, seed=0, row_offset=0):
    """(N, n) snapshot shard generated in HBM: X = G diag(j^-decay) Phi0^T + eps * noise with G and the noise from
    the counter-based generator keyed by (seed, GLOBAL sample index), so that the data does not depend on how
    the samples are sharded over GPUs (``row_offset`` = first global sample of this shard).  Same construction as
    ``snapshots`` (different random stream).  Uses the library's own GEMM for G Phi0^T."""
    import torch
    from . import _lib as K
    side = int(round(np.sqrt(n)))
    r0 = min(r0, n)
    if side * side == n:
        x = torch.linspace(0.0, 1.0, side, dtype=torch.float64, device=device)
        ab = torch.as_tensor(_mode_table(r0), device=device)
        Sx = torch.sin(np.pi * x[:, None] * torch.arange(1, int(ab.max()) + 1, device=device, dtype=torch.float64)[None, :])
        Phi = K.padded_empty(n, r0, device)
        Phi.copy_((Sx[:, None, ab[:, 0] - 1] * Sx[None, :, ab[:, 1] - 1]).reshape(n, r0))
    else:
        x = torch.linspace(0.0, 1.0, n, dtype=torch.float64, device=device)
        Phi = K.padded_empty(n, r0, device)
        Phi.copy_(torch.sin(np.pi * x[:, None] * torch.arange(1, r0 + 1, device=device, dtype=torch.float64)[None, :]))
    G = K.padded_empty(N, r0, device)
    K.fill_random_(G, seed, row_offset=row_offset)
    sig = torch.arange(1, r0 + 1, dtype=torch.float64, device=device) ** (-decay)
    K.colscale_(G, sig)
    X = K.padded_empty(N, n, device)
    K.dgemm(K.HFB_NT, G, Phi, out=X)                       # (N x r0) (n x r0)^T
    del Phi, G
    chunk = max(1, min(N, (1 << 28) // max(n, 1)))         # noise in chunks of <= 2 GiB
    for i0 in range(0, N, chunk):
        i1 = min(N, i0 + chunk)
        noise = K.padded_empty(i1 - i0, n, device)
        K.fill_random_(noise, seed + 0x9E3779B9, row_offset=row_offset + i0)
        K.axpby_(eps, noise, 1.0, X[i0:i1])
        del noise
    return X
