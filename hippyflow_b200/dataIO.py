"""On-disk data contract of the hot path: read the arrays the reference's data generator
writes, shard them by sample straight into device buffers, and write projector files under the names the
reference's drivers and loaders use.

Formats (hippyflow/modeling/dataGenerator.py:636-655): ``mq_data.npz`` {m_data (N,dM), q_data (N,dQ)},
``JstarPhi_data.npz`` {JstarPhi_data (N,dM,rQ), Phi, MPhi}, ``JPsi_data.npz`` {JPsi_data (N,dQ,rM), Psi,
input_encoder}, ``Jsvd_data.npz`` {U_data (N,dQ,r), sigma_data (N,r), V_data (N,dM,r)}; per-sample files
``m_sample_<i>.npy`` / ``q_sample_<i>.npy`` (dataGenerator.py:560-566).  Projector files: ``POD_projector.npy``,
``POD_d.npy`` (PODProjector.py:383-384), ``KLE_decoder.npy`` / ``KLE_d.npy`` (KLEProjector.py:191-192),
``AS_<N>_input_decoder.npy`` / ``AS_<N>_d_GN.npy`` (activeSubspaceProjector.py:475-480).
"""
import os

import numpy as np
import torch

from . import _lib as K


def shard_bounds(n_total, rank, world):
    """Contiguous equal shards (the 'avg' collective assumes equal counts, activeSubspaceProjector.py:429-430);
    a remainder is dropped from the tail so that every rank holds floor(n_total / world) samples."""
    per = n_total // world
    return rank * per, (rank + 1) * per


def load_mq_data(file_path, device, rank=0, world=1, name="mq_data.npz"):
    """(m_shard, q_shard) as padded device blocks, rows = this rank's samples."""
    with np.load(os.path.join(file_path, name)) as z:
        lo, hi = shard_bounds(z["m_data"].shape[0], rank, world)
        m = K.to_padded(z["m_data"][lo:hi], device)
        q = K.to_padded(z["q_data"][lo:hi], device)
    return m, q


def load_mq_samples(data_path, device, rank=0, world=1):
    """Same from the per-sample ``m_sample_<i>.npy`` / ``q_sample_<i>.npy`` files."""
    idx = sorted(int(f[len("m_sample_"):-4]) for f in os.listdir(data_path) if f.startswith("m_sample_") and f.endswith(".npy"))
    lo, hi = shard_bounds(len(idx), rank, world)
    ms = np.stack([np.load(os.path.join(data_path, "m_sample_%d.npy" % i)) for i in idx[lo:hi]])
    qs = np.stack([np.load(os.path.join(data_path, "q_sample_%d.npy" % i)) for i in idx[lo:hi]])
    return K.to_padded(ms, device), K.to_padded(qs, device)


def load_jacobian_svd_factor(file_path, device, rank=0, world=1, name="Jsvd_data.npz"):
    """Stored low-rank Jacobians J_i = U_i diag(sigma_i) V_i^T  ->  the stacked factor rows sigma_i * V_i^T,
    shape (N_loc * r, dM), with block size r: since U_i^T U_i = I, (1/N) sum J_i^T J_i = (1/N) sum V_i sigma_i^2 V_i^T,
    so the factor is the ``Xt`` of linalg.SampleCovariance (block = r).  Returns (Xt, r)."""
    with np.load(os.path.join(file_path, name)) as z:
        lo, hi = shard_bounds(z["sigma_data"].shape[0], rank, world)
        V = z["V_data"][lo:hi]                                   # (N_loc, dM, r)
        s = z["sigma_data"][lo:hi]                               # (N_loc, r)
    F = np.ascontiguousarray(np.transpose(V * s[:, None, :], (0, 2, 1)))   # (N_loc, r, dM)
    n_loc, r, dM = F.shape
    return K.to_padded(F.reshape(n_loc * r, dM), device), r


def save_pod(output_directory, U, d):
    np.save(os.path.join(output_directory, "POD_projector"), _dense(U))
    np.save(os.path.join(output_directory, "POD_d"), np.asarray(d))


def save_kle(output_directory, V, d, name="KLE_decoder"):
    np.save(os.path.join(output_directory, name), _dense(V))
    np.save(os.path.join(output_directory, "KLE_d"), np.asarray(d))


def save_active_subspace(output_directory, V, d, n_samples_total, name_suffix=None, decoder_name="_input_decoder",
                         d_name="_d_GN"):
    name = "AS_" + str(int(n_samples_total)) + (name_suffix or "")
    np.save(os.path.join(output_directory, name + decoder_name), _dense(V))
    np.save(os.path.join(output_directory, name + d_name), np.asarray(d))


def _dense(U):
    if hasattr(U, "to_dense"):
        return U.to_dense()
    if isinstance(U, torch.Tensor):
        return U.cpu().numpy()
    return np.asarray(U)


def reduce_dataset(file_path, input_encoder, output_encoder, device, q_shift=None, rank=0, world=1, out_name=None):
    """Project a stored ``mq_data.npz`` onto the bases: m_r = m_data @ input_encoder, q_r = (q_data - q_shift) @
    output_encoder, for this rank's shard; optionally written as ``<out_name>`` with the keys m_data / q_data.
    Returns (m_r, q_r) device blocks."""
    from .modeling.projection import project_data
    m, q = load_mq_data(file_path, device, rank, world)
    if q_shift is not None:
        K.subtract_row_(q, torch.as_tensor(np.asarray(q_shift, dtype=np.float64), device=device))
    m_r = project_data(m, input_encoder, device)
    q_r = project_data(q, output_encoder, device)
    if out_name is not None:
        np.savez_compressed(os.path.join(file_path, out_name), m_data=m_r.cpu().numpy(), q_data=q_r.cpu().numpy())
    return m_r, q_r
