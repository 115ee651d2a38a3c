"""On-disk data contract of the hot path: read the arrays the reference's data generator
writes, shard them by sample straight into device buffers, and write projector files under the names the
reference's drivers and loaders use.

Formats (hippyflow/modeling/dataGenerator.py:636-655): ``mq_data.npz`` {m_data (N,dM), q_data (N,dQ)},
``JstarPhi_data.npz`` {JstarPhi_data (N,dM,rQ), Phi, MPhi}, ``JPsi_data.npz`` {JPsi_data (N,dQ,rM), Psi,
input_encoder}, ``Jsvd_data.npz`` {U_data (N,dQ,r), sigma_data (N,r), V_data (N,dM,r)}; per-sample files
``m_sample_<i>.npy`` / ``q_sample_<i>.npy`` (dataGenerator.py:560-566).  Projector files: ``POD_projector.npy``,
``POD_d.npy`` (PODProjector.py:383-384), ``KLE_decoder.npy`` / ``KLE_d.npy`` (KLEProjector.py:191-192),
``AS_<N>_input_decoder.npy`` / ``AS_<N>_d_GN.npy`` (activeSubspaceProjector.py:475-480).
Per-rank archives: ``mq_on_rank<r>.npz`` (PODProjector.py:234-237), ``mq_on_proc<r>.npz`` / ``J_on_proc<r>.npz``
(activeSubspaceProjector.py:860-878), and the ``mq_on_rank_<r>.npz`` spelling the application loaders read
(applications/confusion/confusion_utilities.py:50-54).

Everything here is host-side file I/O plus one upload per array; nothing numerical happens on the CPU.
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


def _npz_member_header(fh):
    """(shape, fortran_order, dtype) from the .npy header at the start of an open zip member."""
    version = np.lib.format.read_magic(fh)
    if version == (1, 0):
        return np.lib.format.read_array_header_1_0(fh)
    if version == (2, 0):
        return np.lib.format.read_array_header_2_0(fh)
    raise ValueError("unsupported .npy header version %r" % (version,))


def _npz_member_shape(path, key):
    """Shape of array ``key`` of an .npz archive from the .npy header of the zip member alone (``np.load(path)[key].shape``
    would read and inflate the whole member -- every rank doing that for every archive is O(world x dataset))."""
    import zipfile
    with zipfile.ZipFile(path) as zf, zf.open(key + ".npy") as fh:
        return tuple(_npz_member_header(fh)[0])


def _npz_member_rows(path, key, lo=0, hi=None, chunk_bytes=1 << 24):
    """Rows [lo, hi) of array ``key`` of an .npz archive, streamed from the zip member: the bytes before row ``lo`` are
    skipped (a seek for stored members, inflate-and-discard in 16 MB pieces for deflated ones), the shard is inflated
    straight into its own buffer and nothing after row ``hi`` is read -- host memory stays O(shard), where
    ``np.load(path)[key][lo:hi]`` materialises the whole member on every rank.  Object arrays and Fortran-ordered members
    take the NumPy route."""
    import zipfile
    with zipfile.ZipFile(path) as zf, zf.open(key + ".npy") as fh:
        shape, fortran, dtype = _npz_member_header(fh)
        if dtype.hasobject or fortran or len(shape) == 0:
            with np.load(path, allow_pickle=True) as z:
                a = z[key]
                return a if len(shape) == 0 else a[lo:hi]
        n = shape[0]
        lo = max(0, min(n, lo))
        hi = n if hi is None else max(lo, min(n, hi))
        row_bytes = int(np.prod(shape[1:], dtype=np.int64)) * dtype.itemsize
        out = np.empty((hi - lo,) + tuple(shape[1:]), dtype=dtype)
        skip = lo * row_bytes
        if skip:
            try:
                fh.seek(skip, 1)                                # stored members seek; deflated ones read and discard inside zipfile
            except Exception:
                while skip > 0:
                    got = fh.read(min(skip, chunk_bytes))
                    if not got:
                        raise EOFError("%s[%s]: truncated member" % (path, key))
                    skip -= len(got)
        view = memoryview(out.reshape(-1).view(np.uint8)) if out.size else memoryview(b"")
        done = 0
        while done < len(view):
            got = fh.readinto(view[done:done + chunk_bytes])
            if not got:
                raise EOFError("%s[%s]: truncated member" % (path, key))
            done += got
    return out


def load_mq_data(file_path, device, rank=0, world=1, name="mq_data.npz"):
    """(m_shard, q_shard) as padded device blocks, rows = this rank's samples (only the shard is held on the host)."""
    path = os.path.join(file_path, name)
    lo, hi = shard_bounds(_npz_member_shape(path, "m_data")[0], rank, world)
    m = K.to_padded(_npz_member_rows(path, "m_data", lo, hi), device)
    q = K.to_padded(_npz_member_rows(path, "q_data", lo, hi), device)
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
    path = os.path.join(file_path, name)
    lo, hi = shard_bounds(_npz_member_shape(path, "sigma_data")[0], rank, world)
    V = _npz_member_rows(path, "V_data", lo, hi)                 # (N_loc, dM, r)
    s = _npz_member_rows(path, "sigma_data", lo, hi)             # (N_loc, r)
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


# ------------------------------------------------------------------------------------------- per-rank archives
_RANK_FILE_PATTERNS = ("mq_on_rank%d.npz", "mq_on_rank_%d.npz", "mq_on_proc%d.npz")


def _rank_files(data_dir, patterns):
    """Consecutive per-rank archives <pattern % 0>, <pattern % 1>, ... for the first pattern that exists."""
    for pat in patterns:
        files, r = [], 0
        while os.path.exists(os.path.join(data_dir, pat % r)):
            files.append(os.path.join(data_dir, pat % r))
            r += 1
        if files:
            return files
    raise FileNotFoundError("no per-rank archive (%s) under %s" % (", ".join(p % 0 for p in patterns), data_dir))


def _sharded_rows_from_archives(files, keys, rank, world):
    """Rows [lo, hi) of the virtual concatenation of ``keys`` over ``files`` without materialising the whole dataset:
    only archives that intersect this rank's shard are decompressed.  Returns {key: ndarray}."""
    counts = [int(_npz_member_shape(f, keys[0])[0]) for f in files]     # headers only: nothing is decompressed here
    offs = np.concatenate([[0], np.cumsum(counts)])
    lo, hi = shard_bounds(int(offs[-1]), rank, world)
    parts = {k: [] for k in keys}
    for f, a, b in zip(files, offs[:-1], offs[1:]):
        s0, s1 = max(lo, int(a)), min(hi, int(b))
        if s0 >= s1:
            continue
        for k in keys:
            parts[k].append(_npz_member_rows(f, k, s0 - int(a), s1 - int(a)))
    return {k: np.concatenate(v, axis=0) for k, v in parts.items()}


def load_mq_rank_files(data_dir, device, rank=0, world=1, patterns=_RANK_FILE_PATTERNS):
    """(m_shard, q_shard) from the per-MPI-rank archives the reference's generators leave behind
    (``mq_on_rank<r>.npz`` PODProjector.py:234-237, ``mq_on_proc<r>.npz`` activeSubspaceProjector.py:860-865).  The
    archives are concatenated in rank order (as applications/confusion/confusion_utilities.py:50-58 does) and re-sharded
    over ``world`` GPUs, which need not equal the number of MPI ranks that wrote them."""
    rows = _sharded_rows_from_archives(_rank_files(data_dir, patterns), ("m_data", "q_data"), rank, world)
    return K.to_padded(rows["m_data"], device), K.to_padded(rows["q_data"], device)


def load_jacobian_svd_rank_files(data_dir, device, rank=0, world=1, patterns=("J_on_proc%d.npz", "J_on_rank%d.npz")):
    """Stacked factor rows sigma_i V_i^T (see load_jacobian_svd_factor) from the per-process archives
    ``J_on_proc<r>.npz`` {U_data, sigma_data, V_data} (activeSubspaceProjector.py:867-878).  Returns (Xt, r)."""
    rows = _sharded_rows_from_archives(_rank_files(data_dir, patterns), ("sigma_data", "V_data"), rank, world)
    F = np.ascontiguousarray(np.transpose(rows["V_data"] * rows["sigma_data"][:, None, :], (0, 2, 1)))
    n_loc, r, dM = F.shape
    return K.to_padded(F.reshape(n_loc * r, dM), device), r


# ------------------------------------------------------------------------------------------- reduced Jacobians
def load_reduced_jacobians(file_path, device, kind="JstarPhi", rank=0, world=1):
    """``JstarPhi_data.npz`` {JstarPhi_data (N,dM,rQ), Phi, MPhi} or ``JPsi_data.npz`` {JPsi_data (N,dQ,rM), Psi,
    input_encoder} (dataGenerator.py:649-652).  Returns (block, basis, encoder): ``block`` is this rank's samples as a
    contiguous (N_loc, rows, r) device tensor; the two bases are NumPy arrays (or None when the file holds none)."""
    if kind not in ("JstarPhi", "JPsi"):
        raise ValueError("kind must be 'JstarPhi' or 'JPsi'")
    b_key, e_key = ("Phi", "MPhi") if kind == "JstarPhi" else ("Psi", "input_encoder")
    path = os.path.join(file_path, kind + "_data.npz")
    lo, hi = shard_bounds(_npz_member_shape(path, kind + "_data")[0], rank, world)
    block = torch.as_tensor(np.ascontiguousarray(_npz_member_rows(path, kind + "_data", lo, hi), dtype=np.float64), device=device)
    with np.load(path, allow_pickle=True) as z:                  # the bases are small; members are only read on access
        opt = lambda k: (None if (k not in z.files or z[k].dtype == object) else np.asarray(z[k]))
        return block, opt(b_key), opt(e_key)


def save_reduced_jacobians(file_path, block, basis, encoder, kind="JstarPhi"):
    """Write reduced Jacobians under the reference's names and keys (dataGenerator.py:649-652): ``block`` is
    (N, dM, rQ) = J_i^T (M Phi) for kind='JstarPhi' (modeling.projection.jacobian_transpose_action) or (N, dQ, rM) =
    J_i Psi for kind='JPsi' (jacobian_action)."""
    if kind not in ("JstarPhi", "JPsi"):
        raise ValueError("kind must be 'JstarPhi' or 'JPsi'")
    arr = _dense(block)
    if kind == "JstarPhi":
        np.savez_compressed(os.path.join(file_path, "JstarPhi_data.npz"), JstarPhi_data=arr, Phi=_dense(basis), MPhi=_dense(encoder))
    else:
        np.savez_compressed(os.path.join(file_path, "JPsi_data.npz"), JPsi_data=arr, Psi=_dense(basis), input_encoder=_dense(encoder))


# ------------------------------------------------------------------------------------------- per-sample files -> archives
def _stack_samples(directory, stem, ndata):
    first = np.load(os.path.join(directory, "%s%d.npy" % (stem, 0)))
    out = np.empty((ndata,) + first.shape, dtype=first.dtype)
    out[0] = first
    for i in range(1, ndata):
        out[i] = np.load(os.path.join(directory, "%s%d.npy" % (stem, i)))
    return out


def compress_dataset(file_path, derivatives=(0, 0), clean_up=True, has_z_data=False, input_decoder=None,
                     output_decoder=None, input_encoder=None, output_encoder=None, derivatives_only=False):
    """Same contract as hippyflow/modeling/dataGenerator.py:495-666: gather the per-sample ``.npy`` files under
    ``<file_path>/mq_data/`` (``mzq_data/`` with controls) and ``<file_path>/J_data/`` (``Jz_data/``) into the
    compressed archives ``mq_data.npz`` / ``mzq_data.npz``, ``JstarPhi_data.npz``, ``JPsi_data.npz``, ``Jsvd_data.npz``
    (and the ``Jz*`` variants) that the loaders above and the reference's own loaders read; a Jacobian format is archived
    only if every sample has it, and at least one format must be complete when derivatives are requested.
    ``clean_up`` removes the per-sample directories afterwards.  Returns the list of archives written."""
    import shutil
    sample_dir = os.path.join(file_path, "mzq_data" if has_z_data else "mq_data")
    if derivatives[1] and not has_z_data:
        raise AssertionError("control derivatives need z data")
    indices = [int(f[len("m_sample_"):-4]) for f in os.listdir(sample_dir) if f.startswith("m_sample_") and f.endswith(".npy")]
    if not indices or max(indices) == 0:
        raise RuntimeError("Some issue has arisen, no data found.")            # the reference raises here too (:548-550)
    ndata = max(indices) + 1
    written = []
    if not derivatives_only:
        arrays = {"m_data": _stack_samples(sample_dir, "m_sample_", ndata), "q_data": _stack_samples(sample_dir, "q_sample_", ndata)}
        if has_z_data:
            arrays["z_data"] = _stack_samples(sample_dir, "z_sample_", ndata)
        written.append(os.path.join(file_path, "mzq_data.npz" if has_z_data else "mq_data.npz"))
        np.savez_compressed(written[-1], **arrays)

    def archive_jacobians(jdir, tag, basis_kw):
        # tag '' for parameter Jacobians (J_data/), 'z' for control Jacobians (Jz_data/)
        have = lambda stem: all(os.path.exists(os.path.join(jdir, "%s%d.npy" % (stem, i))) for i in range(ndata))
        formats = {"J%sstarPhi" % tag: ["J%sstarPhi" % tag], "J%sPsi" % tag: ["J%sPsi" % tag],
                   "J%ssvd" % tag: ["U%s_sample_" % tag, "sigma%s_sample_" % tag, "V%s_sample_" % tag]}
        complete = {name: all(have(stem) for stem in stems) for name, stems in formats.items()}
        if not any(complete.values()):
            raise AssertionError("no complete Jacobian format under " + jdir)
        if complete["J%sstarPhi" % tag]:
            written.append(os.path.join(file_path, "J%sstarPhi_data.npz" % tag))
            np.savez_compressed(written[-1], **{"J%sstarPhi_data" % tag: _stack_samples(jdir, "J%sstarPhi" % tag, ndata),
                                                "Phi": basis_kw["output_decoder"], "MPhi": basis_kw["output_encoder"]})
        if complete["J%sPsi" % tag]:
            written.append(os.path.join(file_path, "J%sPsi_data.npz" % tag))
            np.savez_compressed(written[-1], **{"J%sPsi_data" % tag: _stack_samples(jdir, "J%sPsi" % tag, ndata),
                                                "Psi": basis_kw["input_decoder"], "input_encoder": basis_kw["input_encoder"]})
        if complete["J%ssvd" % tag]:
            written.append(os.path.join(file_path, "J%ssvd_data.npz" % tag))
            np.savez_compressed(written[-1], **{"U%s_data" % tag: _stack_samples(jdir, "U%s_sample_" % tag, ndata),
                                                "sigma%s_data" % tag: _stack_samples(jdir, "sigma%s_sample_" % tag, ndata),
                                                "V%s_data" % tag: _stack_samples(jdir, "V%s_sample_" % tag, ndata)})

    bases = dict(input_decoder=input_decoder, output_decoder=output_decoder, input_encoder=input_encoder, output_encoder=output_encoder)
    if derivatives[0]:
        archive_jacobians(os.path.join(file_path, "J_data"), "", bases)
    if derivatives[1]:
        archive_jacobians(os.path.join(file_path, "Jz_data"), "z", bases)
    if clean_up:
        shutil.rmtree(sample_dir)
        if derivatives[0]:
            shutil.rmtree(os.path.join(file_path, "J_data"))
        if derivatives[1]:
            shutil.rmtree(os.path.join(file_path, "Jz_data"))
    return written


def load_projectors(data_dir, names=("POD_projector", "POD_d", "KLE_decoder", "KLE_d", "KLE_projector", "AS_input_projector",
                                     "AS_d_GN", "AS_output_projector", "AS_d_NG")):
    """{name: array} for whichever projector files exist under ``data_dir`` (the names applications/*/…_utilities.py
    ``get_projectors`` reads, plus the KLE decoder name KLEProjector.py:191 writes)."""
    out = {}
    for nm in names:
        f = os.path.join(data_dir, nm + ".npy")
        if os.path.exists(f):
            out[nm] = np.load(f)
    return out
