"""On-disk data contract, host side only (no GPU): archive names, keys, shapes, sharding and
per-rank re-sharding.  Arrays are "uploaded" to the CPU device here; the upload to a GPU is the same torch call.
When /root/reference is present (build container) the reference's own compress_dataset is run verbatim under the
stub importer on the same per-sample files and the archives are compared bit for bit."""
import os

import numpy as np
import pytest
import torch

from hippyflow_b200 import dataIO

CPU = torch.device("cpu")


def _write_samples(root, ndata=7, dM=11, dQ=5, r=3, rQ=4, rM=2, z=False, seed=0):
    rng = np.random.default_rng(seed)
    sdir = os.path.join(root, "mzq_data" if z else "mq_data")
    os.makedirs(sdir)
    os.makedirs(os.path.join(root, "J_data"))
    data = dict(m=rng.standard_normal((ndata, dM)), q=rng.standard_normal((ndata, dQ)), z=rng.standard_normal((ndata, 3)),
                JstarPhi=rng.standard_normal((ndata, dM, rQ)), JPsi=rng.standard_normal((ndata, dQ, rM)),
                U=rng.standard_normal((ndata, dQ, r)), sigma=rng.random((ndata, r)), V=rng.standard_normal((ndata, dM, r)))
    for i in range(ndata):
        np.save(os.path.join(sdir, "m_sample_%d.npy" % i), data["m"][i])
        np.save(os.path.join(sdir, "q_sample_%d.npy" % i), data["q"][i])
        if z:
            np.save(os.path.join(sdir, "z_sample_%d.npy" % i), data["z"][i])
        np.save(os.path.join(root, "J_data", "JstarPhi%d.npy" % i), data["JstarPhi"][i])
        np.save(os.path.join(root, "J_data", "JPsi%d.npy" % i), data["JPsi"][i])
        np.save(os.path.join(root, "J_data", "U_sample_%d.npy" % i), data["U"][i])
        np.save(os.path.join(root, "J_data", "sigma_sample_%d.npy" % i), data["sigma"][i])
        np.save(os.path.join(root, "J_data", "V_sample_%d.npy" % i), data["V"][i])
    return data


def test_compress_dataset_archives_and_loaders(tmp_path):
    root = str(tmp_path) + "/"
    data = _write_samples(root)
    rng = np.random.default_rng(1)
    Phi, MPhi, Psi, enc = (rng.standard_normal(s) for s in ((5, 4), (5, 4), (11, 2), (11, 2)))
    written = dataIO.compress_dataset(root, derivatives=(1, 0), clean_up=True, input_decoder=Psi, output_decoder=Phi,
                                      input_encoder=enc, output_encoder=MPhi)
    assert sorted(os.path.basename(w) for w in written) == ["JPsi_data.npz", "JstarPhi_data.npz", "Jsvd_data.npz", "mq_data.npz"]
    assert not os.path.exists(root + "mq_data") and not os.path.exists(root + "J_data")       # clean_up
    with np.load(root + "mq_data.npz") as z:
        assert np.array_equal(z["m_data"], data["m"]) and np.array_equal(z["q_data"], data["q"])
    # sample-sharded loads: rank r of 2 gets rows [3r, 3r+3) (equal shards, remainder dropped from the tail)
    for r in range(2):
        m, q = dataIO.load_mq_data(root, CPU, rank=r, world=2)
        assert np.array_equal(m.numpy(), data["m"][3 * r:3 * r + 3]) and np.array_equal(q.numpy(), data["q"][3 * r:3 * r + 3])
        blk, B, E = dataIO.load_reduced_jacobians(root, CPU, "JstarPhi", rank=r, world=2)
        assert np.array_equal(blk.numpy(), data["JstarPhi"][3 * r:3 * r + 3]) and np.array_equal(B, Phi) and np.array_equal(E, MPhi)
        blk, B, E = dataIO.load_reduced_jacobians(root, CPU, "JPsi", rank=r, world=2)
        assert np.array_equal(blk.numpy(), data["JPsi"][3 * r:3 * r + 3]) and np.array_equal(B, Psi) and np.array_equal(E, enc)
    Xt, blk = dataIO.load_jacobian_svd_factor(root, CPU)
    assert blk == 3 and tuple(Xt.shape) == (7 * 3, 11)
    np.testing.assert_array_equal(Xt.numpy().reshape(7, 3, 11), np.transpose(data["V"] * data["sigma"][:, None, :], (0, 2, 1)))
    with pytest.raises(ValueError):
        dataIO.load_reduced_jacobians(root, CPU, "Jfoo")


def test_compress_dataset_partial_formats_and_errors(tmp_path):
    root = str(tmp_path) + "/"
    _write_samples(root, ndata=4)
    os.remove(root + "J_data/JPsi2.npy")                     # one sample lacks JPsi -> that archive is skipped
    written = dataIO.compress_dataset(root, derivatives=(1, 0), clean_up=False)
    names = sorted(os.path.basename(w) for w in written)
    assert names == ["JstarPhi_data.npz", "Jsvd_data.npz", "mq_data.npz"]
    blk, B, E = dataIO.load_reduced_jacobians(root, CPU, "JstarPhi")
    assert B is None and E is None and blk.shape[0] == 4     # bases were not supplied (the reference stores None)
    with pytest.raises(AssertionError):
        dataIO.compress_dataset(root, derivatives=(0, 1), clean_up=False)          # control derivatives without z data
    empty = str(tmp_path / "empty") + "/"
    os.makedirs(empty + "mq_data")
    with pytest.raises(RuntimeError):
        dataIO.compress_dataset(empty)


def test_per_rank_archives_are_resharded(tmp_path):
    rng = np.random.default_rng(2)
    root = str(tmp_path)
    ms = [rng.standard_normal((c, 6)) for c in (4, 4, 5)]
    qs = [rng.standard_normal((c, 3)) for c in (4, 4, 5)]
    for r, (m, q) in enumerate(zip(ms, qs)):
        np.savez_compressed(os.path.join(root, "mq_on_rank%d.npz" % r), m_data=m, q_data=q)
    M, Q = np.concatenate(ms), np.concatenate(qs)
    for world in (1, 2, 4):
        per = 13 // world
        for r in range(world):
            m, q = dataIO.load_mq_rank_files(root, CPU, rank=r, world=world)
            assert np.array_equal(m.numpy(), M[r * per:(r + 1) * per]) and np.array_equal(q.numpy(), Q[r * per:(r + 1) * per])
    # the spelling the application loaders use, and the Jacobian archives of the active-subspace generator
    sub = tmp_path / "apps"
    sub.mkdir()
    np.savez_compressed(str(sub / "mq_on_rank_0.npz"), m_data=ms[0], q_data=qs[0])
    m, _ = dataIO.load_mq_rank_files(str(sub), CPU)
    assert np.array_equal(m.numpy(), ms[0])
    U, s, V = rng.standard_normal((5, 3, 2)), rng.random((5, 2)), rng.standard_normal((5, 6, 2))
    np.savez_compressed(str(sub / "J_on_proc0.npz"), U_data=U[:2], sigma_data=s[:2], V_data=V[:2])
    np.savez_compressed(str(sub / "J_on_proc1.npz"), U_data=U[2:], sigma_data=s[2:], V_data=V[2:])
    Xt, blk = dataIO.load_jacobian_svd_rank_files(str(sub), CPU, rank=1, world=2)      # samples 2, 3 (5 // 2 = 2 per rank)
    np.testing.assert_array_equal(Xt.numpy().reshape(2, 2, 6), np.transpose(V[2:4] * s[2:4, None, :], (0, 2, 1)))
    with pytest.raises(FileNotFoundError):
        dataIO.load_mq_rank_files(str(tmp_path / "nothing_here"), CPU)


def test_projector_files_round_trip(tmp_path):
    root = str(tmp_path)
    rng = np.random.default_rng(3)
    U, d = rng.standard_normal((9, 4)), np.arange(4.0)
    dataIO.save_pod(root, U, d)
    dataIO.save_kle(root, U, d)
    dataIO.save_active_subspace(root, U, d, 128)
    got = dataIO.load_projectors(root)
    assert set(got) == {"POD_projector", "POD_d", "KLE_decoder", "KLE_d"}
    assert np.array_equal(got["POD_projector"], U) and np.array_equal(got["KLE_d"], d)
    assert np.array_equal(np.load(os.path.join(root, "AS_128_input_decoder.npy")), U)
    blk = rng.standard_normal((3, 9, 4))
    dataIO.save_reduced_jacobians(root, blk, U[:, :4], U[:, :4], kind="JstarPhi")
    back, B, E = dataIO.load_reduced_jacobians(root, CPU, "JstarPhi")
    assert np.array_equal(back.numpy(), blk) and np.array_equal(B, U[:, :4])


def test_compress_dataset_matches_reference_bit_for_bit(tmp_path):
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present (GPU box)")
    hf_ref = ref_import.import_reference()
    from hippyflow.modeling.dataGenerator import compress_dataset as ref_compress
    rng = np.random.default_rng(4)
    Phi, MPhi, Psi, enc = (rng.standard_normal(s) for s in ((5, 4), (5, 4), (11, 2), (11, 2)))
    ours, theirs = str(tmp_path / "ours") + "/", str(tmp_path / "theirs") + "/"
    for root in (ours, theirs):
        os.makedirs(root)
        _write_samples(root, seed=9)
    dataIO.compress_dataset(ours, derivatives=(1, 0), clean_up=False, input_decoder=Psi, output_decoder=Phi,
                            input_encoder=enc, output_encoder=MPhi)
    ref_compress(theirs, derivatives=(1, 0), clean_up=False, input_decoder=Psi, output_decoder=Phi,
                 input_encoder=enc, output_encoder=MPhi)
    for name in ("mq_data.npz", "JstarPhi_data.npz", "JPsi_data.npz", "Jsvd_data.npz"):
        with np.load(ours + name) as a, np.load(theirs + name) as b:
            assert sorted(a.files) == sorted(b.files), name
            for k in a.files:
                assert np.array_equal(a[k], b[k]), (name, k)


def test_streamed_shard_reader_matches_numpy(tmp_path):
    """dataIO._npz_member_rows / _npz_member_shape: rows [lo, hi) of an archive member streamed from the zip file equal
    ``np.load(path)[key][lo:hi]`` for deflated and stored archives, any dtype, 1-/2-/3-D members, empty and clipped ranges;
    Fortran-ordered, scalar and object members take the NumPy route; shapes come from the member header alone."""
    import numpy as np
    from hippyflow_b200 import dataIO
    rng = np.random.default_rng(0)
    a = rng.standard_normal((301, 57, 3))
    b = rng.integers(0, 9, (301, 5)).astype(np.int32)
    v = rng.standard_normal(301).astype(np.float32)
    f = np.asfortranarray(rng.standard_normal((301, 4)))
    for comp in (True, False):
        fn = str(tmp_path / ("x%d.npz" % comp))
        (np.savez_compressed if comp else np.savez)(fn, a=a, b=b, v=v, f=f, s=np.float64(3.5), o=np.array(None, dtype=object))
        for lo, hi in ((0, None), (0, 0), (17, 18), (100, 301), (250, 400), (301, 301)):
            for k, ref in (("a", a), ("b", b), ("v", v), ("f", f)):
                got = dataIO._npz_member_rows(fn, k, lo, hi, chunk_bytes=4096)
                assert got.dtype == ref.dtype and np.array_equal(got, ref[lo:hi]), (comp, k, lo, hi)
        assert dataIO._npz_member_shape(fn, "a") == (301, 57, 3) and dataIO._npz_member_shape(fn, "s") == ()
        assert float(dataIO._npz_member_rows(fn, "s")) == 3.5
        assert dataIO._npz_member_rows(fn, "o").dtype == object
    # a truncated archive is an error, not a short shard
    raw = open(str(tmp_path / "x0.npz"), "rb").read()
    import zipfile
    with zipfile.ZipFile(str(tmp_path / "t.npz"), "w") as zf:
        import io
        buf = io.BytesIO()
        np.lib.format.write_array(buf, a)
        zf.writestr("a.npy", buf.getvalue()[:-1000])
    with pytest.raises(EOFError):
        dataIO._npz_member_rows(str(tmp_path / "t.npz"), "a", 290, 301)
