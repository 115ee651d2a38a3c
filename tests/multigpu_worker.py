"""Worker for the multi-GPU parity test (launched by torch.distributed.run, one rank per GPU):
sample-sharded randomized POD / KLE / AS over NCCL must reproduce the single-rank oracle with the same Omega."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sharded_mean_shift_case(hf, dev, coll, rank, world):
    """Sample-sharded weighted POD with a mean 20x the fluctuations, 1024 rows per rank (upload in 4 chunks with a
    per-rank provisional mean + per-chunk lift) and the device-resident implicit shift: both must give the eigenpairs
    of the explicitly shifted full data set."""
    from hippyflow_b200 import _lib as K, synthetic as syn
    M = syn.p1_mass_matrix(64)            # 4225 dofs: the lift runs in row blocks with asynchronous allreduces
    n = M.shape[0]
    per = 1024
    u = syn.snapshots(n, per * world, r0=64, decay=1.0, eps=1e-6, seed=31)
    u = u + 20.0 * np.sqrt(np.mean(u ** 2)) * (1.0 + 0.5 * np.cos(np.arange(n) * 0.02))[None, :]
    # make the shards statistically different so that the per-rank provisional means differ from the global mean
    u[per:] *= 1.5
    Om = syn.gaussian_omega(n, 34, seed=32)
    X = u - u.mean(0)
    Y = X.T @ (X @ (M @ Om)) / X.shape[0]
    Q = Y
    for _ in range(2):
        G = Q.T @ (M @ Q)
        w, V = np.linalg.eigh((G + G.T) / 2)
        Q = Q @ (V / np.sqrt(w))
    W = X @ (M @ Q)
    dd, VV = np.linalg.eigh(W.T @ W / X.shape[0])
    d0, U0 = dd[::-1][:24], Q @ VV[:, ::-1][:, :24]
    k = int(np.sum(d0 / d0[0] > 1e-5))
    shard = u[rank * per:(rank + 1) * per]
    proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
    for entry in ("host", "resident"):
        data = shard.copy() if entry == "host" else K.to_padded(shard, dev)
        d, phi, Mphi, shift = proj.construct_subspace(data, 24, shifted=True, method="randomized", Omega=Om, collective=coll)
        assert proj.shift_route.startswith("pipelined" if entry == "host" else "implicit")   # "implicit (lazy mean redone)" when the mean is large
        np.testing.assert_allclose(d, d0, rtol=1e-10)
        from oracle import projectors_np as P
        assert P.principal_angle(phi[:, :k], U0[:, :k], M) < 1e-8
        np.testing.assert_allclose(shift, u.mean(0), rtol=1e-13)


def peer_exchange_case(hf, dev, coll, rank, world):
    """The fused lift + NVLink exchange (hippyflow_b200/peer.py) against the lift GEMM followed by an NCCL allreduce:
    equal to round-off of the summation order, bitwise identical on all ranks and from call to call; then the operator route (first use verifies itself against NCCL)."""
    from hippyflow_b200 import _lib as K
    from hippyflow_b200.peer import PeerExchange
    n, R, ncols = 66049, 192, 67
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    X = K.to_padded(torch.randn((R, n), generator=g, dtype=torch.float64), dev)
    W = K.to_padded(torch.randn((R, ncols), generator=g, dtype=torch.float64), dev)
    ref = K.dgemm(K.HFB_TN, X, W, alpha=0.5)
    refc = ref.contiguous()
    dist.all_reduce(refc)
    ex = PeerExchange.create(None, dev, n, K._ld(ref), ncols)
    assert ex is not None, "peer exchange unavailable on an NVLink box"
    first = None
    for rep in range(3):
        Y = ex.lift_allreduce(X, W, 0.5)
        torch.cuda.synchronize()
        err = float((Y - refc).abs().max() / refc.abs().max())
        assert err < 1e-13, (rep, err)
        first = Y.clone() if first is None else first
        assert torch.equal(Y, first), "exchange not reproducible from call to call"
    gathered = [torch.empty_like(first.contiguous()) for _ in range(world)]
    dist.all_gather(gathered, first.contiguous())
    for t in gathered:
        assert torch.equal(t, gathered[0]), "ranks hold different bits"
    del Y
    from hippyflow_b200.peer import release_all
    release_all([ex], None)
    # operator route: SampleCovarianceOperator.lift_reduced takes the peer route and verified it against NCCL on first use
    exs = [e for e in getattr(coll, "_peer_exchanges", {}).values() if e is not None]
    assert exs and all(e.verified and e.verify_err < 1e-12 for e in exs), "operator lifts did not take the peer route"


def kle_over_peer_exchange_case(hf, dev, coll, rank, world):
    """KLE from sharded parameter draws at n >= 4096 (the lifts take the fused NVLink exchange): 'mass' (doublePassG on M C M,
    un-centred operator: the non-lazy branch of the peer route) and 'identity' with faithful=True (T = (A Q)^T Q: two
    exchanges per solve, so the sketch of the first must survive the second -- the alternating result blocks)."""
    from hippyflow_b200 import synthetic as syn
    from oracle import projectors_np as P
    M = syn.p1_mass_matrix(66)                                     # 4489 dofs
    n = M.shape[0]
    per = 48
    m_data = syn.snapshots(n, per * world, r0=40, seed=41)
    Om = syn.gaussian_omega(n, 30, seed=42)
    pk = hf.KLEParameterList()
    pk["rank"], pk["oversampling"], pk["verbose"], pk["save_and_plot"] = 20, 10, False, False
    proj = hf.KLEProjector(hf.SampleCovariancePrior(m_data[rank * per:(rank + 1) * per].copy(), M, device=dev), collective=coll,
                           parameters=pk)
    for orth, faithful in (("mass", False), ("identity", True), ("mass", True)):
        d, dec, enc = proj.construct_input_subspace(orth, Omega=Om, faithful=faithful)
        d0, V0, E0 = P.kle_from_samples(m_data, M, 20, Om, orth, ranks=world)
        k = int(np.sum(d0 / d0[0] > 1e-5))
        np.testing.assert_allclose(d[:k], d0[:k], rtol=1e-10)
        assert P.principal_angle(hf.mv_to_dense(dec)[:, :k], V0[:, :k], M if orth == "mass" else None) < 1e-8, (orth, faithful)
        if orth == "mass":
            V, E = hf.mv_to_dense(dec), hf.mv_to_dense(enc)
            assert np.linalg.norm(M @ V - E) / np.linalg.norm(E) < 1e-10          # test_KLEProjector.py:102-108


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import hippyflow_b200 as hf
    from hippyflow_b200 import synthetic as syn
    from oracle import projectors_np as P
    coll = hf.MultipleSerialPDEsCollective()
    assert coll.size() == world and coll.rank() == rank

    nx = 24
    M = syn.p1_mass_matrix(nx)
    n = M.shape[0]
    N = 64 * world
    u = syn.snapshots(n, N, r0=48, seed=11) + 0.2 * np.cos(np.linspace(0, 2, n))[None, :]
    Om = syn.gaussian_omega(n, 30, seed=12)
    shard = u[rank * 64:(rank + 1) * 64]

    # weighted randomized POD, shifted (global mean needs an allreduce)
    proj = hf.PODProjectorFromData(None, M_output=M, device=dev)
    d, phi, Mphi, shift = proj.construct_subspace(shard.copy(), 20, shifted=True, method="randomized", Omega=Om, collective=coll)
    d0, U0, E0, s0 = P.pod_randomized_weighted(u, M, 20, Om, shifted=True, ranks=world)
    k = int(np.sum(d0 / d0[0] > 1e-5))
    assert np.abs(d[:k] - d0[:k]).max() / 1.0 < 1e-10 * d0[0] or np.allclose(d[:k], d0[:k], rtol=1e-10), (d, d0)
    np.testing.assert_allclose(d[:k], d0[:k], rtol=1e-10)
    assert P.principal_angle(phi[:, :k], U0[:, :k], M) < 1e-8
    np.testing.assert_allclose(shift, s0, atol=1e-14)

    # PODProjector: Omega drawn on rank 0 and broadcast (PODProjector.py:367-374) -> identical on all ranks
    params = hf.PODParameterList()
    params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = 20, 10, False, False
    pp = hf.PODProjector(hf.StoredSnapshots(shard.copy()), collective=coll, parameters=params, device=dev)
    d2, U2 = pp.construct_subspace()
    Om_used = pp.Omega.to_dense()
    gathered = [torch.zeros_like(pp.Omega.tensor().contiguous()) for _ in range(world)]
    dist.all_gather(gathered, pp.Omega.tensor().contiguous())
    for g in gathered:
        assert torch.equal(g, gathered[0])
    d3, U3 = P.pod_randomized(u, 20, Om_used, ranks=world)
    k = int(np.sum(d3 / d3[0] > 1e-5))
    np.testing.assert_allclose(d2[:k], d3[:k], rtol=1e-10)
    assert P.principal_angle(hf.mv_to_dense(U2)[:, :k], U3[:, :k]) < 1e-8

    # active subspace from sharded stored Jacobians
    J = syn.jacobians(8 * world, 30, 121, r0=16, seed=13)
    OmJ = syn.gaussian_omega(121, 26, seed=14)
    pa = hf.ActiveSubspaceParameterList()
    pa["rank"], pa["oversampling"], pa["verbose"], pa["save_and_plot"] = 16, 10, False, False
    asp = hf.ActiveSubspaceProjector(hf.StoredJacobians(J[rank * 8:(rank + 1) * 8]), None, collective=coll, parameters=pa, device=dev)
    asp.Omega_GN = OmJ
    dj, Vj, _ = asp.construct_input_subspace(prior_preconditioned=False)
    dj0, Vj0, _ = P.as_input_from_jacobians(J, 16, OmJ, ranks=world)
    k = int(np.sum(dj0 / dj0[0] > 1e-5))
    np.testing.assert_allclose(dj[:k], dj0[:k], rtol=1e-10)
    assert P.principal_angle(hf.mv_to_dense(Vj)[:, :k], Vj0[:, :k]) < 1e-8

    # n >= 4096: the lift GEMM is cut into row blocks whose NCCL allreduce overlaps the next block's GEMM
    Mb = syn.p1_mass_matrix(70)                                   # 5041 dofs
    nb = Mb.shape[0]
    ub = syn.snapshots(nb, 32 * world, r0=24, seed=21)
    Omb = syn.gaussian_omega(nb, 18, seed=22)
    pb = hf.PODProjectorFromData(None, M_output=Mb, device=dev)
    db, phib, _, sb = pb.construct_subspace(ub[rank * 32:(rank + 1) * 32].copy(), 8, shifted=True, method="randomized",
                                            Omega=Omb, collective=coll)
    db0, Ub0, _, sb0 = P.pod_randomized_weighted(ub, Mb, 8, Omb, shifted=True, ranks=world)
    np.testing.assert_allclose(db, db0, rtol=1e-10)
    assert P.principal_angle(phib, Ub0, Mb) < 1e-8
    np.testing.assert_allclose(sb, sb0, atol=1e-14)
    # same through the device-resident (non-pipelined) entry
    Xd = hf._lib.to_padded(ub[rank * 32:(rank + 1) * 32], dev)
    db2, _, _, _ = pb.construct_subspace(Xd, 8, shifted=True, method="randomized", Omega=Omb, collective=coll)
    np.testing.assert_allclose(db2, db0, rtol=1e-10)

    sharded_mean_shift_case(hf, dev, coll, rank, world)
    kle_over_peer_exchange_case(hf, dev, coll, rank, world)
    peer_exchange_case(hf, dev, coll, rank, world)

    # collective on device blocks: one NCCL call for the whole padded block, 'avg' = sum / size
    mv = hf.DeviceMultiVector(50, 7, device=dev)
    mv.tensor().fill_(float(rank + 1))
    coll.allReduce(mv, "avg")
    assert torch.allclose(mv.tensor(), torch.full_like(mv.tensor(), (world + 1) / 2.0))
    assert coll.allReduce(float(rank), "sum") == sum(range(world))
    # strided operands under NCCL (advisor r1): a DeviceVector is an (n, 1) column view of a padded block, mv[j] a column
    # of a multivector -- NCCL rejects non-contiguous tensors, the collective must pack / unpack them
    dv = hf.DeviceVector(37, dev)
    dv.set_local(np.full(37, float(rank + 1)))
    assert not dv.storage_tensor().is_contiguous()
    coll.allReduce(dv, "sum")
    np.testing.assert_allclose(dv.get_local(), sum(range(1, world + 1)))
    mv.tensor().fill_(float(rank))
    col = mv[3]
    coll.allReduce(col, "avg")
    assert torch.allclose(mv.tensor()[:, 3], torch.full((50,), (world - 1) / 2.0, dtype=torch.float64, device=dev))
    assert torch.allclose(mv.tensor()[:, 2], torch.full((50,), float(rank), dtype=torch.float64, device=dev))
    bv = hf.DeviceVector(11, dev)
    bv.set_local(np.full(11, float(rank + 5)))
    coll.bcast(bv, root=world - 1)
    np.testing.assert_allclose(bv.get_local(), float(world + 4))
    # operator applies on device VECTORS through the collective (CollectiveOperator.mult path, collectiveOperator.py:31-38)
    from hippyflow_b200.linalg import SampleCovariance
    cov_op = hf.SampleCovarianceOperator(SampleCovariance(hf._lib.to_padded(shard, dev)), coll, "avg")
    xv, yv = hf.DeviceVector(n, dev), hf.DeviceVector(n, dev)
    xvec = np.cos(np.arange(n) * 0.1)
    xv.set_local(xvec)
    cov_op.mult(xv, yv)
    np.testing.assert_allclose(yv.get_local(), u.T @ (u @ xvec) / u.shape[0], rtol=1e-11, atol=1e-13)
    dist.barrier()
    if rank == 0:
        print("MULTIGPU_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
