"""CPU: the collective layer (hippyflow/collectives API) with world_size-2 gloo process groups."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hippyflow_b200.collectives import (CollectiveOperator, MultipleSamePartitioningPDEsCollective,
                                                MultipleSerialPDEsCollective)
        # NCCL rejects strided tensors, gloo does not: make gloo as strict so that the contiguity handling is exercised here
        _ar, _bc = dist.all_reduce, dist.broadcast

        def strict_all_reduce(t, *a, **k):
            assert t.is_contiguous(), "Tensors must be contiguous"
            return _ar(t, *a, **k)

        def strict_broadcast(t, *a, **k):
            assert t.is_contiguous(), "Tensors must be contiguous"
            return _bc(t, *a, **k)

        dist.all_reduce, dist.broadcast = strict_all_reduce, strict_broadcast
        c = MultipleSerialPDEsCollective()
        res = {}
        assert c.size() == world and c.rank() == rank

        class ColumnVector:                         # a DeviceVector: (n, 1) column view of a padded block (stride (2, 1))
            def __init__(self, n, val):
                self.block = torch.zeros(n, 2, dtype=torch.float64)
                self.block[:, 0] = val

            def storage_tensor(self):
                return self.block[:, :1]

        cv = ColumnVector(5, float(rank + 1))
        assert not cv.storage_tensor().is_contiguous()
        assert c.allReduce(cv, "avg") is cv
        res["colvec"] = cv.block.numpy().copy()
        cb = ColumnVector(5, float(rank + 7))
        c.bcast(cb, root=1)
        res["colvec_bcast"] = cb.block.numpy().copy()
        ti = torch.full((3,), 3 * rank + 2, dtype=torch.int64)            # 2 and 5 -> sum 7, 'avg' truncates to 3
        c.allReduce(ti, "avg")
        res["int_avg_tensor"] = ti.numpy().copy()
        ai = np.full(3, 3 * rank + 2, dtype=np.int32)
        c.allReduce(ai, "avg")
        res["int_avg_array"] = ai.copy()
        res["sum_f"] = c.allReduce(float(rank + 1), "sum")
        res["avg_f"] = c.allReduce(float(rank + 1), "AVG")                 # case-insensitive (collective.py:82)
        res["sum_i"] = int(c.allReduce(int(rank + 1), "sum"))
        a = np.arange(6, dtype=np.float64).reshape(2, 3) * (rank + 1)
        out = c.allReduce(a, "avg")
        assert out is a                                                     # in place AND returned
        res["arr"] = a.copy()
        t = torch.full((4, 3), float(rank + 1), dtype=torch.float64)
        c.allReduce(t, "sum")
        res["ten"] = t.numpy().copy()
        nc = torch.zeros(4, 6, dtype=torch.float64)[:, :3]                  # non-contiguous view (padded block)
        nc += rank + 1
        c.allReduce(nc, "avg")
        res["nc"] = nc.clone().numpy()
        b = np.full(3, float(rank))
        c.bcast(b, root=1)
        res["bcast"] = b.copy()
        res["bcast_scalar"] = float(c.bcast(float(rank) + 0.5, root=0))
        try:
            c.allReduce(np.ones(2), "max")
            res["bad_op"] = False
        except NotImplementedError:
            res["bad_op"] = True
        try:
            c.allReduce("a string", "sum")
            res["bad_type"] = False
        except NotImplementedError:
            res["bad_type"] = True

        # CollectiveOperator: local apply then allReduce (collectiveOperator.py:31-38) on host vectors
        from oracle.hippylib_np import Vector

        class LocalOp:
            def init_vector(self, x, dim):
                x.init(3)

            def mult(self, x, y):
                y.set_local((rank + 1) * x.get_local())

            transpmult = mult

        op = CollectiveOperator(LocalOp(), c, "avg")
        x, y = Vector(np.array([1.0, 2.0, 3.0])), Vector(np.zeros(3))
        op.mult(x, y)
        res["op"] = y.get_local()

        # the N > 1 eigensolve path on CPU: each rank holds a sample shard, the repo's CollectiveOperator('avg')
        # over gloo wraps the local operator, Omega is drawn on rank 0 and broadcast (PODProjector.py:360-376)
        from oracle import hippylib_np as hnp
        from oracle import projectors_np as P
        from hippyflow_b200 import synthetic as syn
        M = syn.p1_mass_matrix(8)
        n = M.shape[0]
        u = syn.snapshots(n, 16 * world, r0=20, seed=3)
        shard = u[rank * 16:(rank + 1) * 16]
        local = hnp.LowRankOperator(np.ones(16) / 16, hnp.MultiVector.from_dense(shard.T))
        A = CollectiveOperator(local, c, "avg")
        Om = syn.gaussian_omega(n, 14, seed=4) if rank == 0 else np.zeros((n, 14))
        c.bcast(Om, root=0)
        d, U = hnp.doublePass(A, hnp.MultiVector.from_dense(Om), 8, s=1)
        res["d"], res["Omega00"] = d, float(Om[0, 0])
        res["d_serial"] = P.pod_randomized(u, 8, syn.gaussian_omega(n, 14, seed=4), ranks=1)[0]
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def _shim_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import hippyflow_b200 as hf
        from cpu_device_shim import emulated_device
        from multigpu_worker import sharded_mean_shift_case
        with emulated_device() as dev:
            sharded_mean_shift_case(hf, dev, hf.MultipleSerialPDEsCollective(), rank, world)
        q.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def _split_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import hippyflow_b200 as hf
        mesh_comm, coll_comm = hf.splitCommunicators(None, 2, 2)            # 2 instances x 2 subdomains
        mesh, coll = hf.MultipleSamePartitioningPDEsCollective(mesh_comm), hf.MultipleSerialPDEsCollective(coll_comm)
        res = {"mesh": (mesh.size(), mesh.rank()), "coll": (coll.size(), coll.rank()),
               "mesh_sum": mesh.allReduce(float(rank), "sum"), "coll_avg": coll.allReduce(float(rank), "avg")}
        b = np.array([float(rank)])
        coll.bcast(b, root=1)                                                # root is a rank WITHIN the sample group
        res["coll_bcast"] = float(b[0])
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_split_communicators_grid_gloo_world4():
    """comm_utils.py:19-40: rank r sits in row r // n_subdomain (mesh group, key r % n_subdomain) and column
    r % n_subdomain (sample group, key r // n_subdomain)."""
    world, port = 4, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_split_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert out[r]["mesh"] == (2, r % 2) and out[r]["coll"] == (2, r // 2)
        assert out[r]["mesh_sum"] == float(2 * (r // 2) * 2 + 1)              # ranks {0,1} -> 1, {2,3} -> 5
        assert out[r]["coll_avg"] == float(r % 2) + 1.0                        # ranks {0,2} -> 1, {1,3} -> 2
        assert out[r]["coll_bcast"] == float(2 + r % 2)                        # group rank 1 of column k is world rank 2 + k


def _shim_sharded_worker(rank, world, port, q):
    """Sample-sharded active subspace, collective PODProjector (Omega drawn on rank 0 + bcast) and sharded KLE through the
    CPU test double over gloo: the cases the NCCL worker (tests/multigpu_worker.py) runs on GPUs."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import hippyflow_b200 as hf
        from hippyflow_b200 import synthetic as syn
        from oracle import projectors_np as P
        from cpu_device_shim import emulated_device
        with emulated_device() as dev:
            coll = hf.MultipleSerialPDEsCollective()
            # active subspace from sharded stored Jacobians (activeSubspaceProjector.py:427-463 with 'avg' over ranks)
            J = syn.jacobians(8 * world, 30, 121, r0=16, seed=13)
            OmJ = syn.gaussian_omega(121, 26, seed=14)
            pa = hf.ActiveSubspaceParameterList()
            pa["rank"], pa["oversampling"], pa["verbose"], pa["save_and_plot"] = 16, 10, False, False
            asp = hf.ActiveSubspaceProjector(hf.StoredJacobians(J[rank * 8:(rank + 1) * 8]), None, collective=coll,
                                             parameters=pa, device=dev)
            asp.Omega_GN = OmJ
            dj, Vj, _ = asp.construct_input_subspace(prior_preconditioned=False)
            dj0, Vj0, _ = P.as_input_from_jacobians(J, 16, OmJ, ranks=world)
            k = int(np.sum(dj0 / dj0[0] > 1e-5))
            np.testing.assert_allclose(dj[:k], dj0[:k], rtol=1e-10)
            assert P.principal_angle(hf.mv_to_dense(Vj)[:, :k], Vj0[:, :k]) < 1e-8
            # PODProjector: rank 0 draws Omega, everybody else starts from zeros, bcast (PODProjector.py:367-374)
            M = syn.p1_mass_matrix(12)
            n = M.shape[0]
            u = syn.snapshots(n, 32 * world, r0=24, seed=11)
            params = hf.PODParameterList()
            params["rank"], params["oversampling"], params["verbose"], params["save_and_plot"] = 10, 6, False, False
            pp = hf.PODProjector(hf.StoredSnapshots(u[rank * 32:(rank + 1) * 32].copy()), collective=coll, parameters=params,
                                 device=dev)
            d2, U2 = pp.construct_subspace()
            Om_used = pp.Omega.to_dense()
            Om0 = Om_used.copy()
            coll.bcast(Om0, root=0)
            assert np.array_equal(Om0, Om_used)                               # every rank ended up with rank 0's draw
            d3, U3 = P.pod_randomized(u, 10, Om_used, ranks=world)
            k = int(np.sum(d3 / d3[0] > 1e-5))
            np.testing.assert_allclose(d2[:k], d3[:k], rtol=1e-10)
            assert P.principal_angle(hf.mv_to_dense(U2)[:, :k], U3[:, :k]) < 1e-8
            # KLE 'mass' from sharded draws
            kp = hf.KLEParameterList()
            kp["rank"], kp["oversampling"], kp["verbose"], kp["save_and_plot"] = 8, 6, False, False
            Omk = syn.gaussian_omega(n, 14, seed=15)
            kle = hf.KLEProjector(hf.SampleCovariancePrior(u[rank * 32:(rank + 1) * 32], M, device=dev), collective=coll,
                                  parameters=kp)
            dk, Vk, Ek = kle.construct_input_subspace("mass", Omega=Omk)
            dk0, Vk0, _ = P.kle_from_samples(u, M, 8, Omk, "mass", ranks=world)
            np.testing.assert_allclose(dk, dk0, rtol=1e-10)
            assert P.principal_angle(hf.mv_to_dense(Vk), Vk0, M) < 1e-8
        q.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def test_sharded_projectors_host_logic_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shim_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out == {0: "ok", 1: "ok"}


def test_sharded_mean_shift_host_logic_gloo_world2():
    """The N > 1 host logic of the weighted randomized POD (per-rank provisional means, global mean allreduce, rank-one
    corrections, sketch allreduce) over a world_size-2 gloo group, kernels replaced by the CPU test double."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shim_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == {0: "ok", 1: "ok"}


def test_torch_collective_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):
        r = results[rank]
        assert r["sum_f"] == 3.0 and r["avg_f"] == 1.5 and r["sum_i"] == 3
        np.testing.assert_allclose(r["arr"], np.arange(6).reshape(2, 3) * 1.5)
        np.testing.assert_allclose(r["ten"], 3.0)
        np.testing.assert_allclose(r["nc"], 1.5)
        np.testing.assert_allclose(r["colvec"][:, 0], 1.5)
        np.testing.assert_allclose(r["colvec"][:, 1], 0.0)                    # the padding column is not touched
        np.testing.assert_allclose(r["colvec_bcast"][:, 0], 8.0)
        assert r["int_avg_tensor"].tolist() == [3, 3, 3] and r["int_avg_array"].tolist() == [3, 3, 3]
        np.testing.assert_allclose(r["bcast"], 1.0)
        assert r["bcast_scalar"] == 0.5
        assert r["bad_op"] and r["bad_type"]
        np.testing.assert_allclose(r["op"], 1.5 * np.array([1.0, 2.0, 3.0]))
        np.testing.assert_allclose(r["d"], r["d_serial"], rtol=1e-12)
        assert r["Omega00"] == results[0]["Omega00"]


def test_null_collective_and_operator_protocol():
    from hippyflow_b200.collectives import CollectiveOperator, MatrixMultCollectiveOperator, NullCollective
    c = NullCollective()
    assert c.size() == 1 and c.rank() == 0 and c.bcast(5) == 5
    assert c.allReduce(2.0, "Sum") == 2.0
    with pytest.raises(NotImplementedError):
        c.allReduce(2.0, "prod")

    class NoMult:
        pass

    with pytest.raises(AssertionError):
        CollectiveOperator(NoMult(), c)
    with pytest.raises(AssertionError):
        MatrixMultCollectiveOperator(NoMult(), c)


def test_parameter_lists_have_reference_keys():
    import hippyflow_b200 as hf
    p = hf.PODParameterList()
    assert p["rank"] == 20 and p["oversampling"] == 10 and p["sample_per_process"] == 100
    a = hf.ActiveSubspaceParameterList()
    assert a["rank"] == 128 and a["oversampling"] == 10 and a["samples_per_process"] == 64 and a["serialized_sampling"]
    k = hf.KLEParameterList()
    assert k["rank"] == 128 and k["input_decoder_name"] == "KLE_decoder"
    p["rank"] = 7
    assert p["rank"] == 7
    with pytest.raises(ValueError):
        p["no_such_key"]
