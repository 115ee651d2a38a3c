"""Golden vectors for the composite operator classes of the active-subspace path, from the UNMODIFIED reference
(/root/reference through oracle/ref_import.py) on seeded inputs.  TEST INFRASTRUCTURE ONLY; run in the build container:

    python -m oracle.make_golden_list_ops        ->  tests/golden/list_operators_ref.npz

Verbatim reference code exercised: JTJ.mult / JJT.mult (hippyflow/modeling/jacobian.py:142-193), SummedListOperator.mult
(activeSubspaceProjector.py:69-95; averaged and summed, with a zero and with a non-zero result vector on entry),
MatrixMultCollectiveOperator.matMvMult (hippyflow/collectives/collectiveOperator.py:73-80) and CollectiveOperator.mult
(:31-38) over NullCollective (collective.py:19-38).  The Jacobians are dense NumPy stand-ins (oracle/hippylib_np.DenseOperator:
mult = J x, transpmult = J^T y); vectors / multivectors are the NumPy stand-ins of oracle/hippylib_np.py."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import hippylib_np as hnp  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402
from hippyflow_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class _Jac(hnp.DenseOperator):
    def mpi_comm(self):
        return None


class _StackedJTJ:
    """A local operator with the block protocol MatrixMultCollectiveOperator needs: Y[j] = mean_i J_i^T J_i X[j]."""

    def __init__(self, J):
        self.J = J

    def matMvMult(self, X, Y):
        Xd = X.to_dense()
        Yd = sum(Ji.T @ (Ji @ Xd) for Ji in self.J) / len(self.J)
        for j in range(Y.nvec()):
            Y[j].set_local(Yd[:, j])


def main():
    warnings.simplefilter("ignore")
    hf = import_reference()
    import dolfin
    dolfin.Vector = lambda arg=None: hnp.Vector(arg.get_local().copy()) if isinstance(arg, hnp.Vector) else hnp.Vector()
    rng = np.random.default_rng(23)
    N, dQ, dM = 5, 14, 22
    J = syn.jacobians(N, dQ, dM, r0=8, seed=4)
    xs = rng.standard_normal((3, dM))
    ws = rng.standard_normal((3, dQ))
    out = dict(J=J, x=xs, w=ws)

    # ---- JTJ / JJT of one Jacobian
    jtj, jjt = hf.JTJ(_Jac(J[0])), hf.JJT(_Jac(J[0]))
    a, b = [], []
    for x, w in zip(xs, ws):
        yv = hnp.Vector(np.full(dM, 5.0))
        jtj.mult(hnp.Vector(x.copy()), yv)
        a.append(yv.get_local())
        zv = hnp.Vector(np.full(dQ, -1.0))
        jjt.mult(hnp.Vector(w.copy()), zv)
        b.append(zv.get_local())
    out.update(jtj_mult=np.array(a), jjt_mult=np.array(b))

    # ---- SummedListOperator over the per-sample J_i^T J_i, averaged and summed; y = 0 on entry (what every caller on the path
    # passes) and y != 0 on entry (the reference seeds its accumulator with a copy of y, activeSubspaceProjector.py:84-85)
    ops = [hf.JTJ(_Jac(J[i])) for i in range(N)]
    for tag, average in (("avg", True), ("sum", False)):
        op = hf.modeling.activeSubspaceProjector.SummedListOperator(ops, average=average)
        zero_in, seeded_in = [], []
        for x in xs:
            yv = hnp.Vector(np.zeros(dM))
            op.mult(hnp.Vector(x.copy()), yv)
            zero_in.append(yv.get_local())
            yv = hnp.Vector(np.full(dM, 2.0))
            op.mult(hnp.Vector(x.copy()), yv)
            seeded_in.append(yv.get_local())
        out["sl_%s_zero" % tag] = np.array(zero_in)
        out["sl_%s_seeded" % tag] = np.array(seeded_in)

    # ---- CollectiveOperator / MatrixMultCollectiveOperator over NullCollective
    coll = hf.NullCollective()
    cop = hf.CollectiveOperator(hf.modeling.activeSubspaceProjector.SummedListOperator(ops, average=True), coll, mpi_op="avg")
    c = []
    for x in xs:
        yv = hnp.Vector(np.zeros(dM))
        cop.mult(hnp.Vector(x.copy()), yv)
        c.append(yv.get_local())
    out["coll_mult"] = np.array(c)
    mop = hf.MatrixMultCollectiveOperator(_StackedJTJ(J), coll, mpi_op="avg")
    X = hnp.MultiVector.from_dense(xs.T.copy())
    Y = hnp.MultiVector.from_dense(np.zeros((dM, 3)))
    mop.matMvMult(X, Y)
    out["mm_matMvMult"] = Y.to_dense()

    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "list_operators_ref.npz"), **out)
    print("list_operators_ref.npz", os.path.getsize(os.path.join(OUT, "list_operators_ref.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
