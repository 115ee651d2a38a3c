"""Golden vectors for the small operator classes around the stored-data path, from the UNMODIFIED reference
(/root/reference through oracle/ref_import.py) on seeded inputs.  TEST INFRASTRUCTURE ONLY; run in the build container:

    python -m oracle.make_golden_small_ops        ->  tests/golden/small_operators_ref.npz

Verbatim reference code exercised: LowRankRectangularOperator.mult / transpmult (hippyflow/modeling/
lowRankRectangularOperator.py:50-66), PriorPreconditionedProjector.mult (priorPreconditionedProjector.py:48-55),
npToDolfinOperator.mult / transpmult (operatorWrappers.py:38-52).  The vectors / multivectors they act on are the NumPy
stand-ins of oracle/hippylib_np.py (dolfin and hIPPYlib are not installed here)."""
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


def main():
    warnings.simplefilter("ignore")
    hf = import_reference()
    import dolfin
    dolfin.Vector = lambda comm=None: hnp.Vector()            # PriorPreconditionedProjector.__init__ builds a work vector
    rng = np.random.default_rng(11)
    out = {}

    # ---- LowRankRectangularOperator: A = U s V^T with the truncated SVD of a synthetic Jacobian (100 x 121, rank 12)
    J = syn.jacobians(1, 100, 121, r0=32, seed=9)[0]
    Uf, sf, Vtf = np.linalg.svd(J, full_matrices=False)
    r = 12
    U, s, V = Uf[:, :r].copy(), sf[:r].copy(), Vtf[:r].T.copy()
    op = hf.LowRankRectangularOperator(hnp.MultiVector.from_dense(U), s, hnp.MultiVector.from_dense(V),
                                       U_init_vector=lambda x: x.init(100), V_init_vector=lambda x: x.init(121))
    xs, ws = rng.standard_normal((4, 121)), rng.standard_normal((4, 100))
    ys, zs = [], []
    for x, w in zip(xs, ws):
        xv, yv = hnp.Vector(x.copy()), hnp.Vector(np.full(100, 3.0))
        op.mult(xv, yv)
        ys.append(yv.get_local())
        wv, zv = hnp.Vector(w.copy()), hnp.Vector(np.full(121, -2.0))
        op.transpmult(wv, zv)
        zs.append(zv.get_local())
    out.update(lr_U=U, lr_s=s, lr_V=V, lr_x=xs, lr_w=ws, lr_mult=np.array(ys), lr_transpmult=np.array(zs))

    # ---- PriorPreconditionedProjector: y = U U^T C^-1 x with C^-1 = a P1 mass matrix (121 dofs), U = C-orthonormal basis
    M = syn.p1_mass_matrix(10)
    n = M.shape[0]
    Q = np.linalg.qr(rng.standard_normal((n, 9)))[0]
    Lc = np.linalg.cholesky(Q.T @ (M @ Q))
    Ub = Q @ np.linalg.inv(Lc).T                               # U^T M U = I

    class _Cinv(hnp.SparseOperator):
        def mpi_comm(self):
            return None

    proj = hf.PriorPreconditionedProjector(hnp.MultiVector.from_dense(Ub), _Cinv(M), lambda x, dim: x.init(n))
    px = rng.standard_normal((4, n))
    py = []
    for x in px:
        xv, yv = hnp.Vector(x.copy()), hnp.Vector(np.ones(n))
        proj.mult(xv, yv)
        py.append(yv.get_local())
    out.update(pp_U=Ub, pp_nx=10, pp_x=px, pp_mult=np.array(py))

    # ---- npToDolfinOperator
    A = rng.standard_normal((17, 29))
    dop = hf.npToDolfinOperator(A)
    dx, dw = rng.standard_normal(29), rng.standard_normal(17)
    yv, zv = hnp.Vector(), hnp.Vector()
    dop.init_vector(yv, 0)
    dop.init_vector(zv, 1)
    dop.mult(hnp.Vector(dx.copy()), yv)
    dop.transpmult(hnp.Vector(dw.copy()), zv)
    out.update(np_A=A, np_x=dx, np_w=dw, np_mult=yv.get_local(), np_transpmult=zv.get_local())

    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "small_operators_ref.npz"), **out)
    print("small_operators_ref.npz", os.path.getsize(os.path.join(OUT, "small_operators_ref.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
