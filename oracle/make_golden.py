"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, imported through
oracle/ref_import.py) on small seeded inputs.  TEST INFRASTRUCTURE ONLY; run in the build container:

    python -m oracle.make_golden

What is verbatim reference code and what is restated:
  * PODProjectorFromData.construct_subspace (hep / ghep / inverse_ghep, shifted or not) -- verbatim
    (hippyflow/modeling/PODProjector.py:699-852), SciPy ARPACK / LAPACK / SuperLU underneath;
  * MeanJTJfromDataOperator.mult -- verbatim (hippyflow/modeling/operatorWrappers.py:95-114);
  * CollectiveOperator + NullCollective -- verbatim (hippyflow/collectives/);
  * doublePass / doublePassG / MultiVector / LowRankOperator -- oracle/hippylib_np.py (hIPPYlib itself is a
    third-party dependency that is not available; see that file's header).
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import hippylib_np as hnp  # noqa: E402
from oracle import projectors_np as P  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402
from hippyflow_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    warnings.simplefilter("ignore")
    hf = import_reference()
    os.makedirs(OUT, exist_ok=True)

    # ---------------- (1) deterministic weighted POD, reference verbatim (cfg1: 16x16 mesh, 289 dofs, 100 samples, rank 15)
    nx, N, r = 16, 100, 15
    M = syn.p1_mass_matrix(nx)
    n = M.shape[0]
    u_data = syn.snapshots(n, N, r0=40, decay=1.0, eps=1e-6, seed=0)
    u_data += 0.3 * np.sin(np.linspace(0, 3, n))[None, :]      # non-zero mean so that 'shifted' matters
    pod = object.__new__(hf.PODProjectorFromData)              # __init__ needs FEniCS only to assemble M
    pod.M_csr = M
    out = {"u_data": u_data, "nx": nx, "rank": r}
    for method in ("hep", "ghep", "inverse_ghep"):
        for shifted in (True, False):
            with contextlib.redirect_stdout(io.StringIO()):
                d, phi, Mphi, shift = pod.construct_subspace(u_data.copy(), r, shifted=shifted, method=method)
            key = f"{method}_{int(shifted)}"
            out[key + "_d"], out[key + "_phi"], out[key + "_Mphi"], out[key + "_shift"] = d, phi, Mphi, shift
    np.savez_compressed(os.path.join(OUT, "pod_from_data_ref.npz"), **out)

    # ---------------- (2) MeanJTJfromDataOperator.mult, reference verbatim (cfg1 AS: 64 samples, 100 obs, 121 params)
    Nj, dQ, dM = 64, 100, 121
    J = syn.jacobians(Nj, dQ, dM, r0=32, seed=3)
    rng = np.random.default_rng(7)
    xs = rng.standard_normal((5, dM))
    Ginv = np.linalg.inv(0.5 * np.eye(dQ) + 0.1 * np.cov(rng.standard_normal((dQ, 4 * dQ))))

    class _R:
        def init_vector(self, x, dim):
            x.init(dM)

    class _Prior:
        R = _R()

    ys, ysG = [], []
    for G, acc in ((None, ys), (Ginv, ysG)):
        op = hf.MeanJTJfromDataOperator(J, _Prior(), G)
        for x in xs:
            xv, yv = hnp.Vector(x.copy()), hnp.Vector(np.zeros(dM))
            op.mult(xv, yv)
            acc.append(yv.get_local())
    np.savez_compressed(os.path.join(OUT, "meanjtj_ref.npz"), J=J, xs=xs, noise_cov_inv=Ginv, ys=np.array(ys), ysG=np.array(ysG))

    # ---------------- (3) double-pass eigensolves driven through the reference's CollectiveOperator(NullCollective)
    # POD (PODProjector.py:359-376): LowRankOperator(ones/N) -> CollectiveOperator('avg') -> doublePass
    k, p = 15, 10
    Om = syn.gaussian_omega(n, k + p, seed=1)
    U_loc = hnp.MultiVector.from_dense(u_data.T)
    A = hf.CollectiveOperator(hnp.LowRankOperator(np.ones(N) / N, U_loc), hf.NullCollective(), mpi_op="avg")
    d_pod, U_pod = hnp.doublePass(A, hnp.MultiVector.from_dense(Om), k, s=1)
    # weighted POD / sample KLE 'mass' (KLEProjector.py:146-168): doublePassG(M C M, M, Msolver)
    Aw = hf.CollectiveOperator(P.SandwichedCovarianceOperator(u_data, M), hf.NullCollective(), mpi_op="avg")
    B = hnp.SparseOperator(M)
    d_w, U_w = hnp.doublePassG(Aw, B, B, hnp.MultiVector.from_dense(Om), k, s=1)
    # AS input (activeSubspaceProjector.py:427-463) on the reference's own MeanJTJfromDataOperator
    ka = 64
    Om_as = syn.gaussian_omega(dM, ka + p, seed=2)
    Aj = hf.CollectiveOperator(hf.MeanJTJfromDataOperator(J, _Prior(), None), hf.NullCollective(), mpi_op="avg")
    d_as, V_as = hnp.doublePass(Aj, hnp.MultiVector.from_dense(Om_as), ka, s=1)
    Mp = syn.p1_mass_matrix(10)                                   # 121 dofs: stands for prior.R
    Bp = hnp.SparseOperator(Mp)
    d_asg, V_asg = hnp.doublePassG(Aj, Bp, Bp, hnp.MultiVector.from_dense(Om_as), ka, s=1)
    # KLE from samples, rank 128 (+10) as in test_KLEProjector.py:57-59
    kk = 128
    m_data = syn.snapshots(n, 512, r0=200, decay=2.0, eps=1e-9, seed=5)
    Om_k = syn.gaussian_omega(n, kk + p, seed=4)
    Ak = hf.CollectiveOperator(P.SandwichedCovarianceOperator(m_data, M), hf.NullCollective(), mpi_op="avg")
    d_k, V_k = hnp.doublePassG(Ak, B, B, hnp.MultiVector.from_dense(Om_k), kk, s=1)
    np.savez_compressed(os.path.join(OUT, "doublepass_ref.npz"), Omega=Om, d_pod=d_pod, U_pod=U_pod.to_dense(),
                        d_w=d_w, U_w=U_w.to_dense(), Omega_as=Om_as, d_as=d_as, V_as=V_as.to_dense(),
                        d_asg=d_asg, V_asg=V_asg.to_dense(), m_data_seed=5, Omega_kle=Om_k, d_kle=d_k,
                        V_kle=V_k.to_dense())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
