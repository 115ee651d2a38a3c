"""Time the UNMODIFIED reference's PODProjectorFromData.construct_subspace (hep / ghep) in the build container, where
/root/reference exists (it does not travel to the GPU box, so bench.py's cpu_baseline leg times the oracle port there).
TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Usage: python tools/time_reference_verbatim.py [n_side] [N] [rank]"""
import contextlib
import io
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_import import import_reference  # noqa: E402
from oracle import projectors_np as P  # noqa: E402
from hippyflow_b200 import synthetic as syn  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 257
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    r = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    hf = import_reference()
    M = syn.p1_mass_matrix(side - 1)
    n = M.shape[0]
    u = syn.snapshots(n, N, r0=min(512, N), seed=0)
    pod = object.__new__(hf.PODProjectorFromData)          # __init__ needs FEniCS only to assemble M
    pod.M_csr = M
    out = {"n": n, "N": N, "rank": r, "cores": os.cpu_count()}
    ref_d = None
    for method in ("hep", "ghep"):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            d, phi, Mphi, shift = pod.construct_subspace(u.copy(), r, shifted=True, method=method)
        out[method + "_seconds"] = time.perf_counter() - t0
        if ref_d is None:
            ref_d = d
        out[method + "_max_rel_eig_diff_vs_hep"] = float(np.max(np.abs(d - ref_d) / ref_d))
    # the randomized weighted double pass of the oracle port on the same data (what the GPU path computes)
    Om = syn.gaussian_omega(n, r + 10, seed=1)
    t0 = time.perf_counter()
    d_r = P.pod_randomized_weighted(u, M, r, Om, shifted=True)[0]
    out["oracle_randomized_seconds"] = time.perf_counter() - t0
    # a single range-finding pass resolves the top of the spectrum; modes close to the sketch size k + p carry the
    # truncation error of the randomized method itself (same for hIPPYlib's doublePassG), so report both ends
    rel = np.abs(d_r - ref_d) / ref_d
    out["randomized_vs_hep_rel_eig_diff_first10_max"] = float(rel[:10].max())
    out["randomized_vs_hep_rel_eig_diff_last10_max"] = float(rel[-10:].max())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
