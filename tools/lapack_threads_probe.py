"""Host LAPACK dsyevd (m = 266, the Rayleigh matrix of cfg2) against the BLAS thread count, raised through threadpoolctl as
hippyflow_b200.linalg.top_k_eig does under torchrun (OMP_NUM_THREADS=1).  B200 host, 16 cores: 4.30 / 3.77 / 3.47 / 3.35 ms at
1 / 2 / 4 / 6 threads; started with OMP_NUM_THREADS=1 and raised: 4.25 / 4.10 / 3.84 / 3.72 ms."""
import numpy as np, time, os
from scipy.linalg import lapack
from threadpoolctl import ThreadpoolController
A=np.random.randn(1000,266); T=A.T@A
ctl=ThreadpoolController()
print("cpus", os.cpu_count(), len(os.sched_getaffinity(0)), [(i['internal_api'], i['num_threads']) for i in ctl.info()])
for nt in [t for t in (1, 2, 3, 4, 6, 8, 12, 16) if t <= len(os.sched_getaffinity(0))]:      # never oversubscribe: ~1 s per call
    t0=time.perf_counter()
    with ctl.limit(limits=nt, user_api="blas"):
        t1=time.perf_counter()
        lapack.dsyevd(T,compute_v=1,lower=1)
        ts=[]
        for _ in range(20):
            t=time.perf_counter(); lapack.dsyevd(T,compute_v=1,lower=1); ts.append(time.perf_counter()-t)
    print(nt, "threads: dsyevd median %.2f ms min %.2f ms; ctx enter %.3f ms" % (np.median(ts)*1e3, min(ts)*1e3, (t1-t0)*1e3))
