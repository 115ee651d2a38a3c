import numpy as np, scipy.linalg as sla, time
from scipy.linalg import lapack
from threadpoolctl import threadpool_limits, threadpool_info
A=np.random.randn(1000,266); T=A.T@A; G=np.eye(266)+1e-7*np.random.randn(266,266); G=(G+G.T)/2
def manual():
    R,_=lapack.dpotrf(G,lower=0,clean=1); S,_=lapack.dtrtri(R,lower=0)
    Tp=S.T@T@S; Tp=(Tp+Tp.T)/2
    d,V,_=lapack.dsyevd(Tp,compute_v=1,lower=1)
    return d, S@V
fns={'dpotrf+dtrtri':lambda:(lapack.dtrtri(lapack.dpotrf(G,lower=0,clean=1)[0],lower=0)),'dsyevd':lambda:lapack.dsyevd(T,compute_v=1,lower=1),'manual_gen':manual,'dsygvd':lambda:lapack.dsygvd(T,G,itype=1,jobz='V',uplo='L')}
print([ (i['internal_api'], i['num_threads']) for i in threadpool_info()])
for nt in (None,1,2,4):
    for name,fn in fns.items():
        ctx = threadpool_limits(nt) if nt else threadpool_limits(None)
        with ctx:
            fn(); t=time.perf_counter()
            for _ in range(10): fn()
            print(nt,name,round((time.perf_counter()-t)/10*1e3,2),'ms')
