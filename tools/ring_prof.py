import sys; sys.path.insert(0, ".")
import torch
from hippyflow_b200 import _lib as K, synthetic as syn
from hippyflow_b200.linalg import CsrMatrix
dev = torch.device("cuda:0")
for n, m in ((263169, 266), (263169, 74)):
    M = syn.p1_mass_matrix_for(n)
    Md = CsrMatrix(M, dev); Md.impl = "ring"
    B = K.padded_empty(n, m, dev).normal_(); C = K.padded_empty(n, m, dev)
    for _ in range(3): Md.matmat(B, out=C)
    torch.cuda.synchronize()
