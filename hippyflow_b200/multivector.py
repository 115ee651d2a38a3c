"""Device-resident stand-ins for dolfin.Vector / hippylib.MultiVector at the boundary of the hot path.

A ``DeviceMultiVector`` of k vectors of length n is ONE (n, k) float64 row-major block in HBM with a
padded leading dimension -- the dense layout ``mv_to_dense`` produces (hippyflow/utilities/
mv_utilities.py:31-41; column j = vector j) -- so a whole sketch is a single GEMM operand and a single
NCCL message.  Only the methods the reference's projectors, operators and collectives touch are
provided (nvec, [], zero, swap, dot_mv, dot_v, reduce, get_local/set_local, axpy, inner).
"""
import numpy as np
import torch

from . import _lib as K


class DeviceVector:
    """One vector of length n: a strided column view of a multivector block, or its own (n, 1) block."""

    def __init__(self, n=None, device=None, _view=None):
        if _view is not None:
            self._t = _view  # (n, 1) view
        else:
            self._t = K.padded_zeros(int(n), 1, device, pad=2)

    def init(self, n):
        self._t = K.padded_zeros(int(n), 1, self._t.device, pad=2)

    def storage_tensor(self):
        return self._t

    def size(self):
        return self._t.shape[0]

    def get_local(self):
        return self._t[:, 0].cpu().numpy().copy()

    def set_local(self, a):
        self._t[:, 0].copy_(torch.as_tensor(np.asarray(a, dtype=np.float64)))

    def apply(self, mode=""):
        pass

    def zero(self):
        self._t.zero_()

    def axpy(self, alpha, x):
        K.axpby_(alpha, x._t, 1.0, self._t)

    def inner(self, x):
        return float(K.coldot(self._t, x._t)[0])

    def norm(self, kind="l2"):
        return float(np.sqrt(self.inner(self)))

    def tensor(self):
        return self._t[:, 0]


class DeviceMultiVector:
    def __init__(self, arg, nvec=None, device=None):
        if isinstance(arg, DeviceMultiVector):  # copy constructor, hp.MultiVector(mv)
            self._t = K.padded_empty(arg._t.shape[0], arg._t.shape[1], arg._t.device)
            self._t.copy_(arg._t)
        elif isinstance(arg, DeviceVector):  # hp.MultiVector(vector, nvec): zero block
            self._t = K.padded_zeros(arg.size(), int(nvec), arg._t.device)
        elif isinstance(arg, torch.Tensor):  # adopt a (n, k) device block (no copy)
            self._t = K._req(arg, "block")
        elif isinstance(arg, (int, np.integer)):
            self._t = K.padded_zeros(int(arg), int(nvec), device)
        else:
            raise TypeError("DeviceMultiVector(mv | vector, nvec | tensor | n, nvec, device)")

    # -- boundary conversions (mv_utilities.py:18-49)
    @staticmethod
    def from_dense(a, device):
        return DeviceMultiVector(K.to_padded(np.asarray(a, dtype=np.float64), device))

    def to_dense(self):
        return self._t.cpu().numpy().copy()

    def tensor(self):
        return self._t

    def adopt(self, block):
        """Re-point this multivector at another (n, k) device block of the same shape (no copy): how an operator hands back a
        result that was produced in memory of its own (the NVLink exchange buffer of the sketch allreduce)."""
        assert tuple(block.shape) == tuple(self._t.shape)
        self._t = K._req(block, "block")

    def storage_tensor(self):
        """The full padded block (one contiguous NCCL message)."""
        t = self._t
        ld = K._ld(t)
        return t.as_strided((t.shape[0], ld), (ld, 1)) if t.shape[0] > 0 else t

    # -- hippylib.MultiVector protocol
    def nvec(self):
        return self._t.shape[1]

    def __getitem__(self, j):
        return DeviceVector(_view=self._t[:, j:j + 1])

    def zero(self):
        self._t.zero_()

    def swap(self, other):
        self._t, other._t = other._t, self._t

    def dot_mv(self, mv):
        """(self.nvec x mv.nvec) matrix of inner products -- a split-K TN GEMM."""
        return K.dgemm(K.HFB_TN, self._t, mv._t).cpu().numpy()

    def dot_v(self, v):
        vt = v._t
        if vt.data_ptr() % 16 or K._ld(vt) % 2:          # a column view mv[j] with odd j: stage it for TMA
            vt = K.to_padded(vt, vt.device, pad=2)
        return K.dgemm(K.HFB_TN, self._t, vt).cpu().numpy()[:, 0]

    def reduce(self, y, alpha):
        """y += sum_i alpha_i self[i]."""
        a = K.to_padded(np.asarray(alpha, dtype=np.float64).reshape(-1, 1), self._t.device, pad=2)
        tmp = K.dgemm(K.HFB_NN, self._t, a)
        K.axpby_(1.0, tmp, 1.0, y._t)

    def axpy(self, alpha, mv):
        K.axpby_(float(alpha), mv._t, 1.0, self._t)

    def scale(self, alpha):
        K.axpby_(0.0, self._t, float(alpha), self._t)


def mv_to_dense(multivector):
    """hippyflow/utilities/mv_utilities.py:31-41 for device multivectors (and anything with nvec/[])."""
    if isinstance(multivector, DeviceMultiVector):
        return multivector.to_dense()
    n = multivector[0].get_local().shape[0]
    out = np.zeros((n, multivector.nvec()))
    for i in range(multivector.nvec()):
        out[:, i] = multivector[i].get_local()
    return out


mv_to_dense_local = mv_to_dense


def dense_to_mv_local(dense_array, device_or_vector):
    """mv_utilities.py:43-53: dense (n, k) -> multivector."""
    device = device_or_vector._t.device if hasattr(device_or_vector, "_t") else device_or_vector
    return DeviceMultiVector.from_dense(dense_array, device)
