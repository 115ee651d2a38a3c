"""GPU: the small operator classes around the stored-data path -- JJT, npToDolfinOperator, LowRankRectangularOperator
(hippyflow/modeling/jacobian.py:169-193, operatorWrappers.py:19-52, lowRankRectangularOperator.py:19-72) -- against NumPy.
The same bodies run under the CPU test double in test_host_logic_cpu.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture
def hf():
    import hippyflow_b200
    return hippyflow_b200


def test_jjt_and_dense_operator(hf, cuda_device):
    rng = np.random.default_rng(0)
    J = rng.standard_normal((10, 37))
    X = rng.standard_normal((10, 6))
    op = hf.JJT(J, device=cuda_device)
    Xd = hf.DeviceMultiVector.from_dense(X, cuda_device)
    Yd = hf.DeviceMultiVector(10, 6, device=cuda_device)
    op.matMvMult(Xd, Yd)
    np.testing.assert_allclose(Yd.to_dense(), J @ (J.T @ X), rtol=1e-12, atol=1e-12)
    x, y = hf.DeviceVector(10, cuda_device), hf.DeviceVector(1, cuda_device)
    op.init_vector(y, 0)
    x.set_local(X[:, 3])
    op.mult(x, y)
    np.testing.assert_allclose(y.get_local(), J @ (J.T @ X[:, 3]), rtol=1e-12, atol=1e-12)
    op.mult(Xd[1], y)                                            # a column view of a block (odd byte offset) as input
    np.testing.assert_allclose(y.get_local(), J @ (J.T @ X[:, 1]), rtol=1e-12, atol=1e-12)

    A = hf.npToDolfinOperator(J, device=cuda_device)
    u, v = hf.DeviceVector(1, cuda_device), hf.DeviceVector(1, cuda_device)
    A.init_vector(u, 0)
    A.init_vector(v, 1)
    assert u.size() == 10 and v.size() == 37
    z = rng.standard_normal(37)
    v.set_local(z)
    A.mult(v, u)
    np.testing.assert_allclose(u.get_local(), J @ z, rtol=1e-12, atol=1e-13)
    A.transpmult(x, v)
    np.testing.assert_allclose(v.get_local(), J.T @ X[:, 3], rtol=1e-12, atol=1e-13)
    Zd = hf.DeviceMultiVector(37, 6, device=cuda_device)
    A.matMvTranspmult(Xd, Zd)
    np.testing.assert_allclose(Zd.to_dense(), J.T @ X, rtol=1e-12, atol=1e-13)
    with pytest.raises(ValueError):
        A.init_vector(u, 2)


def test_low_rank_rectangular_operator(hf, cuda_device):
    rng = np.random.default_rng(1)
    dQ, dM, r = 12, 45, 5
    U, V, s = np.linalg.qr(rng.standard_normal((dQ, r)))[0], np.linalg.qr(rng.standard_normal((dM, r)))[0], rng.random(r) + 0.5
    A = U @ np.diag(s) @ V.T
    op = hf.LowRankRectangularOperator(hf.DeviceMultiVector.from_dense(U, cuda_device), s,
                                       hf.DeviceMultiVector.from_dense(V, cuda_device))
    x, y = hf.DeviceVector(1, cuda_device), hf.DeviceVector(1, cuda_device)
    op.init_vector(y, 0)
    op.init_vector(x, 1)
    assert (y.size(), x.size()) == (dQ, dM)
    z = rng.standard_normal(dM)
    x.set_local(z)
    y.set_local(np.full(dQ, 7.0))                                # mult zeroes y first (lowRankRectangularOperator.py:55)
    op.mult(x, y)
    np.testing.assert_allclose(y.get_local(), A @ z, rtol=1e-12, atol=1e-13)
    w = rng.standard_normal(dQ)
    y.set_local(w)
    op.transpmult(y, x)
    np.testing.assert_allclose(x.get_local(), A.T @ w, rtol=1e-12, atol=1e-13)
    X = rng.standard_normal((dM, 7))
    Yd = hf.DeviceMultiVector(dQ, 7, device=cuda_device)
    op.matMvMult(hf.DeviceMultiVector.from_dense(X, cuda_device), Yd)
    np.testing.assert_allclose(Yd.to_dense(), A @ X, rtol=1e-12, atol=1e-13)
    Zd = hf.DeviceMultiVector(dM, 7, device=cuda_device)
    op.matMvTranspmult(Yd, Zd)
    np.testing.assert_allclose(Zd.to_dense(), A.T @ (A @ X), rtol=1e-12, atol=1e-13)
    with pytest.raises(ValueError):
        op.init_vector(x, 3)
