"""The archived FFT / tanh score variant (holE-20170724/graph.pbtxt:6221-6521): the NumPy oracle against an
independent torch restatement of the graph's op list (torch.fft + autograd) and against finite differences."""
import numpy as np
import pytest
import torch

from oracle import hole_ccorr as oc
from oracle import hole_oracle as ho


def _table(n, dim, seed, scale):
    rng = np.random.default_rng(seed)
    E = rng.standard_normal((n, dim))
    E *= (scale * rng.uniform(0.5, 1.5, size=(n, 1))) / np.linalg.norm(E, axis=1, keepdims=True)
    return E


def _torch_loss(E, pos, neg, margin):
    """The op list of the archived graph, written with torch ops: clip_by_norm (graph.pbtxt:5307-5744),
    Complex, FFT, Conj, Mul, IFFT, Mul, Real + Imag, Sum, Tanh, Sub, Add, Maximum."""
    H = E.shape[1] // 2

    def emb(ids):
        x = E[ids]
        y = x * torch.minimum(torch.rsqrt((x * x).sum(1, keepdim=True)), torch.ones((), dtype=x.dtype))
        return torch.complex(y[:, :H], y[:, H:])

    def value(tr):
        h, t, r = emb(tr[:, 0]), emb(tr[:, 1]), emb(tr[:, 2])
        c = torch.fft.ifft(torch.conj(torch.fft.fft(h, dim=1)) * torch.fft.fft(t, dim=1), dim=1)
        m = r * c
        return torch.tanh((m.real + m.imag).sum(1))

    vp, vn = value(pos), value(neg)
    return torch.clamp(vp - vn + margin, min=0.0), vp, vn


@pytest.mark.parametrize("dim,side", [(8, 0), (8, 1), (30, 0), (30, 1)])
def test_ccorr_step_matches_torch_autograd(dim, side):
    n, B = 40, 24
    E = _table(n, dim, 7 + dim + side, 1.0)          # row norms 0.5..1.5: both clip branches
    rng = np.random.default_rng(3)
    pos = np.stack([rng.integers(5, n, B), rng.integers(5, n, B), rng.integers(0, 5, B)], axis=1).astype(np.int32)
    neg_ent = rng.integers(5, n, B).astype(np.int32)
    margin, lr = 1.0, 0.1
    Et = torch.tensor(E, dtype=torch.float64, requires_grad=True)
    neg = ho.corrupt_triples(pos, neg_ent, side)
    loss_t, vp_t, vn_t = _torch_loss(Et, torch.as_tensor(pos, dtype=torch.long), torch.as_tensor(neg, dtype=torch.long), margin)
    loss_t.sum().backward()
    want = E - lr * Et.grad.numpy()
    got = E.copy()
    loss, vp, vn = oc.sgd_step(got, pos, neg_ent, side, margin, lr, dtype=np.float64)
    np.testing.assert_allclose(vp, vp_t.detach().numpy(), atol=1e-13)
    np.testing.assert_allclose(vn, vn_t.detach().numpy(), atol=1e-13)
    np.testing.assert_allclose(loss, loss_t.detach().numpy(), atol=1e-13)
    np.testing.assert_allclose(got, want, atol=1e-12)
    assert np.abs(got - E).max() > 1e-3


def test_ccorr_direct_sum_equals_fft():
    rng = np.random.default_rng(0)
    for H in (4, 15, 75):
        h = rng.standard_normal((3, H)) + 1j * rng.standard_normal((3, H))
        t = rng.standard_normal((3, H)) + 1j * rng.standard_normal((3, H))
        np.testing.assert_allclose(oc.ccorr_direct(h, t), oc.ccorr_fft(h, t), atol=1e-11)


def test_ccorr_gradient_by_finite_differences():
    n, dim = 6, 10
    E = _table(n, dim, 11, 1.0)
    tri = np.array([[3, 4, 1], [5, 3, 0]], dtype=np.int32)
    g = np.array([1.0, -0.5])
    dh, dt_, dr = oc._side_grads(E, tri, g, np.float64)
    eps = 1e-6
    for col, d in ((0, dh), (1, dt_), (2, dr)):
        for i in range(len(tri)):
            row = tri[i, col]
            # rows are distinct within a triple here, so the derivative w.r.t. one row's entry is d[i]
            for k in (0, dim // 2, dim - 1):
                Ep, Em = E.copy(), E.copy()
                Ep[row, k] += eps
                Em[row, k] -= eps
                fd = (oc.raw_score(Ep, tri[i:i + 1], np.float64)[0] - oc.raw_score(Em, tri[i:i + 1], np.float64)[0]) / (2 * eps)
                assert abs(g[i] * fd - d[i, k]) < 1e-7


def test_ccorr_fp32_close_to_fp64():
    E = _table(50, 150, 5, 1.0)
    rng = np.random.default_rng(1)
    tri = np.stack([rng.integers(5, 50, 64), rng.integers(5, 50, 64), rng.integers(0, 5, 64)], axis=1).astype(np.int32)
    s64 = oc.raw_score(E, tri, np.float64)
    s32 = oc.raw_score(E.astype(np.float32), tri, np.float32)
    s32d = oc.raw_score(E.astype(np.float32), tri, np.float32, direct=True)
    assert np.abs(s32 - s64).max() < 5e-6 and np.abs(s32d - s64).max() < 5e-6
