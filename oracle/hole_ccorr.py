"""NumPy restatement of the ARCHIVED score variant of the reference (run directory holE-20170724).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py for who may import this.

What the archived graph computes (holE-20170724/graph.pbtxt:6221-6521, ops FFT, Conj, FFT_1, Mul, IFFT,
Mul_1, Real, Imag, add, Sum, Tanh on the complex vectors h, r, t of length H = D/2 built exactly like the
live model's: embedding_lookup(max_norm=1), split halves, tf.complex -- graph.pbtxt:5307-6187):

    c = ifft(conj(fft(h)) * fft(t))          # complex circular correlation, c_k = sum_j conj(h_j) t_{j+k}
    m = r * c
    s = sum_k Re(m_k) + Im(m_k)
    value = tanh(s)

and the loss max(tanh(s+) - tanh(s-) + margin, 0) with margin 1.0 in that run (graph.pbtxt:15874-15988;
the live script's --margin default is 0.2).  Corruption, the clip, the ones gradient seed, the slice order
and the ScatterSub update are the live model's (oracle/hole_oracle.py).

PARITY UNPINNED BY REFERENCE EXECUTION.  The live holE.py no longer contains this variant (its line 176
is a dead rfft helper), so there is no reference source to run, and TensorFlow 1.2 itself is not
installable here.  The op list above is transcribed from the GraphDef; tests/test_ccorr_oracle.py checks
this module against an independent torch.fft + autograd restatement of the same op list and against
finite differences.
"""
import numpy as np

from .hole_oracle import clip_rows, corrupt_triples


def _cplx(y, H):
    return y[:, :H] + 1j * y[:, H:]


def ccorr_fft(h, t):
    """c_k = sum_j conj(h_j) t_{(j+k) mod H} through the FFT, as the graph does (6221-6314)."""
    return np.fft.ifft(np.conj(np.fft.fft(h, axis=1)) * np.fft.fft(t, axis=1), axis=1)


def ccorr_direct(h, t):
    """The same sum, term by term in index order j = 0..H-1 (what the device kernel does)."""
    H = h.shape[1]
    idx = (np.arange(H)[:, None] + np.arange(H)[None, :]) % H          # [j, k] -> (j + k) mod H
    out = np.zeros_like(t)
    for j in range(H):
        out = out + np.conj(h[:, j:j + 1]) * t[:, idx[j]]
    return out


def raw_score(E, triples, dtype=np.float32, direct=False):
    """s = sum_k Re(r_k c_k) + Im(r_k c_k) (graph.pbtxt:6334-6482).  [B] array."""
    triples = np.asarray(triples)
    H = E.shape[1] // 2
    cd = np.complex64 if np.dtype(dtype) == np.float32 else np.complex128
    ys = [clip_rows(E[triples[:, col].astype(np.int64)].astype(dtype, copy=False))[0] for col in range(3)]
    h, t, r = (_cplx(y, H).astype(cd) for y in ys)
    c = (ccorr_direct if direct else ccorr_fft)(h, t).astype(cd)
    m = r * c
    return np.sum(m.real + m.imag, axis=1).astype(dtype)


def evaluate_triples(E, triples, dtype=np.float32, direct=False):
    """tanh(s) (graph.pbtxt:6521)."""
    return np.tanh(raw_score(E, triples, dtype, direct))


def _side_grads(E, triples, g, dtype):
    """Gradient of sum_i g_i * s_i w.r.t. the unclipped gathered rows (h, t, r) of one side, [B, D] each.

    With rho = (1 - i) r:  s = Re sum_k rho_k c_k, hence
      d s / d h_j = u_j = sum_k rho_k t_{j+k}            -> [Re u ;  Im u]
      d s / d t_m = conj(w_m), w_m = sum_k rho_k conj(h_{m-k})   -> [Re w ; -Im w]
      d s / d r_k = [Re c_k + Im c_k ; Re c_k - Im c_k]
    then through the clip exactly as in hole_oracle._side_grads (App. A.3)."""
    triples = np.asarray(triples)
    H = E.shape[1] // 2
    ys, invs, clipped = [], [], []
    for col in range(3):
        x = E[triples[:, col].astype(np.int64)].astype(dtype, copy=False)
        y, inv, cl = clip_rows(x)
        ys.append(y); invs.append(inv); clipped.append(cl)
    h, t, r = (_cplx(y, H) for y in ys)
    rho = (1 - 1j) * r
    c = ccorr_fft(h, t)
    u = ccorr_fft(np.conj(rho), t)                                   # sum_k rho_k t_{j+k}
    w = np.fft.ifft(np.fft.fft(rho, axis=1) * np.fft.fft(np.conj(h), axis=1), axis=1)   # circular convolution
    g = g.astype(dtype)[:, None]
    dy_h = g * np.concatenate([u.real, u.imag], axis=1).astype(dtype)
    dy_t = g * np.concatenate([w.real, -w.imag], axis=1).astype(dtype)
    dy_r = g * np.concatenate([c.real + c.imag, c.real - c.imag], axis=1).astype(dtype)
    out = []
    for dy, y, inv, cl in zip((dy_h, dy_t, dy_r), ys, invs, clipped):
        proj = np.sum(y * dy, axis=1, keepdims=True)
        with np.errstate(invalid="ignore", over="ignore"):
            dx_clip = (dy - y * proj) * inv
        out.append(np.where(cl, dx_clip, dy))
    return out


def indexed_slices(E, pos, neg_ent, side, margin=1.0, dtype=np.float32):
    """The six IndexedSlices in the live graph's order [r+, r-, t+, t-, h+, h-] plus the loss rows:
    loss = max(tanh(s+) - tanh(s-) + margin, 0), active on >= 0, d tanh = 1 - tanh^2."""
    pos = np.asarray(pos)
    neg = corrupt_triples(pos, neg_ent, side)
    vp = evaluate_triples(E, pos, dtype)
    vn = evaluate_triples(E, neg, dtype)
    dt = vp.dtype.type
    pre = vp - vn + dt(margin)
    loss = np.maximum(pre, dt(0.0))
    act = (pre >= dt(0.0)).astype(dtype)
    gp = act * (dt(1.0) - vp * vp)
    gn = -act * (dt(1.0) - vn * vn)
    dh_p, dt_p, dr_p = _side_grads(E, pos, gp, dtype)
    dh_n, dt_n, dr_n = _side_grads(E, neg, gn, dtype)
    slices = [
        (pos[:, 2], dr_p), (neg[:, 2], dr_n),
        (pos[:, 1], dt_p), (neg[:, 1], dt_n),
        (pos[:, 0], dh_p), (neg[:, 0], dh_n),
    ]
    return slices, loss, vp, vn


def sgd_step(E, pos, neg_ent, side, margin, lr, dtype=np.float32):
    """One training step in place on E: E[idx] -= lr * g for each of the 6B pairs (sequential ScatterSub
    in concat order, duplicates all applying).  Returns (loss[B], tanh+[B], tanh-[B])."""
    slices, loss, vp, vn = indexed_slices(E, pos, neg_ent, side, margin, dtype)
    dt = np.dtype(dtype).type
    Ew = E if E.dtype == np.dtype(dtype) else E.astype(dtype)
    idx = np.concatenate([s[0] for s in slices]).astype(np.int64)
    upd = np.concatenate([s[1] for s in slices], axis=0) * dt(lr)
    np.subtract.at(Ew, idx, upd)
    if Ew is not E:
        E[...] = Ew
    return loss, vp, vn


def all_scores(E, queries, side, cand_ids, dtype=np.float64):
    """Raw scores s of every candidate for every query, [Q, C]: side "tail" scores (h, c, r), "head" (c, t, r).
    Straight from the definition (raw_score on the expanded triples), not through the query-vector form the
    device kernel uses."""
    queries = np.asarray(queries)
    cand_ids = np.asarray(cand_ids)
    out = np.empty((len(queries), len(cand_ids)), dtype=dtype)
    for i, (h, t, r) in enumerate(queries.tolist()):
        tri = np.empty((len(cand_ids), 3), dtype=np.int64)
        tri[:, 0] = cand_ids if side == "head" else h
        tri[:, 1] = cand_ids if side == "tail" else t
        tri[:, 2] = r
        out[i] = raw_score(E, tri, dtype)
    return out
