"""NumPy restatement of holE.py's training step and link-prediction ranking.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py for who may import this.

PARITY PINNED MODULO A TENSORFLOW SHIM.  The reference computes this arithmetic inside TensorFlow 1.2
(GraphDef producer 22), which is neither vendored under /root/reference nor installable offline, and it
ships no tests or golden vectors for this path.  The pin is
tests/golden/tfshim_step.npz: the outputs of /root/reference/holE.py's OWN corrupt_batch / get_embedding /
evaluate_triples / evaluate_batch / GradientDescentOptimizer.minimize source, executed unmodified on the
torch-backed `tensorflow` stand-in tests/golden/tfshim.py (which states the TF-internal conventions of
SURVEY.md App. B once: clip formula and axes, Minimum/Maximum tie routing, ones gradient seed, slice
order, sequential ScatterSub, fp32 inverse_time_decay).  tests/test_tfshim_golden.py holds sgd_step and
logloss_step to 1e-12 (fp64) and 5e-7 (fp32) against it.  Every function below cites the holE.py lines
it restates.  Also pinned:
  * the ranking / metric routines, against the reference's own pure-Python
    eval_link_prediction / score_mrr (holE.py:427-490) executed here with TensorFlow
    stubbed out -- see tests/golden/make_golden.py and tests/golden/ranking_ref.json;
  * the Xavier constant 0.00151367869694 (holE-20170724/graph.pbtxt:1791-1830);
  * the gradient, by finite differences and by an independent torch-autograd
    restatement of holE.py:161-168,191-192,198,231 (tests/test_oracle.py);
  * the Philox sampler, against Random123 known-answer vectors.

All entry points take ``dtype`` (np.float32 mirrors the fp32 TF graph op by op;
np.float64 is the high-precision yardstick).
Triple column order is (head, tail, relation) -- holE.py:80-81.
"""
import heapq

import numpy as np

from . import philox

# --------------------------------------------------------------------------------------
# init (holE.py:263-264; App. A.7)
# --------------------------------------------------------------------------------------


def xavier_stddev(n_rows, dim):
    """tf.contrib.layers.xavier_initializer(uniform=False) for a [N, D] variable:
    truncated normal with stddev sqrt(1.3 * 2 / (fan_in + fan_out)) (holE.py:263-264;
    pinned by the const 0.00151367869694 at holE-20170724/graph.pbtxt:1791-1830)."""
    return float(np.sqrt(2.6 / (n_rows + dim)))


def xavier_init(n_rows, dim, rng, dtype=np.float32):
    """Truncated normal (resample outside +-2 sigma), mean 0 (TF TruncatedNormal op)."""
    sd = xavier_stddev(n_rows, dim)
    x = rng.standard_normal((n_rows, dim))
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * sd).astype(dtype)


# --------------------------------------------------------------------------------------
# forward (holE.py:161-168, 179-202, 222-234)
# --------------------------------------------------------------------------------------


def clip_rows(x):
    """tf.nn.embedding_lookup(..., max_norm=1) == clip_by_norm(row, 1) (holE.py:162).

    TF 1.2: y = x * 1.0 * min(rsqrt(sum x^2), 1/1.0)  (App. B, graph.pbtxt:3107-3595).
    Returns (y, inv_norm, clipped_mask) where clipped_mask = (rsqrt <= 1), the branch the
    Minimum gradient takes (graph.pbtxt:40692).
    """
    dt = x.dtype.type
    n2 = np.sum(x * x, axis=-1, keepdims=True)
    with np.errstate(divide="ignore"):
        inv = dt(1.0) / np.sqrt(n2)
    scale = np.minimum(inv, dt(1.0))
    return x * scale, inv, (inv <= dt(1.0))


def get_embedding(E, ids, dtype=np.float32):
    """holE.py:161-168: gather, clip, split halves -> (real, imag) each [B, D/2]."""
    x = E[np.asarray(ids, dtype=np.int64)].astype(dtype, copy=False)
    y, _, _ = clip_rows(x)
    H = E.shape[1] // 2
    return y[:, :H], y[:, H:]


def score(E, triples, dtype=np.float32):
    """holE.py:191-192: s = sum_k Re(h_k * (r_k * conj(t_k))).  [B] array."""
    triples = np.asarray(triples)
    a, b = get_embedding(E, triples[:, 0], dtype)
    e, f = get_embedding(E, triples[:, 1], dtype)
    c, d = get_embedding(E, triples[:, 2], dtype)
    # r * conj(t) = (c + i d)(e - i f) = (c e + d f) + i (d e - c f)
    pr = c * e + d * f
    pi = d * e - c * f
    # Re(h * p) = a pr - b pi
    return np.sum(a * pr - b * pi, axis=1)


def sigmoid(s):
    dt = s.dtype.type
    return dt(1.0) / (dt(1.0) + np.exp(-s))


def evaluate_triples(E, triples, dtype=np.float32):
    """holE.py:179-202 (non-log-loss branch): sigma(score).  Returns [B]."""
    return sigmoid(score(E, triples, dtype))


def corrupt_triples(pos, neg_ent, side):
    """Rebuild the corrupt [B,3] batch (holE.py:113, 132): side 1 replaces heads."""
    neg = np.array(pos, copy=True)
    neg[:, 0 if side else 1] = neg_ent
    return neg


def evaluate_batch(E, pos, neg_ent, side, margin=0.2, dtype=np.float32):
    """holE.py:222-234: hinge loss max(sigma+ - sigma- + margin, 0) per triple."""
    vp = evaluate_triples(E, pos, dtype)
    vn = evaluate_triples(E, corrupt_triples(pos, neg_ent, side), dtype)
    dt = vp.dtype.type
    return np.maximum(vp - vn + dt(margin), dt(0.0)), vp, vn


# --------------------------------------------------------------------------------------
# backward + SGD (holE.py:291-296; App. A.3, App. B)
# --------------------------------------------------------------------------------------


def inverse_time_decay(lr0, step, decay_steps, decay_rate, dtype=np.float32):
    """tf.train.inverse_time_decay (holE.py:292-294): lr0 / (1 + rate * step / decay_steps),
    evaluated in fp32 with global_step cast to float (graph.pbtxt:16151-16412)."""
    dt = np.dtype(dtype).type
    p = dt(step) / dt(decay_steps)
    return dt(lr0) / (dt(1.0) + dt(decay_rate) * p)


def _side_grads(E, triples, g, dtype):
    """Gradient of sum_i g_i * s_i w.r.t. the *unclipped* gathered rows of one side.

    Returns (dx_h, dx_t, dx_r), each [B, D] (App. A.3)."""
    triples = np.asarray(triples)
    H = E.shape[1] // 2
    out = []
    xs, ys, invs, clipped = [], [], [], []
    for col in range(3):
        x = E[triples[:, col].astype(np.int64)].astype(dtype, copy=False)
        y, inv, cl = clip_rows(x)
        xs.append(x); ys.append(y); invs.append(inv); clipped.append(cl)
    a, b = ys[0][:, :H], ys[0][:, H:]
    e, f = ys[1][:, :H], ys[1][:, H:]
    c, d = ys[2][:, :H], ys[2][:, H:]
    g = g.astype(dtype)[:, None]
    dy_h = g * np.concatenate([c * e + d * f, c * f - d * e], axis=1)
    dy_t = g * np.concatenate([a * c - b * d, a * d + b * c], axis=1)
    dy_r = g * np.concatenate([a * e + b * f, a * f - b * e], axis=1)
    for dy, y, inv, cl in zip((dy_h, dy_t, dy_r), ys, invs, clipped):
        proj = np.sum(y * dy, axis=1, keepdims=True)
        with np.errstate(invalid="ignore", over="ignore"):
            dx_clip = (dy - y * proj) * inv
        out.append(np.where(cl, dx_clip, dy))
    return out


def indexed_slices(E, pos, neg_ent, side, margin=0.2, dtype=np.float32):
    """The six IndexedSlices TF concatenates before ScatterSub, in graph order
    [r+, r-, t+, t-, h+, h-] (App. B, graph.pbtxt:47849-48049), plus the loss.

    Gradient seed is ones (sum of losses, graph.pbtxt:16484-16550); hinge active on
    ``>= 0`` (graph.pbtxt:16739)."""
    pos = np.asarray(pos)
    neg = corrupt_triples(pos, neg_ent, side)
    vp = evaluate_triples(E, pos, dtype)
    vn = evaluate_triples(E, neg, dtype)
    dt = vp.dtype.type
    pre = vp - vn + dt(margin)
    loss = np.maximum(pre, dt(0.0))
    act = (pre >= dt(0.0)).astype(dtype)
    gp = act * vp * (dt(1.0) - vp)
    gn = -act * vn * (dt(1.0) - vn)
    dh_p, dt_p, dr_p = _side_grads(E, pos, gp, dtype)
    dh_n, dt_n, dr_n = _side_grads(E, neg, gn, dtype)
    slices = [
        (pos[:, 2], dr_p), (neg[:, 2], dr_n),
        (pos[:, 1], dt_p), (neg[:, 1], dt_n),
        (pos[:, 0], dh_p), (neg[:, 0], dh_n),
    ]
    return slices, loss, vp, vn


def sgd_step(E, pos, neg_ent, side, margin, lr, dtype=np.float32, order="tf"):
    """One training step in place on E (holE.py:296): E[idx] -= lr * g for each of the 6B
    (index, gradient-row) pairs, duplicates all applying.

    order="tf":     sequential application in concat order (TF CPU ScatterSub loop).
    order="merged": the device kernel's fixed order -- gradients of one row are first
                    summed over (slot, batch index) with slots [r, t, h, neg-entity] and
                    the +/- contributions of a shared row pre-added, then applied once.
                    Same real-number result; differs from "tf" only in rounding.
    Returns (loss[B], sigma_pos[B], sigma_neg[B]).
    """
    slices, loss, vp, vn = indexed_slices(E, pos, neg_ent, side, margin, dtype)
    dt = np.dtype(dtype).type
    Ew = E if E.dtype == np.dtype(dtype) else E.astype(dtype)
    if order == "tf":
        idx = np.concatenate([s[0] for s in slices]).astype(np.int64)
        upd = np.concatenate([s[1] for s in slices], axis=0) * dt(lr)
        np.subtract.at(Ew, idx, upd)
    elif order == "merged":
        (ri, rp), (_, rn), (ti, tp), (tni, tn), (hi, hp), (hni, hn) = slices
        if side:  # heads corrupted: t and r shared
            merged = [(ri, rp + rn), (ti, tp + tn), (hi, hp), (hni, hn)]
        else:     # tails corrupted: h and r shared
            merged = [(ri, rp + rn), (ti, tp), (hi, hp + hn), (tni, tn)]
        idx = np.concatenate([m[0] for m in merged]).astype(np.int64)
        g = np.concatenate([m[1] for m in merged], axis=0)
        order_ix = np.argsort(idx, kind="stable")
        acc = np.zeros_like(Ew)
        # sequential sum in (slot, batch) order per row
        np.add.at(acc, idx[order_ix], g[order_ix])
        rows = np.unique(idx)
        Ew[rows] -= dt(lr) * acc[rows]
    else:
        raise ValueError(order)
    if Ew is not E:
        E[...] = Ew
    return loss, vp, vn


def logloss_step(E, pos, neg_ents, sides, lr, l2=0.0, dtype=np.float32):
    """One --log_loss training step, in place on E (holE.py:194-196, 206-220, 296).

    Loss rows: the B positives with label +1 (holE.py:210), then one corrupt batch per
    entry of ``sides`` / ``neg_ents`` with label -1 (holE.py:213-217), concatenated
    (holE.py:221).  Row loss = log(1 + exp(-label * score)) + l2 * l2_loss(embeddings)
    (holE.py:194-196; tf.nn.l2_loss = sum(x^2)/2 over the WHOLE variable, the same scalar in
    every row).  minimize() differentiates the sum of the rows: the sparse score gradients
    (through clip_by_norm, as in the hinge branch) plus  n_rows_of_loss * l2 * E  for the
    whole table.  Returns (loss [(1+k), B] without the L2 scalar, l2_loss scalar)."""
    pos = np.asarray(pos)
    dt = np.dtype(dtype).type
    B, k = len(pos), len(sides)
    Ew = E.astype(dtype, copy=True)
    s_p = score(Ew, pos, dtype)
    losses = [np.log(dt(1.0) + np.exp(-s_p))]
    acc = np.zeros_like(Ew)

    def scatter(triples, g):
        for col, dx in zip((0, 1, 2), _side_grads(Ew, triples, g, dtype)):
            np.add.at(acc, triples[:, col].astype(np.int64), dx)

    scatter(pos, sigmoid(s_p) - dt(1.0))              # d/ds log(1+exp(-s)) = -sigmoid(-s)
    for j in range(k):
        neg = corrupt_triples(pos, neg_ents[j], sides[j])
        s_n = score(Ew, neg, dtype)
        losses.append(np.log(dt(1.0) + np.exp(s_n)))
        scatter(neg, sigmoid(s_n))                    # d/ds log(1+exp(+s)) = sigmoid(s)
    l2_loss = dt(0.5) * np.sum(Ew.astype(np.float64) ** 2)
    decay = dt(lr) * dt((1 + k) * B) * dt(l2)
    E[...] = Ew - dt(lr) * acc - decay * Ew
    return np.stack(losses), dtype(l2_loss)


# --------------------------------------------------------------------------------------
# corruption (holE.py:97-140, 343-347; App. A.6)
# --------------------------------------------------------------------------------------


def build_type_csr(type_of, n_types=None):
    """type -> entity-id CSR from a dense type_of[N] array (ids ascending inside a type,
    which is the order holE.py:61 appends them in)."""
    type_of = np.asarray(type_of, dtype=np.int64)
    T = int(type_of.max()) + 1 if n_types is None else n_types
    order = np.argsort(type_of, kind="stable")
    counts = np.bincount(type_of, minlength=T)
    off = np.zeros(T + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    return off, order.astype(np.int32)


def corrupt(pos, type_of, csr_off, csr_ids, seed, step):
    """Type-safe corruption (holE.py:97-140): one coin per batch picks the side; each
    triple's replacement is uniform over entities of the replaced entity's type.

    Returns (side, neg_ent[B] int32).  The draw is the Philox stream of oracle/philox.py.
    """
    pos = np.asarray(pos)
    side = philox.side_coin(seed, step)
    ent = pos[:, 0 if side else 1].astype(np.int64)
    ty = np.asarray(type_of)[ent].astype(np.int64)
    lo = np.asarray(csr_off)[ty]
    cnt = np.asarray(csr_off)[ty + 1] - lo
    j = philox.entity_draw(seed, step, np.arange(len(pos)), cnt)
    return side, np.asarray(csr_ids)[lo + j].astype(np.int32)


# --------------------------------------------------------------------------------------
# ranking + metrics (holE.py:427-490; App. A.4, A.5)
# --------------------------------------------------------------------------------------


def query_vectors(E, queries, side, dtype=np.float32):
    """App. A.4: q such that score(candidate j) = clip(E_j) . q.

    side "tail": q = [a c - b d ; a d + b c]   (h * r)
    side "head": q = [c e + d f ; c f - d e]   (r * conj(t), imaginary half negated)
    """
    queries = np.asarray(queries)
    a, b = get_embedding(E, queries[:, 0], dtype)
    e, f = get_embedding(E, queries[:, 1], dtype)
    c, d = get_embedding(E, queries[:, 2], dtype)
    if side == "tail":
        return np.concatenate([a * c - b * d, a * d + b * c], axis=1)
    if side == "head":
        return np.concatenate([c * e + d * f, c * f - d * e], axis=1)
    raise ValueError(side)


def all_scores(E, queries, side, cand_ids, dtype=np.float32):
    """Raw scores s[q, j] of every candidate for every query (GEMM form, App. A.4)."""
    q = query_vectors(E, queries, side, dtype)
    y, _, _ = clip_rows(E[np.asarray(cand_ids, dtype=np.int64)].astype(dtype, copy=False))
    return q @ y.T


def rank_counts(scores, cand_ids, true_ids, filter_lists=None):
    """Counts of candidates ranked strictly before the true one, ascending by
    (value, candidate id) -- the heap's tuple order (holE.py:434, 446-453).

    scores[q, j] is candidate cand_ids[j] for query q (lower is better, holE.py:231).
    Returns (raw_before[Q], filtered_before[Q]); rank = 1 + count.
    filter_lists[q] = iterable of train/valid-true candidate ids (holE.py:454-461); they
    are skipped without advancing the filtered rank.
    """
    cand_ids = np.asarray(cand_ids, dtype=np.int64)
    Q = scores.shape[0]
    raw = np.zeros(Q, dtype=np.int64)
    filt = np.zeros(Q, dtype=np.int64)
    pos_of = {int(c): j for j, c in enumerate(cand_ids)}
    for q in range(Q):
        jt = pos_of[int(true_ids[q])]
        thr = scores[q, jt]
        before = (scores[q] < thr) | ((scores[q] == thr) & (cand_ids < cand_ids[jt]))
        raw[q] = int(before.sum())
        nf = 0
        if filter_lists is not None:
            for fid in set(int(x) for x in filter_lists[q]):
                j = pos_of.get(fid)
                if j is not None and before[j]:
                    nf += 1
        filt[q] = raw[q] - nf
    return raw, filt


def eval_link_prediction_heap(values, triples, true_triples, test_triples, threshold=None):
    """Heap restatement of holE.py:427-469 for ONE (head, candidate-set) group.

    values[k] is sigma(s) of triples[k] = (h, t, r).  Returns lists (raw_positions,
    filtered_positions) appended in pop order.  ``threshold=None`` disables the
    ``min sigma < infer_threshold`` gate (holE.py:438), which is unreachable for the live
    model (surprise 2)."""
    heap = []
    min_loss = 100
    for v, tr in zip(values, triples):
        v = float(v)
        min_loss = min(min_loss, v)
        heapq.heappush(heap, (v, tuple(int(x) for x in tr)))
    confident = True if threshold is None else (min_loss < threshold)
    raw_positions, filtered_positions = [], []
    raw_rank = filtered_rank = 0
    while heap:
        v, (h, t, r) = heapq.heappop(heap)
        raw_rank += 1
        in_sample = t in true_triples.get(h, {}).get(r, ())
        if confident and in_sample:
            continue
        filtered_rank += 1
        if confident and t in test_triples.get(h, {}).get(r, ()):
            raw_positions.append(raw_rank)
            filtered_positions.append(filtered_rank)
    return raw_positions, filtered_positions


def score_mrr(raw_positions, filtered_positions):
    """holE.py:475-490.  Hits are percentages of the *filtered* positions."""
    raw = np.array(raw_positions, dtype=np.float64)
    fil = np.array(filtered_positions, dtype=np.float64)
    return {
        "raw_mrr": float(np.mean(1.0 / raw)),
        "raw_mean_pos": float(np.mean(raw)),
        "filtered_mrr": float(np.mean(1.0 / fil)),
        "filtered_mean_pos": float(np.mean(fil)),
        "hits1": float(np.mean(fil <= 1) * 100),
        "hits3": float(np.mean(fil <= 3) * 100),
        "hits10": float(np.mean(fil <= 10) * 100),
    }
