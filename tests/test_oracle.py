"""The oracle against everything that can pin it:
reference-generated ranking goldens, the Xavier constant, finite differences, an
independent torch-autograd restatement, and the HolE / GEMM-form identities."""
import json
import os
import re

import numpy as np
import pytest
import torch

from graphembeddings_b200 import data as D
from oracle import hole_oracle as O


def _kg(seed=3, dim=16, trained=True, n_triples=40):
    return D.synthetic_kg(n_relations=5, n_entities=60, n_triples=n_triples, n_types=4,
                          dim=dim, seed=seed, trained_scale=trained, zipf_entities=True)


# ---------------------------------------------------------------- pinned constants


def test_xavier_constant_from_archived_graph():
    # holE-20170724/graph.pbtxt:1791-1830: stddev const for the [1134637, 128] variable
    # (the graph stores it as an fp32 const, printed to 12 significant digits)
    assert abs(O.xavier_stddev(1134637, 128) - 0.00151367869694) < 1e-10
    assert abs(D.xavier_stddev(1134637, 128) - 0.00151367869694) < 1e-10
    assert float(np.float32(O.xavier_stddev(1134637, 128))) == float(np.float32(0.00151367869694))


def test_xavier_init_is_truncated():
    rng = np.random.default_rng(0)
    x = O.xavier_init(2000, 32, rng)
    sd = O.xavier_stddev(2000, 32)
    assert np.abs(x).max() <= 2 * sd * (1 + 1e-6)
    assert abs(x.std() / sd - 0.88) < 0.02          # std of a +-2 sigma truncated normal


def test_lr_schedule():
    # holE.py:292-294 with the defaults: lr0 .1, rate .5, decay_steps = 32 * batch_count
    assert O.inverse_time_decay(0.1, 0, 32 * 943, 0.5) == np.float32(0.1)
    got = O.inverse_time_decay(0.1, 32 * 943, 32 * 943, 0.5)
    assert abs(float(got) - 0.1 / 1.5) < 1e-8


# ---------------------------------------------------------------- score identities


def test_score_is_hole_circular_correlation_in_fourier_domain():
    """SURVEY section 0 surprise 1: r . ccorr(h, t) == (1/d) Re sum F(h) F(r) conj(F(t))...
    here the rows *are* the spectra, so s == d * <r, ccorr(h, t)> for real vectors whose
    FFTs are the stored complex rows.  Checked by building real vectors first."""
    rng = np.random.default_rng(1)
    d = 12
    h, t, r = rng.standard_normal((3, d))
    ccorr = np.array([sum(h[i] * t[(i + k) % d] for i in range(d)) for k in range(d)])
    hole = float(r @ ccorr)
    Fh, Ft, Fr = np.fft.fft(h), np.fft.fft(t), np.fft.fft(r)
    # holE.py:191: Re sum h * (r * conj(t)) with complex rows; HolE's r.ccorr(h,t) pairs
    # conj(F h) F t with conj(F r); taking the real part makes both conjugations agree.
    s = np.sum((np.conj(Fh) * Ft * np.conj(Fr)).real) / d
    assert abs(hole - s) < 1e-9
    # and the oracle's formula on rows [Re | Im] of (conj Fh, conj... ) equals the same sum
    rows = np.stack([np.concatenate([Ft.real, Ft.imag]),       # "h" slot
                     np.concatenate([Fh.real, Fh.imag]),       # "t" slot (conjugated in score)
                     np.concatenate([Fr.real, -Fr.imag])])     # "r" slot = conj(F r)
    rows64 = rows / 100.0                                       # norms < 1: no clip
    sc = O.score(rows64, np.array([[0, 1, 2]]), np.float64)[0] * 100.0 ** 3 / d
    assert abs(sc - hole) < 1e-6


def test_gemm_form_matches_pointwise_score():
    kg = _kg()
    E = kg.E.astype(np.float64)
    q = kg.triples[:10]
    cand = np.arange(kg.n_relations, kg.n_rows)
    for side, col in (("tail", 1), ("head", 0)):
        S = O.all_scores(E, q, side, cand, np.float64)
        for i in range(len(q)):
            for j in (0, 7, 31):
                tr = q[i].copy()
                tr[col] = cand[j]
                assert abs(S[i, j] - O.score(E, tr[None], np.float64)[0]) < 1e-12


def test_sigma_range_under_clip():
    kg = _kg(trained=True)
    v = O.evaluate_triples(kg.E, kg.triples)
    assert v.min() >= 0.2689 and v.max() <= 0.7311   # |s| <= 1 (App. A.1)


# ---------------------------------------------------------------- gradients


def _total_loss(E, pos, neg, side, margin):
    return float(O.evaluate_batch(E, pos, neg, side, margin, np.float64)[0].sum())


@pytest.mark.parametrize("side", [0, 1])
def test_gradient_finite_differences(side):
    kg = _kg(seed=4 + side, dim=8, n_triples=12)
    E = kg.E.astype(np.float64)
    pos = kg.triples
    rng = np.random.default_rng(5)
    neg = rng.integers(kg.n_relations, kg.n_rows, size=len(pos)).astype(np.int32)
    slices, _, _, _ = O.indexed_slices(E, pos, neg, side, 0.2, np.float64)
    G = np.zeros_like(E)
    for idx, g in slices:
        np.add.at(G, idx.astype(np.int64), g)
    eps = 1e-6
    worst = 0.0
    for row in np.unique(np.concatenate([pos.ravel(), neg]))[:12]:
        for k in range(E.shape[1]):
            Ep = E.copy(); Ep[row, k] += eps
            Em = E.copy(); Em[row, k] -= eps
            num = (_total_loss(Ep, pos, neg, side, 0.2) - _total_loss(Em, pos, neg, side, 0.2)) / (2 * eps)
            worst = max(worst, abs(num - G[row, k]))
    assert worst < 1e-8


def _torch_loss(E, pos, neg_triples, margin):
    """Independent restatement with torch autograd: holE.py:161-168, 191-192, 198, 231."""
    def emb(ids):
        x = E[ids]
        inv = torch.rsqrt((x * x).sum(dim=1, keepdim=True))
        y = x * torch.minimum(inv, torch.ones_like(inv))
        H = x.shape[1] // 2
        return torch.complex(y[:, :H], y[:, H:])

    def value(tr):
        h, t, r = emb(tr[:, 0]), emb(tr[:, 1]), emb(tr[:, 2])
        return torch.sigmoid((h * (r * torch.conj(t))).real.sum(dim=1))

    return torch.clamp(value(pos) - value(neg_triples) + margin, min=0).sum()


@pytest.mark.parametrize("side", [0, 1])
def test_gradient_matches_torch_autograd(side):
    kg = _kg(seed=9, dim=20, n_triples=64)
    E = kg.E.astype(np.float64)
    pos = kg.triples
    rng = np.random.default_rng(6)
    neg = rng.integers(kg.n_relations, kg.n_rows, size=len(pos)).astype(np.int32)
    slices, loss, _, _ = O.indexed_slices(E, pos, neg, side, 0.2, np.float64)
    G = np.zeros_like(E)
    for idx, g in slices:
        np.add.at(G, idx.astype(np.int64), g)
    Et = torch.tensor(E, requires_grad=True)
    L = _torch_loss(Et, torch.tensor(pos, dtype=torch.long),
                    torch.tensor(O.corrupt_triples(pos, neg, side), dtype=torch.long), 0.2)
    L.backward()
    assert abs(float(L) - float(loss.sum())) < 1e-12
    assert np.abs(Et.grad.numpy() - G).max() < 1e-12


def _torch_logloss(E, pos, neg_list, l2):
    """Independent torch-autograd restatement of the --log_loss branch: holE.py:161-168
    (max_norm lookup), 191-196 (score, -label * score, log(1 + exp), + l2 * l2_loss(embeddings)),
    206-221 (positives then negative_ratio corrupt batches, concatenated), 296 (sum)."""
    def emb(ids):
        x = E[ids]
        inv = torch.rsqrt((x * x).sum(dim=1, keepdim=True))
        y = x * torch.minimum(inv, torch.ones_like(inv))
        H = x.shape[1] // 2
        return torch.complex(y[:, :H], y[:, H:])

    def rows(tr, label):
        h, t, r = emb(tr[:, 0]), emb(tr[:, 1]), emb(tr[:, 2])
        s = (h * (r * torch.conj(t))).real.sum(dim=1)
        return torch.log(1.0 + torch.exp(-label * s)) + l2 * 0.5 * (E * E).sum()

    return torch.cat([rows(pos, 1.0)] + [rows(n, -1.0) for n in neg_list]).sum()


@pytest.mark.parametrize("k,l2", [(1, 0.0), (3, 0.0), (2, 1e-3)])
def test_logloss_step_matches_torch_autograd(k, l2):
    kg = _kg(seed=12, dim=20, n_triples=48)
    E = kg.E.astype(np.float64)
    pos = kg.triples
    rng = np.random.default_rng(8)
    negs = [rng.integers(kg.n_relations, kg.n_rows, size=len(pos)).astype(np.int32) for _ in range(k)]
    sides = [int(rng.integers(0, 2)) for _ in range(k)]
    lr = 0.05
    Et = torch.tensor(E, requires_grad=True)
    L = _torch_logloss(Et, torch.tensor(pos, dtype=torch.long),
                       [torch.tensor(O.corrupt_triples(pos, n, sd), dtype=torch.long) for n, sd in zip(negs, sides)], l2)
    L.backward()
    want = E - lr * Et.grad.numpy()
    got = E.copy()
    loss, l2_loss = O.logloss_step(got, pos, negs, sides, lr, l2, np.float64)
    assert loss.shape == (1 + k, len(pos))
    assert np.abs(got - want).max() < 1e-12
    assert abs(float(loss.sum() + loss.size * l2 * l2_loss) - float(L.detach())) < 1e-9 * max(1.0, abs(float(L.detach())))
    # fp32 follows fp64
    got32 = kg.E.astype(np.float32).copy()
    O.logloss_step(got32, pos, negs, sides, lr, l2, np.float32)
    assert np.abs(got32 - want).max() < 5e-6


def test_sgd_orders_agree_and_duplicates_accumulate():
    kg = _kg(seed=12, dim=16, n_triples=200)       # 60 entities, 200 triples: many duplicates
    pos = kg.triples
    off, ids = O.build_type_csr(kg.type_of)
    side, neg = O.corrupt(pos, kg.type_of, off, ids, seed=1, step=0)
    E_tf = kg.E.astype(np.float64); E_m = E_tf.copy()
    l1, _, _ = O.sgd_step(E_tf, pos, neg, side, 0.2, 0.1, np.float64, "tf")
    l2, _, _ = O.sgd_step(E_m, pos, neg, side, 0.2, 0.1, np.float64, "merged")
    assert np.array_equal(l1, l2)
    assert np.abs(E_tf - E_m).max() < 1e-13
    E32 = kg.E.copy(); E32m = kg.E.copy()
    O.sgd_step(E32, pos, neg, side, 0.2, 0.1, np.float32, "tf")
    O.sgd_step(E32m, pos, neg, side, 0.2, 0.1, np.float32, "merged")
    assert np.abs(E32 - E_tf).max() < 2e-6
    assert np.abs(E32m - E_tf).max() < 2e-6
    assert np.abs(E_tf - kg.E).max() > 1e-3        # the step moved something


def test_train_step_golden_fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "train_step_oracle.npz"))
    off, ids = O.build_type_csr(z["type_of"])
    for step in (0, 1):
        side, neg = O.corrupt(z["pos"], z["type_of"], off, ids, seed=5, step=step)
        assert side == int(z[f"side{step}"])
        assert np.array_equal(neg, z[f"neg{step}"])
        E = z["E0"].copy()
        loss, vp, vn = O.sgd_step(E, z["pos"], neg, side, 0.2, 0.1, np.float32, "tf")
        np.testing.assert_allclose(loss, z[f"loss{step}"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(E, z[f"E32_{step}"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(E, z[f"E64_{step}"], rtol=0, atol=2e-6)


def test_logloss_step_golden_fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "logloss_step_oracle.npz"))
    off, ids = O.build_type_csr(z["type_of"])
    k, seed, step = int(z["k"]), int(z["seed"]), int(z["step"])
    for j in range(k):
        side, neg = O.corrupt(z["pos"], z["type_of"], off, ids, seed=seed, step=step * k + j)
        assert side == int(z["sides"][j]) and np.array_equal(neg, z["negs"][j])
    E = z["E0"].copy()
    loss, l2_loss = O.logloss_step(E, z["pos"], list(z["negs"]), [int(x) for x in z["sides"]],
                                   float(z["lr"]), float(z["l2"]), np.float32)
    np.testing.assert_allclose(loss, z["loss32"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(E, z["E32"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(E, z["E64"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(loss, z["loss64"], rtol=0, atol=1e-6)
    assert abs(float(l2_loss) - float(z["l2_64"])) <= 1e-6 * float(z["l2_64"])


# ---------------------------------------------------------------- corruption


def test_corruption_is_type_safe_and_relations_untouched():
    kg = D.synthetic_kg(n_relations=9, n_entities=500, n_triples=1000, n_types=7, dim=8,
                        seed=2, with_embeddings=False)
    off, ids = O.build_type_csr(kg.type_of)
    sides = set()
    for step in range(6):
        side, neg = O.corrupt(kg.triples, kg.type_of, off, ids, seed=77, step=step)
        sides.add(side)
        replaced = kg.triples[:, 0 if side else 1]
        assert np.array_equal(kg.type_of[neg], kg.type_of[replaced])
        assert neg.min() >= kg.n_relations          # never a relation row (type 0 is theirs)
    assert sides == {0, 1}


# ---------------------------------------------------------------- ranking vs the reference's own code


def _load_ranking_golden(golden_dir):
    with open(os.path.join(golden_dir, "ranking_ref.json")) as f:
        return json.load(f)


def _dd(d):
    return {int(h): {int(r): set(ts) for r, ts in rr.items()} for h, rr in d.items()}


def test_heap_restatement_matches_reference_outputs(golden_dir):
    g = _load_ranking_golden(golden_dir)
    assert len(g["cases"]) >= 10
    for c in g["cases"]:
        vals = np.array(c["values"], dtype=np.float32)
        raw, filt = O.eval_link_prediction_heap(vals, c["triples"], _dd(c["true_triples"]),
                                                _dd(c["test_triples"]))
        assert raw == c["raw_positions"]
        assert filt == c["filtered_positions"]


def test_rank_counts_matches_reference_outputs(golden_dir):
    """The vectorised count form (what the GPU kernel computes) reproduces the reference's
    heap ranks for single-relation groups; multi-relation groups are checked through the
    (tail, relation) composite id."""
    g = _load_ranking_golden(golden_dir)
    checked = 0
    for c in g["cases"]:
        triples = np.array(c["triples"])
        vals = np.array(c["values"], dtype=np.float32)
        true_t, test_t = _dd(c["true_triples"]), _dd(c["test_triples"])
        # composite candidate id keeps the heap's (tail, relation) tuple order
        comp = triples[:, 1].astype(np.int64) * 1000 + triples[:, 2]
        order = np.argsort(comp)
        comp, vals_s, tr_s = comp[order], vals[order], triples[order]
        in_sample = [int(cid) for cid, (h, t, r) in zip(comp, tr_s)
                     if t in true_t.get(int(h), {}).get(int(r), ())]
        queries = [(k, int(cid)) for k, (cid, (h, t, r)) in enumerate(zip(comp, tr_s))
                   if t in test_t.get(int(h), {}).get(int(r), ()) and int(cid) not in in_sample]
        if not queries:
            continue
        S = np.tile(vals_s, (len(queries), 1))
        raw, filt = O.rank_counts(S, comp, [cid for _, cid in queries],
                                  [in_sample] * len(queries))
        got = sorted(zip((raw + 1).tolist(), (filt + 1).tolist()))
        want = sorted(zip(c["raw_positions"], c["filtered_positions"]))
        assert got == want
        checked += len(queries)
    assert checked > 100


def test_score_mrr_matches_reference_stdout(golden_dir):
    g = _load_ranking_golden(golden_dir)
    raw = [x for c in g["cases"] for x in c["raw_positions"]]
    filt = [x for c in g["cases"] for x in c["filtered_positions"]]
    m = O.score_mrr(raw, filt)
    nums = [float(x) for x in re.findall(r"[-+]?\d+\.\d+(?:e[-+]?\d+)?", g["score_mrr_stdout"])]
    want = dict(zip(["raw_mrr", "raw_mean_pos", "filtered_mrr", "filtered_mean_pos",
                     "hits1", "hits3", "hits10"], nums))
    for k, v in want.items():
        assert abs(m[k] - v) < 1e-12, k
