"""Full-size parity against the oracle at BASELINE.json's sizes (the full-size configs were
only property-tested).

  config 0  FB15k shape, d=150, 16,296 rows: chained steps at B = 32768 (uniform and Zipf entities,
            Xavier and trained-scale tables) against oracle/hole_ref.c on identical corruption ids
  config 1  d=256, 1,200,014 rows, B = 32768, trained-scale table (clips fire): steps on both
            corruption sides against oracle/hole_ref.c
  config 2  the REAL 59,071 FB15k test queries x 2 sides x 14,951 candidates, filtered, on a table trained
            by the engine itself: filtered MRR within 1e-4 (split-bf16, the --infer default) and 5e-3 (plain
            bf16) of the fp64 oracle, Hits@1/3/10 within 1e-4 / 5e-3, counts inside the +-band of an fp64 contraction of the kernel's operands

Tolerances: loss <= 2e-6; rows <= 2e-6 + 1e-5 |x| per step against the fp32 C port (itself checked against
the reference-executed fixture in tests/test_tfshim_golden.py); corruption ids bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from graphembeddings_b200 import data as D
from oracle import hole_oracle as O
from oracle import hole_ref as R

pytestmark = pytest.mark.gpu
ROW_ATOL, ROW_RTOL = 2e-6, 1e-5


@pytest.fixture(scope="module")
def eng_mod():
    from graphembeddings_b200 import build
    build.build()
    from graphembeddings_b200 import engine
    return engine


def _steps_vs_c_port(eng_mod, kg, B, n_steps, seed, lr=0.1, margin=0.2):
    off, ids = D.build_type_csr(kg.type_of)
    e = eng_mod.HoleEngine(kg.n_rows, kg.dim).set_embeddings(kg.E).set_types(kg.type_of, off, ids)
    e.set_relation_count(kg.n_relations)
    E = np.ascontiguousarray(kg.E, np.float32).copy()
    sc = R.TrainScratch(B, kg.dim)
    sides = set()
    for s in range(n_steps):
        pos = kg.triples[s * B:(s + 1) * B]
        side, neg = e.corrupt_batch(pos, seed, s)
        oside, oneg = O.corrupt(pos, kg.type_of, off, ids, seed, s)
        assert side == oside and np.array_equal(neg.cpu().numpy(), oneg)          # bit-exact
        sides.add(side)
        loss = e.train_step(pos, neg, side, margin, lr).cpu().numpy()
        _, want = R.train_step(E, pos, oneg, oside, margin, lr, sc)
        assert np.abs(loss - want).max() <= 2e-6
        got = e.embeddings().cpu().numpy()
        err = np.abs(got - E)
        tol = (s + 1) * (ROW_ATOL + ROW_RTOL * np.abs(E))
        assert (err <= tol).all(), (s, float((err - tol).max()))
    touched = np.zeros(kg.n_rows, bool)
    touched[kg.triples[: n_steps * B].reshape(-1)] = True
    assert np.abs(got - kg.E).max() > 1e-4
    e.close()
    return sides, got, E


@pytest.mark.parametrize("zipf", [False, True])
@pytest.mark.parametrize("trained", [False, True])
def test_config0_full_size_steps_match_c_port(eng_mod, zipf, trained):
    """16,296 x 150 table, 4 chained steps of 32,768 triples: with Zipf entities the hottest row is used
    thousands of times per step (deep combine trees); trained-scale rows exercise the clip backward."""
    kg = D.make_config("fb15k_d150", n_triples=4 * 32768, zipf_entities=zipf, trained_scale=trained)
    sides, _, _ = _steps_vs_c_port(eng_mod, kg, 32768, 4, seed=7 if zipf else 3)
    assert len(sides) >= 1


def test_config1_full_size_steps_match_c_port(eng_mod):
    """1,200,014 x 256 trained-scale table (every clip branch runs), B = 32768, both corruption sides."""
    kg = D.make_config("diffbot_d256", n_triples=3 * 32768, trained_scale=True)
    norms = np.linalg.norm(kg.E[:4096].astype(np.float64), axis=1)
    assert (norms > 1).any() and (norms < 1).any()
    sides, _, _ = _steps_vs_c_port(eng_mod, kg, 32768, 3, seed=2)     # seed 2: steps 0..2 draw both sides
    assert sides == {0, 1}


def test_config1_device_loop_matches_c_port(eng_mod):
    """hole_train_steps (the loop the bench times: Philox corruption drawn on the device) at config 1's
    size against the C port fed with the oracle's Philox draws."""
    B, n_steps = 32768, 4
    kg = D.make_config("diffbot_d256", n_triples=n_steps * B, trained_scale=True)
    off, ids = D.build_type_csr(kg.type_of)
    e = eng_mod.HoleEngine(kg.n_rows, kg.dim).set_embeddings(kg.E).set_types(kg.type_of, off, ids)
    e.set_relation_count(kg.n_relations)
    lrs = [eng_mod.inverse_time_decay(0.1, s, 32 * 915, 0.5) for s in range(n_steps)]
    sums = e.train_steps(kg.triples, B, 5, 0, 0.2, lrs).cpu().numpy()
    E = np.ascontiguousarray(kg.E, np.float32).copy()
    sc = R.TrainScratch(B, kg.dim)
    for s in range(n_steps):
        pos = kg.triples[s * B:(s + 1) * B]
        side, neg = O.corrupt(pos, kg.type_of, off, ids, 5, s)
        tot, _ = R.train_step(E, pos, neg, side, 0.2, float(lrs[s]), sc)
        assert abs(sums[s] - tot) <= 2e-6 * B
    got = e.embeddings().cpu().numpy()
    assert (np.abs(got - E) <= n_steps * (ROW_ATOL + ROW_RTOL * np.abs(E))).all()
    e.close()


# ------------------------------------------------------------------------------------------ config 2
def _oracle_ranks(E64, queries, known, R_, N, side, chunk=2048):
    """fp64 oracle on the fp32 table: (raw_before, filt_before) for every query, tuple tie order
    (holE.py:434, 446-463), vectorised per chunk of queries."""
    name, col = ("tail", 1) if side == 0 else ("head", 0)
    cand = np.arange(R_, N)
    Y, _, _ = O.clip_rows(E64[cand])
    foff, fids = D.build_filter_csr(queries, known, name)
    raw = np.zeros(len(queries), np.int64)
    filt = np.zeros(len(queries), np.int64)
    for q0 in range(0, len(queries), chunk):
        q = queries[q0:q0 + chunk]
        S = O.query_vectors(E64, q, name, np.float64) @ Y.T
        tj = q[:, col].astype(np.int64) - R_
        thr = S[np.arange(len(q)), tj]
        before = (S < thr[:, None]) | ((S == thr[:, None]) & (cand[None, :] - R_ < tj[:, None]))
        raw[q0:q0 + len(q)] = before.sum(1)
        lo, hi = foff[q0], foff[q0 + len(q)]
        qi = np.repeat(np.arange(len(q)), np.diff(foff[q0:q0 + len(q) + 1]))
        fj = fids[lo:hi].astype(np.int64) - R_
        ok = (fj >= 0) & (fj < len(cand))
        hit = np.zeros(len(q), np.int64)
        np.add.at(hit, qi[ok], before[qi[ok], fj[ok]])
        filt[q0:q0 + len(q)] = raw[q0:q0 + len(q)] - hit
    return raw, filt, (foff, fids)


def test_config2_real_fb15k_queries_filtered_metrics_match_fp64_oracle(eng_mod, golden_dir):
    z = np.load(os.path.join(golden_dir, "fb15k_test_valid_triples.npz"))
    test, valid = z["test"], z["valid"]
    assert test.shape == (59071, 3) and valid.shape == (50000, 3)
    R_, N, dim = 1345, 16296, 150
    assert test[:, :2].min() >= R_ and test[:, :2].max() < N and test[:, 2].max() < R_
    # A table trained by the engine itself on the real test + valid triples (this is a numerical parity test,
    # not a held-out evaluation: fitting the queries gives the true candidates non-trivial ranks, so that
    # MRR / Hits are not the ~ln(N)/N of a random table); the filter sets are the valid triples + synthetic
    # "train" triples of the same shape (the real train file is not in the reference).
    kg = D.make_config("fb15k_d150", n_triples=100000, zipf_entities=True)
    fit = np.concatenate([test, valid]).astype(np.int32)
    off, ids = D.build_type_csr(kg.type_of)
    e = eng_mod.HoleEngine(N, dim).set_embeddings(kg.E).set_types(kg.type_of, off, ids).set_relation_count(R_)
    rng = np.random.default_rng(0)
    for epoch in range(60):
        perm = rng.permutation(len(fit))[: (len(fit) // 2048) * 2048]
        e.train_steps(fit[perm], 2048, 1, epoch * 100, 0.2, [0.5] * (len(perm) // 2048))
    train = np.concatenate([valid, kg.triples]).astype(np.int32)
    E = e.embeddings().cpu().numpy()
    E64 = E.astype(np.float64)
    known = train
    for side in (0, 1):
        oraw, ofilt, (foff, fids) = _oracle_ranks(E64, test, known, R_, N, side)
        want = O.score_mrr(oraw + 1, ofilt + 1)
        # Tolerances.  Split-bf16 (what `hole.py --infer` uses): MRR 1e-4, Hits@k 0.01 points, every sampled
        # count inside the 3e-5 band of the fp64 oracle.  Plain bf16 operands (what the throughput bench times):
        # MRR 5e-3, Hits@k 0.5 points -- measured on this table: |dMRR| 1.7e-3, |dHits@10| 0.11 (tail side) and
        # 0.38 points (head side); a score moves by up to 4e-3 |q||e| and the fitted queries' top ranks sit that
        # close together, so SURVEY 8c's suggested 1e-3 does not hold for bf16 here and is not claimed.
        for prec, tol_mrr, tol_hits, band in ((eng_mod.HOLE_RANK_BF16, 5e-3, 0.5, None),
                                              (eng_mod.HOLE_RANK_BF16X3, 1e-4, 0.01, 3e-5)):
            raw, filt, ts = e.rank(test, side, R_, N, foff, fids, precision=prec)
            raw, filt = raw.cpu().numpy().astype(np.int64), filt.cpu().numpy().astype(np.int64)
            got = O.score_mrr(raw + 1, filt + 1)
            assert abs(got["filtered_mrr"] - want["filtered_mrr"]) <= tol_mrr, (side, prec, got, want)
            assert abs(got["raw_mrr"] - want["raw_mrr"]) <= tol_mrr
            for k in ("hits1", "hits3", "hits10"):                       # percentages: 1e-3 absolute = 0.1 %
                assert abs(got[k] - want[k]) <= tol_hits, (side, prec, k, got[k], want[k])
            assert np.all(filt <= raw) and np.all(filt >= 0)
            if band is not None:           # split-bf16: counts inside the band of the fp64 oracle itself
                name, col = ("tail", 1) if side == 0 else ("head", 0)
                Y, _, _ = O.clip_rows(E64[R_:N])
                idx = np.arange(0, len(test), 29)                         # 2,037 sampled queries
                S = O.query_vectors(E64, test[idx], name, np.float64) @ Y.T
                thr = S[np.arange(len(idx)), test[idx, col].astype(np.int64) - R_]
                lo = (S < thr[:, None] - band).sum(1)
                hi = (S <= thr[:, None] + band).sum(1) - 1
                assert np.all(raw[idx] >= lo) and np.all(raw[idx] <= hi)
    # the metrics are not the trivial ones of a random table (ln(N)/N = 6e-4)
    assert want["filtered_mrr"] > 0.002
    e.close()


def test_config2_counts_inside_band_of_own_operands(eng_mod, golden_dir):
    """Same queries, plain bf16: every count lies inside the +-2e-6 band of an fp64 contraction of the
    kernel's OWN bf16 operands (the contraction + epilogue are exact up to fp32 accumulation noise)."""
    z = np.load(os.path.join(golden_dir, "fb15k_test_valid_triples.npz"))
    test = z["test"]
    R_, N, dim = 1345, 16296, 150
    kg = D.make_config("rank_fb15k_d150", trained_scale=True)
    e = eng_mod.HoleEngine(N, dim).set_embeddings(kg.E)
    for side, col in ((0, 1), (1, 0)):
        raw, filt, ts = e.rank(test, side, R_, N)
        cand, qp = e.rank_debug_operands()
        Cm = cand.double()[: N - R_]
        raw = raw.long()
        eps = 2e-6
        for q0 in range(0, len(test), 8192):
            q1 = min(len(test), q0 + 8192)
            S = qp[q0:q1].double() @ Cm.T
            tj = torch.as_tensor(test[q0:q1, col].astype(np.int64) - R_, device="cuda")
            thr = S[torch.arange(q1 - q0, device="cuda"), tj]
            assert float((ts[q0:q1].double() - thr).abs().max()) < eps
            lo = (S < (thr - eps)[:, None]).sum(1)
            hi = (S <= (thr + eps)[:, None]).sum(1) - 1
            assert bool(((raw[q0:q1] >= lo) & (raw[q0:q1] <= hi)).all())
    e.close()
