"""The C/OpenMP restatement (bench.py's CPU baseline) against the NumPy oracle."""
import numpy as np

from graphembeddings_b200 import data as D
from oracle import hole_oracle as O
from oracle import hole_ref as R


def test_score_matches_numpy_oracle():
    kg = D.synthetic_kg(6, 400, 500, 4, 150, seed=3, trained_scale=True)
    got = R.score(kg.E, kg.triples)
    want = O.evaluate_triples(kg.E.astype(np.float64), kg.triples, np.float64)
    assert np.abs(got - want).max() < 1e-6


def test_train_step_matches_numpy_oracle_both_sides():
    kg = D.synthetic_kg(6, 300, 2000, 4, 64, seed=4, trained_scale=True, zipf_entities=True)
    off, ids = O.build_type_csr(kg.type_of)
    for side_force in (0, 1):
        _, neg = O.corrupt(kg.triples, kg.type_of, off, ids, 3, side_force)
        E = kg.E.copy()
        tot, loss = R.train_step(E, kg.triples, neg, side_force, 0.2, 0.1)
        E64 = kg.E.astype(np.float64)
        l64, _, _ = O.sgd_step(E64, kg.triples, neg, side_force, 0.2, 0.1, np.float64, "tf")
        assert np.abs(loss - l64).max() < 2e-6
        assert abs(tot - l64.sum()) < 1e-3
        assert np.abs(E - E64).max() < 1e-5
        assert np.abs(E - kg.E).max() > 1e-4


def test_logloss_step_matches_numpy_oracle():
    """Second, independent implementation of the --log_loss step (holE.py:194-196, 206-220)."""
    kg = D.synthetic_kg(6, 300, 1500, 4, 64, seed=6, trained_scale=True, zipf_entities=True)
    off, ids = O.build_type_csr(kg.type_of)
    for k, l2 in ((1, 0.0), (3, 1e-5)):
        drawn = [O.corrupt(kg.triples, kg.type_of, off, ids, 9, 4 * k + j) for j in range(k)]
        sides, negs = [d[0] for d in drawn], [d[1] for d in drawn]
        E = kg.E.copy()
        loss, l2_loss = R.logloss_step(E, kg.triples, negs, sides, 0.05, l2)
        E64 = kg.E.astype(np.float64)
        loss64, l2_64 = O.logloss_step(E64, kg.triples, negs, sides, 0.05, l2, np.float64)
        assert np.abs(loss - loss64).max() < 2e-6
        assert abs(l2_loss - float(l2_64)) < 1e-6 * float(l2_64)
        assert np.abs(E - E64).max() < 2e-5
        assert np.abs(E - kg.E).max() > 1e-4


def test_rank_matches_numpy_oracle():
    kg = D.synthetic_kg(5, 600, 80, 3, 32, seed=5, trained_scale=True)
    cand = np.arange(kg.n_relations, kg.n_rows)
    known = D.synthetic_kg(5, 600, 3000, 3, 32, seed=6, with_embeddings=False).triples
    for side, col in (("tail", 1), ("head", 0)):
        foff, fids = D.build_filter_csr(kg.triples, known, side)
        S = O.all_scores(kg.E, kg.triples, side, cand, np.float32)
        flists = [fids[foff[q]:foff[q + 1]] for q in range(len(kg.triples))]
        raw, filt = O.rank_counts(S, cand, kg.triples[:, col], flists)
        Y = R.clip_rows(kg.E[kg.n_relations:])
        qv = O.query_vectors(kg.E, kg.triples, side, np.float32)
        graw, gfilt = R.rank(Y, kg.n_relations, qv, kg.triples[:, col], foff, fids)
        # fp32 dot products in a different order: ranks may differ only at near-ties
        assert np.abs(graw - raw).max() <= 1 and np.mean(graw == raw) > 0.97
        assert np.abs(gfilt - filt).max() <= 1
