"""Property tests (hypothesis) of the oracle and of the host logic -- CPU only.

They pin size-independent facts the GPU parity tests rely on: the corruption sampler is
type-safe for any (seed, step); the ranking count form equals the reference's heap walk on
random scores with ties; the filter CSR equals a brute-force set construction; the score is
invariant under the Hermitian swap; the checkpoint bundle round-trips."""
import os

import numpy as np
from hypothesis import given, settings, strategies as st

from graphembeddings_b200 import data as D
from graphembeddings_b200 import tf_bundle
from oracle import hole_oracle as O
from oracle import philox

FAST = settings(max_examples=25, deadline=None)


def _kg(seed, dim=12, n_ent=60, n_rel=4, n_tri=40, n_types=3):
    return D.synthetic_kg(n_rel, n_ent, n_tri, n_types, dim, seed=seed, trained_scale=True)


@FAST
@given(seed=st.integers(0, 2**63 - 1), step=st.integers(0, 2**40), kgseed=st.integers(0, 50))
def test_corruption_type_safe_for_any_seed_and_step(seed, step, kgseed):
    kg = _kg(kgseed)
    off, ids = O.build_type_csr(kg.type_of)
    side, neg = O.corrupt(kg.triples, kg.type_of, off, ids, seed, step)
    assert side in (0, 1) and side == philox.side_coin(seed, step)
    replaced = kg.triples[:, 0 if side else 1]
    assert np.array_equal(kg.type_of[neg], kg.type_of[replaced])       # same type (holE.py:111-131)
    assert (neg >= kg.n_relations).all() and (neg < kg.n_rows).all()    # never a relation row
    # deterministic in (seed, step, index): a permuted batch draws per position, not per content
    side2, neg2 = O.corrupt(kg.triples, kg.type_of, off, ids, seed, step)
    assert side2 == side and np.array_equal(neg2, neg)


@FAST
@given(seed=st.integers(0, 2**32 - 1), cnt=st.integers(1, 2**31 - 1), n=st.integers(1, 64))
def test_entity_draw_is_in_range(seed, cnt, n):
    j = philox.entity_draw(seed, 7, np.arange(n), np.full(n, cnt, dtype=np.int64))
    assert (j >= 0).all() and (j < cnt).all()


@FAST
@given(seed=st.integers(0, 10_000), levels=st.integers(2, 6))
def test_rank_counts_equal_heap_walk_with_ties(seed, levels):
    """Quantised scores force exact ties; the count form must reproduce the (value, id) tuple
    order of the reference's heap (holE.py:434, 446-463), filtered ranks included."""
    rng = np.random.default_rng(seed)
    n_cand, Q = 30, 6
    cand = np.arange(5, 5 + n_cand)
    S = rng.integers(0, levels, size=(Q, n_cand)).astype(np.float32) / levels
    h, r = 1, 2
    true_ids = rng.choice(cand, size=Q, replace=False)
    known = [set(int(x) for x in rng.choice(cand, size=rng.integers(0, 6), replace=False)) - {int(t)}
             for t in true_ids]
    raw, filt = O.rank_counts(S, cand, true_ids, known)
    for q in range(Q):
        triples = [(h, int(c), r) for c in cand]
        true_t = {h: {r: set(known[q])}}
        rr, ff = O.eval_link_prediction_heap(S[q], triples, true_t, {h: {r: {int(true_ids[q])}}})
        assert rr == [int(raw[q]) + 1] and ff == [int(filt[q]) + 1]


@FAST
@given(seed=st.integers(0, 10_000))
def test_filter_csr_equals_brute_force(seed):
    rng = np.random.default_rng(seed)
    known = np.stack([rng.integers(0, 8, 60), rng.integers(0, 8, 60), rng.integers(0, 3, 60)], axis=1)
    queries = np.stack([rng.integers(0, 8, 10), rng.integers(0, 8, 10), rng.integers(0, 3, 10)], axis=1)
    for side, qcol, vcol in (("tail", 0, 1), ("head", 1, 0)):
        off, ids = D.build_filter_csr(queries, known, side)
        for q, row in enumerate(queries):
            want = sorted({int(k[vcol]) for k in known if k[qcol] == row[qcol] and k[2] == row[2]})
            assert list(ids[off[q]:off[q + 1]]) == want


@FAST
@given(seed=st.integers(0, 10_000))
def test_score_hermitian_swap(seed):
    """Re<h, r conj(t)> = Re<t, conj(r) conj(h)>: swapping head and tail and conjugating the
    relation leaves the score unchanged (the identity behind the head-side query vector)."""
    kg = _kg(seed, dim=16)
    E = kg.E.astype(np.float64)
    H = E.shape[1] // 2
    Ec = E.copy()
    Ec[:kg.n_relations, H:] *= -1.0                     # conjugate the relation rows
    swapped = kg.triples[:, [1, 0, 2]]
    np.testing.assert_allclose(O.score(E, kg.triples, np.float64), O.score(Ec, swapped, np.float64),
                               rtol=0, atol=1e-12)
    # and the GEMM form agrees with the pointwise score on both sides
    cand = np.arange(kg.n_relations, kg.n_rows)
    for side, col in (("tail", 1), ("head", 0)):
        S = O.all_scores(E, kg.triples, side, cand, np.float64)
        got = S[np.arange(len(kg.triples)), kg.triples[:, col] - kg.n_relations]
        np.testing.assert_allclose(got, O.score(E, kg.triples, np.float64), rtol=0, atol=1e-12)


@FAST
@given(n=st.integers(1, 40), d=st.integers(1, 33), seed=st.integers(0, 1000), step=st.integers(0, 2**31 - 1))
def test_bundle_roundtrip(tmp_path_factory, n, d, seed, step):
    rng = np.random.default_rng(seed)
    E = rng.standard_normal((n, d)).astype(np.float32)
    prefix = os.path.join(str(tmp_path_factory.mktemp("b")), "model.ckpt")
    tf_bundle.save_bundle(prefix, {"embeddings": E, "batch/Variable": np.array(step, dtype=np.int32)})
    b = tf_bundle.load_bundle(prefix)
    assert np.array_equal(b["embeddings"], E) and int(np.asarray(b["batch/Variable"]).reshape(-1)[0]) == step


@FAST
@given(seed=st.integers(0, 10_000), k=st.integers(1, 3))
def test_logloss_step_with_zero_lr_is_identity_and_loss_is_softplus(seed, k):
    kg = _kg(seed, dim=16)
    rng = np.random.default_rng(seed)
    negs = [rng.integers(kg.n_relations, kg.n_rows, size=len(kg.triples)).astype(np.int32) for _ in range(k)]
    sides = [int(rng.integers(0, 2)) for _ in range(k)]
    E = kg.E.astype(np.float64)
    E0 = E.copy()
    loss, l2 = O.logloss_step(E, kg.triples, negs, sides, 0.0, 0.3, np.float64)
    assert np.array_equal(E, E0)
    s = O.score(E0, kg.triples, np.float64)
    np.testing.assert_allclose(loss[0], np.logaddexp(0.0, -s), rtol=0, atol=1e-12)
    for j in range(k):
        sn = O.score(E0, O.corrupt_triples(kg.triples, negs[j], sides[j]), np.float64)
        np.testing.assert_allclose(loss[1 + j], np.logaddexp(0.0, sn), rtol=0, atol=1e-12)
    assert abs(float(l2) - 0.5 * float((E0 ** 2).sum())) < 1e-9
