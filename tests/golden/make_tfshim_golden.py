#!/usr/bin/env python
"""Generate tests/golden/tfshim_step.npz by EXECUTING the reference's own hot-path source
(/root/reference/holE.py: corrupt_batch, get_embedding, evaluate_triples, evaluate_batch, and
GradientDescentOptimizer.minimize with inverse_time_decay) on top of tests/golden/tfshim.py.

Run in the build container only (needs /root/reference):  python tests/golden/make_tfshim_golden.py

The corruption ids are the Philox oracle's (oracle/philox.py); they are fed to the reference's
own corrupt_heads / corrupt_tails through its two hash tables and tf.random_uniform, so the
reference code and every implementation under test see identical triples and identical
corruption indices (north_star's parity condition).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import tfshim  # noqa: E402
from graphembeddings_b200 import data as D  # noqa: E402
from oracle import hole_oracle as O  # noqa: E402


def flags(dim, **kw):
    f = dict(embedding_dim=dim, log_loss=False, l2_regularization=0.1, negative_ratio=1, margin=0.2,
             padded_size=8, learning_rate=0.1, learning_decay_steps=32, learning_decay_rate=0.5)
    f.update(kw)
    return types.SimpleNamespace(**f)


def feed_corruption(ref, tf, kg, off, ids, pos, draws):
    """Fill the reference's two hash tables (holE.py:37-41, 267-277, 322-326, 343-347) and queue
    the tf.random_uniform values so that its corrupt_batch reproduces `draws` =
    [(side, neg_ent[B]), ...] exactly.  Returns (type_to_ids_table, id_to_type_table)."""
    type_names = {t: "T%d" % t for t in np.unique(kg.type_of)}
    P = int(np.diff(off).max())
    ref.FLAGS.padded_size = P
    type_to_ids = ref.init_table(tf.string, tf.int64, 'type_to_ids', type_to_ids=True)
    id_to_type = ref.init_table(tf.int64, tf.string, 'id_to_type')
    id_to_type.insert(np.arange(kg.n_rows, dtype=np.int64),
                      np.array([type_names[t] for t in kg.type_of], dtype=object))
    # holE.py:343-344 draws P ids per type with replacement; here row = the type's full id list,
    # cycled up to P (every id present, so any wanted replacement can be indexed)
    rows = np.stack([np.resize(ids[off[t]:off[t + 1]], P) for t in sorted(type_names)])
    type_to_ids.insert(np.array([type_names[t] for t in sorted(type_names)], dtype=object), rows.astype(np.int64))
    fed = []
    for side, neg in draws:
        fed.append(0.25 if side else 0.75)              # should_corrupt_heads = u < 0.5 (holE.py:137)
        ent = pos[:, 0 if side else 1]
        ty = kg.type_of[ent]
        fed.append(np.array([int(np.flatnonzero(ids[off[t]:off[t + 1]] == n)[0]) for t, n in zip(ty, neg)],
                            dtype=np.int32))
    tf.feed_random(fed)
    return type_to_ids, id_to_type


def run_case(ref, tf, name, kg, pos, seed, steps, dtype, out, log_loss=False, k=1, l2=0.0, lr0=0.1,
             batch_count=100, margin=0.2):
    tf.set_float(dtype)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    off, ids = O.build_type_csr(kg.type_of)
    ref.FLAGS = flags(kg.dim, log_loss=log_loss, negative_ratio=k, l2_regularization=l2, learning_rate=lr0,
                      margin=margin)
    emb = tf.Variable(torch.as_tensor(kg.E.astype(dtype)), name="embeddings")
    gstep = tf.Variable(torch.zeros((), dtype=torch.int32), name="global_step")
    tag = "%s_%s" % (name, "f64" if dtype == np.float64 else "f32")
    tri = torch.as_tensor(pos, dtype=torch.int32)
    for s in range(steps):
        if log_loss:
            draws = [O.corrupt(pos, kg.type_of, off, ids, seed, s * k + j) for j in range(k)]
        else:
            draws = [O.corrupt(pos, kg.type_of, off, ids, seed, s)]
        t2i, i2t = feed_corruption(ref, tf, kg, off, ids, pos, draws)
        # holE.py:292-296
        lr = tf.train.inverse_time_decay(ref.FLAGS.learning_rate, gstep,
                                         decay_steps=ref.FLAGS.learning_decay_steps * batch_count,
                                         decay_rate=ref.FLAGS.learning_decay_rate)
        if not log_loss:
            # sigma of both sides, through the reference's evaluate_triples (holE.py:179-202)
            side, neg = draws[0]
            with torch.no_grad():
                vp = ref.evaluate_triples(tri, emb).numpy().reshape(-1)
                vn = ref.evaluate_triples(torch.as_tensor(O.corrupt_triples(pos, neg, side), dtype=torch.int32),
                                          emb).numpy().reshape(-1)
            emb.slices = []
            out["%s_vp%d" % (tag, s)], out["%s_vn%d" % (tag, s)] = vp, vn
        loss = ref.evaluate_batch(tri, emb, t2i, i2t, kg.n_relations)       # holE.py:205-234
        assert not tf._STATE["random"], "the reference did not consume every fed random value"
        tf.train.GradientDescentOptimizer(lr).minimize(loss, gstep)         # holE.py:296
        l = loss.detach().numpy()
        if log_loss:
            # every row carries the same scalar l2 * l2_loss(E_old) (holE.py:196); store it apart
            l2_loss = float(0.5 * np.sum(np.asarray(out.get("_E_before", kg.E), dtype=np.float64) ** 2))
            out["%s_l2loss%d" % (tag, s)] = np.asarray(l2_loss)
            out["%s_loss%d" % (tag, s)] = l.reshape(1 + k, len(pos))
            out["%s_sides%d" % (tag, s)] = np.array([d[0] for d in draws])
            out["%s_negs%d" % (tag, s)] = np.stack([d[1] for d in draws])
        else:
            out["%s_loss%d" % (tag, s)] = l.reshape(-1)
            out["%s_side%d" % (tag, s)] = np.asarray(draws[0][0])
            out["%s_neg%d" % (tag, s)] = draws[0][1]
        out["%s_lr%d" % (tag, s)] = np.asarray(lr, dtype=np.float32)
        out["%s_E%d" % (tag, s)] = emb.numpy().copy()
        out["_E_before"] = emb.numpy().copy()
    out.pop("_E_before", None)
    assert int(gstep.t) == steps


def main():
    tf = tfshim.install()
    ref = tfshim.import_reference_hole()
    out = {}
    # A: clipped rows (trained scale), Zipf entities (heavy duplicates), one row of norm exactly 1
    kgA = D.synthetic_kg(n_relations=7, n_entities=300, n_triples=96, n_types=5, dim=20, seed=21,
                         trained_scale=True, zipf_entities=True)
    hot = int(np.bincount(kgA.triples[:, :2].reshape(-1)).argmax())
    kgA.E[hot] = 0.0
    kgA.E[hot, 3] = -1.0                       # rsqrt(sum x^2) == 1: Minimum gradient tie -> clip branch
    # B: Xavier-scale rows (no clip), d = 150 (odd half), uniform entities
    kgB = D.synthetic_kg(n_relations=11, n_entities=500, n_triples=128, n_types=6, dim=150, seed=22)
    # C: --log_loss, k = 2, l2 = 1e-3
    kgC = D.synthetic_kg(n_relations=7, n_entities=300, n_triples=64, n_types=5, dim=20, seed=23,
                         trained_scale=True, zipf_entities=True)
    # D: margin 0.01 on a trained-scale table: about half of the hinges are inactive (zero rows in
    # the IndexedSlices, rows left untouched)
    kgD = D.synthetic_kg(n_relations=5, n_entities=200, n_triples=80, n_types=3, dim=32, seed=24,
                         trained_scale=True)
    for dtype in (np.float32, np.float64):
        run_case(ref, tf, "D", kgD, kgD.triples, seed=8, steps=2, dtype=dtype, out=out, margin=0.01)
        run_case(ref, tf, "A", kgA, kgA.triples, seed=5, steps=3, dtype=dtype, out=out)
        run_case(ref, tf, "B", kgB, kgB.triples, seed=6, steps=2, dtype=dtype, out=out)
        run_case(ref, tf, "C", kgC, kgC.triples, seed=7, steps=2, dtype=dtype, out=out, log_loss=True, k=2,
                 l2=1e-3, lr0=0.05)
    for nm, kg in (("A", kgA), ("B", kgB), ("C", kgC), ("D", kgD)):
        out[nm + "_E0"], out[nm + "_pos"], out[nm + "_type_of"] = kg.E, kg.triples, kg.type_of
        out[nm + "_n_relations"] = np.asarray(kg.n_relations)
    out["generator"] = np.asarray("tests/golden/make_tfshim_golden.py: /root/reference/holE.py executed on tests/golden/tfshim.py")
    np.savez_compressed(os.path.join(HERE, "tfshim_step.npz"), **out)
    sides = {k: np.asarray(v).tolist() for k, v in out.items() if "_side" in k and "f64" in k}
    print("tfshim_step.npz written:", len(out), "arrays; sides", sides)


if __name__ == "__main__":
    main()
