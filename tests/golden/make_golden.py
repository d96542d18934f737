#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ (run in the BUILD container only).

1. ranking_ref.json -- outputs of the REFERENCE's own pure-Python ranking / metric code
   (holE.py:427-490 ``eval_link_prediction`` and ``score_mrr``), executed by importing
   /root/reference/holE.py with the ``tensorflow`` import stubbed out (those two functions
   touch only heapq, numpy, print and FLAGS.infer_threshold).  Inputs are seeded random
   sigma values; they are stored with the outputs so tests need neither the reference
   nor this script.
2. graphembeddings_b200/fb15k_types.json (package data) -- the type histogram of diffbot_data/FB15k/entity_metadata.tsv
   (815 classes, 372 singletons) used by the synthetic FB15k-shape generator, and
   fb15k_head.tsv files: the first lines of the real triple/metadata files as loader
   fixtures (data, not source).
3. train_step_oracle.npz -- oracle outputs (loss, sigma, updated rows) for one seeded
   step.  The oracle is unpinned for this arithmetic (TF absent); the fixture guards the
   oracle against silent drift, and gives the GPU tests a reference-free comparison point.

Usage: python tests/golden/make_golden.py
"""
import contextlib
import importlib.util
import io
import json
import os
import sys
import types
from collections import defaultdict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def import_reference_hole():
    """Import /root/reference/holE.py with TensorFlow stubbed (it is not installable)."""
    tf = types.ModuleType("tensorflow")
    contrib = types.ModuleType("tensorflow.contrib")
    tb = types.ModuleType("tensorflow.contrib.tensorboard")
    plugins = types.ModuleType("tensorflow.contrib.tensorboard.plugins")
    projector = types.ModuleType("tensorflow.contrib.tensorboard.plugins.projector")
    plugins.projector = projector
    tb.plugins = plugins
    contrib.tensorboard = tb
    tf.contrib = contrib
    for name, mod in [("tensorflow", tf), ("tensorflow.contrib", contrib),
                      ("tensorflow.contrib.tensorboard", tb),
                      ("tensorflow.contrib.tensorboard.plugins", plugins),
                      ("tensorflow.contrib.tensorboard.plugins.projector", projector)]:
        sys.modules[name] = mod
    spec = importlib.util.spec_from_file_location("ref_holE", os.path.join(REF, "holE.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def gen_ranking_golden():
    ref = import_reference_hole()
    ref.FLAGS = types.SimpleNamespace(infer_threshold=10.0)  # gate always open
    rng = np.random.default_rng(20170903)
    cases = []
    cwd = os.getcwd()
    os.chdir("/tmp")  # eval_link_prediction appends inference_results.tsv to CWD
    try:
        for case in range(12):
            n_cand = int(rng.integers(5, 60))
            n_rel = int(rng.integers(1, 4))
            head = int(rng.integers(0, 100))
            tails = sorted(int(x) for x in rng.choice(1000, size=n_cand, replace=False))
            rels = sorted(int(x) for x in rng.choice(20, size=n_rel, replace=False))
            triples = [(head, t, r) for t in tails for r in rels]
            # quantised values so exact ties occur and the tuple tie-break is exercised
            vals = np.round(rng.uniform(0.27, 0.73, size=len(triples)), 2).astype(np.float32)
            true_triples = defaultdict(lambda: defaultdict(set))
            test_triples = defaultdict(lambda: defaultdict(set))
            for (h, t, r) in triples:
                u = rng.random()
                if u < 0.15:
                    true_triples[h][r].add(t)
                if 0.10 < u < 0.35:  # overlap on purpose: test tails that are also in-sample
                    test_triples[h][r].add(t)
            raw, filt = [], []
            id_to_metadata = defaultdict(lambda: "x")
            with contextlib.redirect_stdout(io.StringIO()):
                ref.eval_link_prediction(zip(vals.reshape(-1, 1), triples), id_to_metadata,
                                         true_triples, test_triples, 3, raw, filt)
            cases.append({
                "values": [float(v) for v in vals],
                "triples": [list(t) for t in triples],
                "true_triples": {str(h): {str(r): sorted(ts) for r, ts in d.items()}
                                 for h, d in true_triples.items()},
                "test_triples": {str(h): {str(r): sorted(ts) for r, ts in d.items()}
                                 for h, d in test_triples.items()},
                "raw_positions": [int(x) for x in raw],
                "filtered_positions": [int(x) for x in filt],
            })
        # score_mrr prints its results; capture and parse the printed numbers
        all_raw = [x for c in cases for x in c["raw_positions"]]
        all_filt = [x for c in cases for x in c["filtered_positions"]]
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ref.score_mrr(all_raw, all_filt)
        printed = buf.getvalue()
    finally:
        os.chdir(cwd)
        with contextlib.suppress(OSError):
            os.remove("/tmp/inference_results.tsv")
    out = {"generator": "tests/golden/make_golden.py via /root/reference/holE.py:427-490",
           "cases": cases, "score_mrr_stdout": printed}
    with open(os.path.join(HERE, "ranking_ref.json"), "w") as f:
        json.dump(out, f)
    print("ranking_ref.json:", len(cases), "cases;", printed.strip().replace("\n", " | "))


def gen_fb15k_fixtures():
    meta = os.path.join(REF, "diffbot_data/FB15k/entity_metadata.tsv")
    types_count = defaultdict(int)
    n_rel = 0
    with open(meta) as f:
        header = next(f)
        lines = []
        for i, line in enumerate(f):
            cols = line.rstrip("\n").split("\t")
            if cols[3] == "RELATION":
                n_rel += 1
            else:
                types_count[cols[3]] += 1
            if i < 40 or 1340 <= i < 1400:
                lines.append(line)
    hist = sorted(types_count.values(), reverse=True)
    with open(os.path.join(ROOT, "graphembeddings_b200", "fb15k_types.json"), "w") as f:
        json.dump({"source": "diffbot_data/FB15k/entity_metadata.tsv col 4",
                   "n_relation_rows": n_rel, "entity_type_histogram": hist}, f)
    print("fb15k_types.json:", n_rel, "relations,", len(hist), "types,", sum(hist), "entities,",
          sum(1 for h in hist if h == 1), "singletons")
    with open(os.path.join(HERE, "fb15k_metadata_head.tsv"), "w") as f:
        f.write(header)
        f.writelines(lines)
    for name in ("test_positive_triples.txt", "triples-valid.txt"):
        with open(os.path.join(REF, "diffbot_data/FB15k", name)) as src, \
                open(os.path.join(HERE, "fb15k_" + name.replace(".txt", "_head.txt")), "w") as dst:
            for i, line in enumerate(src):
                if i >= 2000:
                    break
                dst.write(line)


def gen_fb15k_full_triples():
    """The complete id-mapped FB15k test / valid triple files (59,071 / 50,000 rows: BASELINE config 2's
    real queries) as one compressed int32 fixture -- data, not source; the GPU box has no /root/reference."""
    out = {}
    for name, key in (("test_positive_triples.txt", "test"), ("triples-valid.txt", "valid")):
        arr = np.loadtxt(os.path.join(REF, "diffbot_data/FB15k", name), dtype=np.int64, delimiter="\t")
        assert arr.min() >= 0 and arr.max() < 2 ** 31
        out[key] = arr.astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "fb15k_test_valid_triples.npz"), **out)
    print("fb15k_test_valid_triples.npz:", {k: v.shape for k, v in out.items()})


def gen_train_step_golden():
    from oracle import hole_oracle as O
    from graphembeddings_b200 import data as D
    kg = D.synthetic_kg(n_relations=7, n_entities=300, n_triples=64, n_types=5, dim=20,
                        seed=11, trained_scale=True)
    E = kg.E.copy()
    pos = kg.triples[:64]
    off, ids = O.build_type_csr(kg.type_of)
    out = {}
    for step in (0, 1):
        side, neg = O.corrupt(pos, kg.type_of, off, ids, seed=5, step=step)
        E32 = E.copy()
        loss, vp, vn = O.sgd_step(E32, pos, neg, side, 0.2, 0.1, np.float32, order="tf")
        E64 = E.astype(np.float64)
        loss64, _, _ = O.sgd_step(E64, pos, neg, side, 0.2, 0.1, np.float64, order="tf")
        out.update({f"side{step}": side, f"neg{step}": neg, f"loss{step}": loss,
                    f"vp{step}": vp, f"vn{step}": vn, f"E32_{step}": E32, f"E64_{step}": E64,
                    f"loss64_{step}": loss64})
    np.savez_compressed(os.path.join(HERE, "train_step_oracle.npz"), E0=E, pos=pos,
                        type_of=kg.type_of, **out)
    print("train_step_oracle.npz written")


def gen_logloss_step_golden():
    """--log_loss step of the oracle (holE.py:194-196, 206-220): k = 2 corrupt batches drawn with
    virtual steps step*k + j, l2 = 1e-3, fp32 and fp64."""
    from oracle import hole_oracle as O
    from graphembeddings_b200 import data as D
    kg = D.synthetic_kg(n_relations=7, n_entities=300, n_triples=64, n_types=5, dim=20,
                        seed=13, trained_scale=True)
    pos = kg.triples[:64]
    off, ids = O.build_type_csr(kg.type_of)
    k, seed, step, lr, l2 = 2, 5, 3, 0.05, 1e-3
    drawn = [O.corrupt(pos, kg.type_of, off, ids, seed=seed, step=step * k + j) for j in range(k)]
    sides = [d[0] for d in drawn]
    negs = [d[1] for d in drawn]
    E32 = kg.E.copy()
    loss32, l2_32 = O.logloss_step(E32, pos, negs, sides, lr, l2, np.float32)
    E64 = kg.E.astype(np.float64)
    loss64, l2_64 = O.logloss_step(E64, pos, negs, sides, lr, l2, np.float64)
    np.savez_compressed(os.path.join(HERE, "logloss_step_oracle.npz"), E0=kg.E, pos=pos, type_of=kg.type_of,
                        k=k, seed=seed, step=step, lr=lr, l2=l2, sides=np.array(sides), negs=np.array(negs),
                        loss32=loss32, l2_32=l2_32, E32=E32, loss64=loss64, l2_64=l2_64, E64=E64)
    print("logloss_step_oracle.npz written")


if __name__ == "__main__":
    gen_ranking_golden()
    gen_fb15k_fixtures()
    gen_fb15k_full_triples()
    gen_train_step_golden()
    gen_logloss_step_golden()
