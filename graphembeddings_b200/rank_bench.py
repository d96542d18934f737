"""All-entity ranking throughput (BASELINE.json metric part 2: candidate scores/s), timed with
CUDA events on the launching stream.  Used by bench.py ("ranking" key) and tools/.

Roofline: tensor pipe.  2*D FLOP per candidate score, counted on the true D (not the padded
K); peak = measured dense bf16 GEMM throughput (MEASURED_PEAKS.json, burst figure for the
isolated kernel call).
"""
import json
import os

import numpy as np
import torch

from . import data as D
from .engine import HOLE_SIDE_BOTH, HOLE_SIDE_HEAD, HOLE_SIDE_TAIL, HoleEngine

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tensor_peak():
    path = os.path.join(_ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    return 1590.0, "fallback (B200_PROFILING.md)"


def _tensor_peak_sustained():
    """The pool's back-to-back bf16 GEMM figure (power-capped clocks): what a 50 ms tensor kernel can reach."""
    path = os.path.join(_ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            v = json.load(f).get("bf16_tflops_sustained")
            return float(v) if v else None
    return None


def time_rank(eng, queries, ent_begin, ent_end, sides=(HOLE_SIDE_TAIL, HOLE_SIDE_HEAD), reps=3,
              filters=None):
    """Best-of-reps milliseconds for ranking `queries` on every side in `sides`."""
    q = torch.as_tensor(queries, dtype=torch.int32).cuda()
    Q = q.shape[0] * (2 if HOLE_SIDE_BOTH in sides else 1)
    raw = torch.zeros(Q, dtype=torch.int32, device="cuda")
    filt = torch.zeros(Q, dtype=torch.int32, device="cuda")
    ts = torch.zeros(Q, dtype=torch.float32, device="cuda")
    best = float("inf")
    eng.rank_prepare(ent_begin, ent_end)     # the table is static during an evaluation: pack the candidates once
    for rep in range(reps + 1):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for side in sides:
            raw.zero_(); filt.zero_()
            fo, fi = (filters[side] if filters is not None else (None, None))
            eng.rank(q, side, ent_begin, ent_end, fo, fi, true_score=ts, raw_before=raw, filt_before=filt)
        ev1.record()
        torch.cuda.synchronize()
        if rep > 0:       # rep 0 is the warm-up (workspace allocation, attribute set)
            best = min(best, ev0.elapsed_time(ev1))
    return best, raw, filt


def _report(name, Q, n_sides, N, dim, ms, extra=None):
    peak, src = _tensor_peak()
    scores = float(Q) * n_sides * N
    tflops = scores * 2 * dim / (ms * 1e-3) / 1e12
    out = {"workload": name, "queries": int(Q), "sides": n_sides, "candidates": int(N), "dim": dim,
           "ms": ms, "scores_per_s": scores / (ms * 1e-3),
           "roofline": {"bound": "tensor", "achieved": tflops, "peak": peak, "unit": "TFLOP/s",
                        "frac": tflops / peak, "peak_source": src,
                        "frac_of_sustained_peak": (tflops / _tensor_peak_sustained()) if _tensor_peak_sustained() else None,
                        "kernel": "hole_rank_kernel (tcgen05 bf16, rank-count epilogue)"},
           "dtype": "bf16 operands, f32 accumulate"}
    if extra:
        out.update(extra)
    return out


def run(eng=None, kg=None, quick=False):
    """-> dict with the FB15k-shape (config 2) and 1.2M-entity (config 3) ranking numbers."""
    res = {}
    # config 2: FB15k shape, 59,071 test queries x {tail, head} x 14,951 candidates, d=150
    kg2 = D.make_config("rank_fb15k_d150", trained_scale=True)
    e2 = HoleEngine(kg2.n_rows, kg2.dim).set_embeddings(kg2.E)
    known = D.make_config("fb15k_d150", n_triples=100000, with_embeddings=False).triples
    fo_t, fi_t = D.build_filter_csr(kg2.triples, known, "tail")
    fo_h, fi_h = D.build_filter_csr(kg2.triples, known, "head")
    # one pass over both sides: filter CSR rows [0,Q) = tail queries, [Q,2Q) = head queries
    fo = np.concatenate([fo_t, fo_h[1:] + fo_t[-1]])
    fi = np.concatenate([fi_t, fi_h])
    filters = {HOLE_SIDE_BOTH: (torch.as_tensor(fo).cuda(), torch.as_tensor(fi).cuda())}
    ms, raw, filt = time_rank(e2, kg2.triples, kg2.n_relations, kg2.n_rows, sides=(HOLE_SIDE_BOTH,),
                              filters=filters)
    r = (filt.cpu().numpy()[len(kg2.triples):] + 1).astype(np.float64)
    res["fb15k_shape"] = _report("rank_fb15k_d150: 59,071 queries x 2 sides x 14,951 candidates (filtered, one pass)",
                                 len(kg2.triples), 2, kg2.n_entities, kg2.dim, ms,
                                 {"filter_entries": int(len(fi)),
                                  "head_side_filtered_mrr_random_table": float(np.mean(1.0 / r))})
    e2.close()
    # config 3: 100k queries x 1.2M candidates, d=256 (the training table of config 1)
    if eng is not None and kg is not None:
        nq = 20000 if quick else 100000
        rng = np.random.default_rng(20170906)
        q = kg.triples[rng.integers(0, len(kg.triples), size=nq)]
        ms, _, _ = time_rank(eng, q, kg.n_relations, kg.n_rows, sides=(HOLE_SIDE_TAIL,), reps=2)
        res["diffbot_shape"] = _report(f"rank_diffbot_d256: {nq} queries x 1,200,000 candidates (tail side)",
                                       nq, 1, kg.n_entities, kg.dim, ms)
    return res
