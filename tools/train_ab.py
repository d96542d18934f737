"""Same-box A/B of training-step variants on BASELINE config 1 (d=256, 1,200,014 rows, B=32768).
Each variant runs in its own process (library path / environment differ); prints one line each.

    python tools/train_ab.py                # parent: runs every variant in VARIANTS
    python tools/train_ab.py --child NAME   # child: measures with the current environment
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "graphembeddings_b200")

PREV = os.path.join(PKG, "libhole_b200_prev.so")   # the previous commit's sources built beside the current library
VARIANTS = [  # (name, env); earlier rounds of variants: profiles/r02_train_ab_README.md
    ("default", {}),
    ("trained-scale table (clips fire)", {"AB_TRAINED": "1"}),
    ("B=512", {"AB_BATCH": "512"}),
    ("B=8192", {"AB_BATCH": "8192"}),
    ("radix chain (HOLE_SORT_SMALL=0)", {"HOLE_SORT_SMALL": "0"}),
    ("one-launch sort from the first call (HOLE_SORT_SMALL=2)", {"HOLE_SORT_SMALL": "2"}),
    ("default (again)", {}),
]


def child(name):
    import numpy as np
    import torch
    import bench as B_
    from graphembeddings_b200.engine import HoleEngine
    B = int(os.environ.get("AB_BATCH", 32768))
    trained = os.environ.get("AB_TRAINED") == "1"
    from graphembeddings_b200 import data as D
    kg = D.make_config(B_.WORKLOAD, n_triples=230 * B, trained_scale=trained)
    off, ids = D.build_type_csr(kg.type_of)
    eng = HoleEngine(kg.n_rows, kg.dim).set_embeddings(kg.E).set_types(kg.type_of, off, ids)
    eng.set_relation_count(kg.n_relations)
    tri = torch.from_numpy(kg.triples).cuda()

    def timed(k0, n):
        lrs = B_.lr_schedule(n, k0, 30_000_000 // B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.train_steps(tri[k0 * B:(k0 + n) * B], B, 1, k0, B_.MARGIN, lrs)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / n     # us per step

    timed(0, 5)
    t20_runs = [timed(5 + 20 * r, 20) for r in range(4)]     # (the first call after the 5-step warm-up is bench.py's timed call)
    t20 = min(t20_runs)
    t200 = min(timed(10, 200) for r in range(2))
    eng.profile(True)
    timed(10, 50)
    k1, k3, n = eng.profile_read()
    eng.profile(False)
    alg = (32 * kg.dim + 20) * B
    print(json.dumps({"variant": name, "us_per_step_20_runs": [round(t, 2) for t in t20_runs], "us_per_step_20": round(t20, 2), "us_per_step_200": round(t200, 2),
                      "k1_us": round(k1 * 1e3 / n, 2), "k3_us": round(k3 * 1e3 / n, 2),
                      "Mtriples_s_20": round(B / t20, 1), "Mtriples_s_200": round(B / t200, 1),
                      "frac20": round(alg / (t20 * 1e-6) / 1e9 / 6548.5, 3),
                      "frac_k": round(alg / ((k1 + k3) / n * 1e-3) / 1e9 / 6548.5, 3)}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        only = os.environ.get("AB_ONLY")
        for name, env in VARIANTS:
            if only and only not in name:
                continue
            if "HOLE_B200_LIB" in env and not os.path.exists(env["HOLE_B200_LIB"]):
                print(json.dumps({"variant": name, "skipped": "library not built"}), flush=True)
                continue
            e = dict(os.environ)
            e.update(env)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name], env=e, cwd=ROOT,
                               capture_output=True, text=True)
            sys.stdout.write(r.stdout if r.returncode == 0 else json.dumps(
                {"variant": name, "error": r.stderr[-600:]}) + "\n")
            sys.stdout.flush()
