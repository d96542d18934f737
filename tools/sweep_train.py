"""Quick device-resident training throughput sweep (not the bench contract; exploration)."""
import argparse
import json
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from graphembeddings_b200 import data as D
from graphembeddings_b200.engine import HoleEngine


def run(name, n_triples, batches, steps, trained=False):
    t0 = time.time()
    kg = D.make_config(name, n_triples=n_triples, trained_scale=trained)
    off, ids = D.build_type_csr(kg.type_of)
    e = HoleEngine(kg.n_rows, kg.dim).set_embeddings(kg.E).set_types(kg.type_of, off, ids)
    tri = torch.from_numpy(kg.triples).cuda()
    print(f"# {name}: N={kg.n_rows} D={kg.dim} gen {time.time()-t0:.1f}s", flush=True)
    for B in batches:
        n_steps = min(steps, n_triples // B)
        if n_steps < 1:
            continue
        lrs = [0.1] * n_steps
        t = tri[: n_steps * B]
        for _ in range(2):
            e.train_steps(t, B, 1, 0, 0.2, lrs)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for rep in range(3):
            ev0.record()
            e.train_steps(t, B, 1, rep * n_steps, 0.2, lrs)
            ev1.record()
            torch.cuda.synchronize()
            best = min(best, ev0.elapsed_time(ev1))
        e.profile(True)
        e.train_steps(t, B, 1, 7 * n_steps, 0.2, lrs)
        k1, k3, npf = e.profile_read()
        e.profile(False)
        tps = n_steps * B / (best * 1e-3)
        gbs = tps * (32 * kg.dim + 20) / 1e9
        print(json.dumps({"cfg": name, "B": B, "steps": n_steps, "ms_per_step": best / n_steps,
                          "k1_us": k1 * 1e3 / npf, "k3_us": k3 * 1e3 / npf, "Mtriples_s": tps / 1e6, "alg_GBs": gbs, "frac_6548": gbs / 6548.5}), flush=True)
    e.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="both")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--batches", type=str, default="")
    a = ap.parse_args()
    bl = [int(x) for x in a.batches.split(",")] if a.batches else None
    if a.cfg in ("both", "fb15k"):
        run("fb15k_d150", 483142, bl or [512, 2048, 8192, 32768], a.steps)
    if a.cfg in ("both", "diffbot"):
        run("diffbot_d256", 4_000_000, bl or [512, 2048, 4096, 8192, 16384, 32768, 131072], a.steps)
