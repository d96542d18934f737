"""Same-box A/B of the host-buffer entry point (hole_train_steps_host) with different first-chunk sizes."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    import bench as B_
    from graphembeddings_b200.engine import HoleEngine
    B, K = 32768, 20
    kg, off, ids = B_.make_workload(6 * K * B)
    eng = HoleEngine(kg.n_rows, kg.dim).set_embeddings(kg.E).set_types(kg.type_of, off, ids)
    eng.set_relation_count(kg.n_relations)
    host = torch.from_numpy(kg.triples).pin_memory()
    lrs = B_.lr_schedule(K, 0, 915)
    eng.train_steps_host(host[: K * B], B, 1, 0, B_.MARGIN, lrs)
    best = 1e9
    for r in range(1, 6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.train_steps_host(host[r * K * B:(r + 1) * K * B], B, 1, r * K, B_.MARGIN, lrs)
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({"host_first": os.environ.get("HOLE_HOST_FIRST", "default"), "us_per_step": best / K * 1e6,
                      "Mtriples_s": K * B / best / 1e6}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        for hf in ("0", "2", "4", "8"):
            e = dict(os.environ, HOLE_HOST_FIRST=hf)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=e, cwd=ROOT, capture_output=True, text=True)
            sys.stdout.write(r.stdout if r.returncode == 0 else r.stderr[-500:])
