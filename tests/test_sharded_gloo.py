"""The multi-GPU protocol (row-sharded table, all-to-all of rows and row deltas) on CPU:
world_size 2 and 3 over gloo, with the oracle as the local step, against the single-process
oracle on the concatenated global batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphembeddings_b200 import data as D
from graphembeddings_b200.sharded import RowShardedTrainer, row_partition
from oracle import hole_oracle as O


class OracleBackend:
    """Local compute = the NumPy oracle (fp64); rows are unpadded."""

    def __init__(self, n_relations, dim, max_batch, type_of, csr_off, csr_ids):
        self.R, self.dim, self.width = n_relations, dim, dim
        self.W = torch.zeros((n_relations + 3 * max_batch, dim), dtype=torch.float64)
        self.type_of, self.csr_off, self.csr_ids = type_of, csr_off, csr_ids

    def pad_rows(self, E):
        return torch.as_tensor(E, dtype=torch.float64).clone()

    def unpad_rows(self, P):
        return P

    def corrupt(self, triples, seed, step, index_base):
        tr = triples.numpy()
        from oracle import philox
        side = philox.side_coin(seed, step)
        ent = tr[:, 0 if side else 1].astype(np.int64)
        ty = self.type_of[ent].astype(np.int64)
        lo = self.csr_off[ty]
        cnt = self.csr_off[ty + 1] - lo
        j = philox.entity_draw(seed, step, index_base + np.arange(len(tr)), cnt)
        return side, torch.from_numpy(self.csr_ids[lo + j].astype(np.int64))

    def step(self, n_rows, pos, neg, side, margin, lr):
        Wn = self.W.numpy()
        loss, _, _ = O.sgd_step(Wn, pos.numpy(), neg.numpy(), side, margin, lr, np.float64, "tf")
        return torch.from_numpy(loss)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Bl, steps = 40, 3
        kg = D.synthetic_kg(5, 97, Bl * world * steps, 4, 12, seed=31, trained_scale=True, zipf_entities=True)
        off, ids = O.build_type_csr(kg.type_of)
        be = OracleBackend(kg.n_relations, kg.dim, Bl, kg.type_of, off, ids)
        tr = RowShardedTrainer(kg.n_relations, kg.n_entities, kg.dim, be, dist).load_embeddings(kg.E)
        losses = []
        for s in range(steps):
            gb = kg.triples[s * Bl * world:(s + 1) * Bl * world]
            losses.append(tr.train_step(torch.from_numpy(gb[rank * Bl:(rank + 1) * Bl]), 7, s, 0.2, 0.1).numpy())
        full = tr.gather_embeddings().numpy()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), E=full, loss=np.concatenate(losses))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_training_matches_single_process_oracle(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    Bl, steps = 40, 3
    kg = D.synthetic_kg(5, 97, Bl * world * steps, 4, 12, seed=31, trained_scale=True, zipf_entities=True)
    off, ids = O.build_type_csr(kg.type_of)
    E = kg.E.astype(np.float64)
    want_loss = []
    for s in range(steps):
        gb = kg.triples[s * Bl * world:(s + 1) * Bl * world]
        side, neg = O.corrupt(gb, kg.type_of, off, ids, 7, s)
        l, _, _ = O.sgd_step(E, gb, neg, side, 0.2, 0.1, np.float64, "tf")
        want_loss.append(l)
    outs = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    for r in range(world):
        assert np.abs(outs[r]["E"] - E).max() < 1e-12          # every rank gathers the same table
        got = outs[r]["loss"].reshape(steps, Bl)
        for s in range(steps):
            assert np.abs(got[s] - want_loss[s][r * Bl:(r + 1) * Bl]).max() < 1e-12
    assert np.abs(E - kg.E).max() > 1e-3


def test_row_partition_covers_all_rows():
    for n, w in ((97, 2), (97, 3), (100, 8), (7, 8)):
        per = row_partition(n, w)
        assert per * w >= n and per * (w - 1) < n or n <= w
