"""Multi-GPU HolE: entity table row-sharded over the ranks of one node.

One process per GPU (torch.distributed for rendezvous; gloo in the CPU tests).

Training is batch-synchronous and gives the same result as one GPU on the concatenated batch up to
fp32 summation order: every rank corrupts its slice of the global batch (Philox keyed on the
GLOBAL triple index, so the draw does not depend on the number of ranks), the rows a rank's
triples touch come from their owners, the rank runs the step kernels on them, and every row's
change travels back to its owner, who adds the changes in rank order (deterministic).  Relation
rows are replicated and take every rank's change in the same rank order.

Two trainers implement this:
  * P2PRowShardedTrainer (GPUs with peer access; the default) -- one library call per step
    (hole_shard_step): the training kernel gathers rows straight from the owners' shards over
    NVLink and stores the row deltas into the owners' staging buffers; no NCCL, no host
    synchronisation inside a step (include/hole_b200.h, csrc/hole_shard.cuh);
  * RowShardedTrainer -- generic all-to-all version (NCCL or gloo) with a pluggable local
    backend, which is what the CPU tests drive against the oracle.
Ranking: candidates are sharded by row block; the true candidate's score comes from its owner
(all-reduce of a zero-initialised vector), the int32 counts are all-reduced.
"""
import json
import os
import sys
import time

import numpy as np

import torch


TIMING = None      # set to a dict to collect a per-section breakdown (tools/sharded_breakdown.py)
EVENTS = None      # set to a list to collect (name, start, end) CUDA events without synchronising


class _Section:
    """Section timer: synchronising wall-clock when sharded.TIMING is a dict, stream events
    (no synchronisation, the GPU timeline as it runs) when sharded.EVENTS is a list."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if TIMING is not None:
            torch.cuda.synchronize()
            self.t0 = time.perf_counter()
        elif EVENTS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if TIMING is not None:
            torch.cuda.synchronize()
            TIMING[self.name] = TIMING.get(self.name, 0.0) + (time.perf_counter() - self.t0) * 1e3
        elif EVENTS is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            EVENTS.append((self.name, self.e0, e1))


def row_partition(n_entities, world):
    """Contiguous blocks of entity rows; returns rows_per_rank."""
    return (n_entities + world - 1) // world


class CudaBackend:
    """libhole_b200 on this rank's GPU.  The engine's table is the step table W."""

    def __init__(self, n_relations, dim, max_batch, device_index, type_of=None, csr_off=None, csr_ids=None):
        from .engine import HoleEngine
        self.R = n_relations
        self.eng = HoleEngine(n_relations + 3 * max_batch, dim, device_index)
        self.eng.set_relation_count(n_relations)
        self.width = self.eng.row_stride
        self.dim = dim
        self.device = self.eng.device
        self.max_batch = int(max_batch)
        self._W = self._D = None
        if type_of is not None:
            self.eng.set_types(type_of, csr_off, csr_ids)

    @property
    def W(self):
        """Step table [relations | fetched rows] of the generic trainer (allocated on first use)."""
        if self._W is None:
            self._W = torch.zeros((self.R + 3 * self.max_batch, self.width), dtype=torch.float32, device=self.device)
            self.eng.table = self._W
        return self._W

    @property
    def D(self):
        """Row deltas of one step (delta-mode kernels) of the generic trainer."""
        if self._D is None:
            self._D = torch.zeros_like(self.W)
        return self._D

    def pad_rows(self, E):
        """checkpoint layout [n, dim] -> device layout [n, width]"""
        E = torch.as_tensor(E, dtype=torch.float32, device=self.device)
        H, Hp = self.dim // 2, self.width // 2
        out = torch.zeros((E.shape[0], self.width), dtype=torch.float32, device=self.device)
        out[:, :H] = E[:, :H]
        out[:, Hp:Hp + H] = E[:, H:]
        return out

    def unpad_rows(self, P):
        H, Hp = self.dim // 2, self.width // 2
        return torch.cat([P[:, :H], P[:, Hp:Hp + H]], dim=1)

    def corrupt(self, triples, seed, step, index_base):
        side, neg = self.eng.corrupt_batch(triples, seed, step, index_base)
        return side, neg.long()

    def step(self, n_rows, pos, neg, side, margin, lr):
        self.W                                   # the engine's table is the step table
        return self.eng.train_step(pos.to(torch.int32), neg.to(torch.int32), side, margin, lr)

    # fast path: plan on a side stream, deltas written by the kernels, owner-side row adds
    def plan(self, pos_i32, neg_i32):
        self.eng.train_step_plan(pos_i32, neg_i32)

    def step_delta(self, n_rows, pos_i32, neg_i32, side, margin, lr):
        self.W
        self.D[:n_rows].zero_()
        return self.eng.train_step_delta(pos_i32, neg_i32, side, margin, lr, self.D)

    def add_rows(self, table, ids, id_offset, rows):
        self.eng.add_rows(table, ids.contiguous(), id_offset, rows.contiguous())


class RowShardedTrainer:
    def __init__(self, n_relations, n_entities, dim, backend, dist=None, max_batch=None):
        self.R, self.n_ent, self.dim = int(n_relations), int(n_entities), int(dim)
        self.dist = dist
        self.my_rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.rows_per = row_partition(self.n_ent, self.world)
        self.begin = self.R + self.my_rank * self.rows_per                   # first global row I own
        self.end = min(self.R + self.n_ent, self.begin + self.rows_per)
        self.be = backend
        self.shard = None          # [R + n_mine, width]: replicated relations, then my block

    # ------------------------------------------------------------------ table
    def load_embeddings(self, E):
        """E: full [N, dim] table (checkpoint layout) available on every rank; keeps my part."""
        E = torch.as_tensor(E)
        mine = torch.cat([E[: self.R], E[self.begin:self.end]], dim=0)
        self.shard = self.be.pad_rows(mine)
        return self

    def gather_embeddings(self):
        """Full [N, dim] table on every rank (tests / checkpointing)."""
        mine = self.be.unpad_rows(self.shard[self.R:]).contiguous()
        if self.world == 1:
            ents = mine
        else:
            pad = torch.zeros((self.rows_per - mine.shape[0], mine.shape[1]), dtype=mine.dtype, device=mine.device)
            buf = torch.cat([mine, pad], dim=0)
            out = [torch.empty_like(buf) for _ in range(self.world)]
            self.dist.all_gather(out, buf)
            ents = torch.cat(out, dim=0)[: self.n_ent]
        return torch.cat([self.be.unpad_rows(self.shard[: self.R]), ents], dim=0)

    # ------------------------------------------------------------------ exchange helpers
    def _a2a(self, send, send_counts, recv_counts, out=None):
        if self.world == 1:
            if out is not None:
                out.copy_(send)
                return out
            return send
        shape = (int(sum(recv_counts)),) + tuple(send.shape[1:])
        recv = out if out is not None else torch.empty(shape, dtype=send.dtype, device=send.device)
        self.dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=list(recv_counts),
                                    input_split_sizes=list(send_counts))
        return recv

    def _route(self, uniq):
        """uniq: sorted unique global entity rows I need.  -> (send_counts, recv_counts) lists."""
        dev = uniq.device
        bounds = self.R + self.rows_per * torch.arange(1, self.world + 1, device=dev)
        cut = torch.searchsorted(uniq, bounds)
        send = torch.diff(cut, prepend=torch.zeros(1, dtype=cut.dtype, device=dev))
        if self.world == 1:
            sc = [int(send[0])]
            return sc, sc
        recv = torch.empty_like(send)
        self.dist.all_to_all_single(recv, send)
        both = torch.stack([send, recv]).tolist()          # one host synchronisation
        return [int(x) for x in both[0]], [int(x) for x in both[1]]

    # ------------------------------------------------------------------ training
    def train_step(self, pos_local, seed, step, margin, lr, next_pos=None):
        """pos_local: this rank's [B,3] slice (global row ids) of the global batch of
        world*B triples.  Returns the per-triple hinge loss of the slice.  (next_pos is used by
        the peer-memory trainer to work one step ahead; ignored here.)"""
        dev = self.shard.device
        pos = torch.as_tensor(pos_local).to(dev).long()
        B, R = pos.shape[0], self.R
        with _Section("corrupt"):
            side, neg = self.be.corrupt(pos.to(torch.int32), seed, step, self.my_rank * B)
            neg = neg.to(dev)
        with _Section("unique"):
            ents = torch.cat([pos[:, 0], pos[:, 1], neg]).to(torch.int32)    # 32-bit keys sort faster
            uniq, inv = torch.unique(ents, return_inverse=True)
            uniq = uniq.long()
            U = uniq.shape[0]
        with _Section("route (counts, host sync)"):
            send_counts, recv_counts = self._route(uniq)
        with _Section("a2a ids"):
            ids_in = self._a2a(uniq, send_counts, recv_counts)              # rows others want from me
        fast = hasattr(self.be, "step_delta")
        if fast:
            with _Section("plan (side stream)"):
                pos_w = torch.stack([R + inv[:B], R + inv[B:2 * B], pos[:, 2]], dim=1).to(torch.int32)
                neg_w = (R + inv[2 * B:]).to(torch.int32)
                self.be.plan(pos_w, neg_w)        # overlaps the row exchange below
        with _Section("gather rows"):
            rows_out = self.shard.index_select(0, ids_in - self.begin + R)
        W = self.be.W
        if fast:
            with _Section("a2a rows"):
                self._a2a(rows_out, recv_counts, send_counts, out=W[R:R + U])   # lands in the step table
                W[:R].copy_(self.shard[:R])
            with _Section("local step (K1+K3, delta mode)"):
                loss = self.be.step_delta(R + U, pos_w, neg_w, side, margin, lr)
                delta = self.be.D
        else:
            with _Section("a2a rows"):
                rows_in = self._a2a(rows_out, recv_counts, send_counts)         # in `uniq` order
            with _Section("assemble W, W0"):
                W[:R].copy_(self.shard[:R])
                W[R:R + U].copy_(rows_in)
                W0 = W[: R + U].clone()
                pos_w = torch.stack([R + inv[:B], R + inv[B:2 * B], pos[:, 2]], dim=1)
                neg_w = R + inv[2 * B:]
            with _Section("local step (plan+K1+K3)"):
                loss = self.be.step(R + U, pos_w, neg_w, side, margin, lr)
            with _Section("delta"):
                delta = W[: R + U] - W0
        with _Section("allreduce relations"):
            d_rel = delta[:R].contiguous()
            if self.world > 1:
                self.dist.all_reduce(d_rel)
            self.shard[:R] += d_rel
        with _Section("a2a deltas"):
            d_in = self._a2a(delta[R:R + U], send_counts, recv_counts)      # grouped by source rank
        with _Section("apply deltas"):
            off = 0
            for k in range(self.world):                                     # rank order: deterministic
                n_k = recv_counts[k]
                if n_k:
                    if fast:
                        self.be.add_rows(self.shard, ids_in[off:off + n_k], R - self.begin, d_in[off:off + n_k])
                    else:
                        self.shard.index_add_(0, ids_in[off:off + n_k] - self.begin + R, d_in[off:off + n_k])
                off += n_k
        return loss

    # ------------------------------------------------------------------ ranking
    def rank(self, queries, side, filter_off=None, filter_ids=None):
        """All-entity ranking of (replicated) queries [Q,3] with candidates sharded by row
        block.  Returns (raw_before, filt_before) int32[Q], identical on every rank."""
        from .engine import HOLE_SIDE_TAIL, HoleEngine
        dev = self.shard.device
        q = torch.as_tensor(queries).to(dev).long()
        Q, R = q.shape[0], self.R
        other_col, true_col = (0, 1) if side == HOLE_SIDE_TAIL else (1, 0)
        uniq, inv = torch.unique(q[:, other_col], return_inverse=True)
        send_counts, recv_counts = self._route(uniq)
        ids_in = self._a2a(uniq, send_counts, recv_counts)
        rows_in = self._a2a(self.shard.index_select(0, ids_in - self.begin + R), recv_counts, send_counts)
        n_mine = self.end - self.begin                                  # (the last shard may carry padding rows)
        table = torch.cat([self.shard[:R + n_mine], rows_in], dim=0)    # [R | my block | fetched]
        tr = q[:, true_col]
        # true candidate: local index if mine, else below / above my candidate range
        tr_loc = torch.where(tr < self.begin, torch.zeros_like(tr) - 1 + R,       # < ent_begin
                             torch.where(tr >= self.end, torch.full_like(tr, R + n_mine + 1),
                                         tr - self.begin + R))
        ql = torch.empty_like(q)
        ql[:, other_col] = R + n_mine + inv
        ql[:, true_col] = tr_loc
        ql[:, 2] = q[:, 2]
        fo = fi = None
        if filter_off is not None:
            fo = torch.as_tensor(filter_off).to(dev)
            fi = (torch.as_tensor(filter_ids).to(dev).long() - self.begin + R).to(torch.int32)
        # a ranking context sized for [relations | my block | fetched rows]
        eng = getattr(self, "_rank_eng", None)
        if eng is None or eng.n_rows < table.shape[0]:
            if eng is not None:
                eng.close()
            eng = self._rank_eng = HoleEngine(int(table.shape[0] * 1.25) + 1024, self.dim, dev.index or 0)
        eng.table = table
        try:
            ts = torch.zeros(Q, dtype=torch.float32, device=dev)
            scratch = torch.zeros(Q, dtype=torch.int32, device=dev)
            eng.rank(ql.to(torch.int32), side, R, R + n_mine, None, None, true_score=ts,
                     compute_true=True, raw_before=scratch, filt_before=scratch.clone())
            if self.world > 1:
                self.dist.all_reduce(ts)          # exactly one owner wrote each entry, others hold 0
            raw = torch.zeros(Q, dtype=torch.int32, device=dev)
            filt = torch.zeros(Q, dtype=torch.int32, device=dev)
            eng.rank(ql.to(torch.int32), side, R, R + n_mine, fo, fi, true_score=ts,
                     compute_true=False, raw_before=raw, filt_before=filt)
            if self.world > 1:
                self.dist.all_reduce(raw)
                self.dist.all_reduce(filt)
        finally:
            eng.table = None
        return raw, filt


class PeerMemoryUnavailable(RuntimeError):
    """The GPUs of this job cannot map each other's memory (no P2P / CUDA IPC)."""


def make_trainer(n_relations, n_entities, dim, backend, dist, log=None):
    """P2PRowShardedTrainer when every rank can map every other rank's buffers, else (all ranks
    together) the NCCL all-to-all RowShardedTrainer.  HOLE_SHARDED_NCCL=1 forces the latter."""
    if os.environ.get("HOLE_SHARDED_NCCL") != "1" and hasattr(backend, "eng"):
        try:
            return P2PRowShardedTrainer(n_relations, n_entities, dim, backend, dist)
        except PeerMemoryUnavailable as e:
            if log:
                log(f"peer memory unavailable, using NCCL all-to-alls: {e}")
    return RowShardedTrainer(n_relations, n_entities, dim, backend, dist)


class P2PRowShardedTrainer(RowShardedTrainer):
    """Row-sharded training with the exchange fused into the training kernel over NVLink peer
    memory: one library call per step (hole_shard_step, csrc/hole_shard.cuh).  Every rank maps
    (CUDA IPC) every other rank's shard, delta staging, relation staging, request inbox and flags.

      side stream, one step ahead (pass `next_pos`):  Philox corruption -> request routing -> update plan
      compute stream:  post request lists -> K1 gathers rows from the owners' shards and stores row
                       deltas into the owners' staging buffers -> K3 -> relation deltas + "delivered"
                       flag -> owner adds the staged deltas in rank order + "current" flag

    No NCCL call and no host synchronisation inside a step."""

    def __init__(self, n_relations, n_entities, dim, backend, dist, timeout_s=0.0, check_every=256):
        super().__init__(n_relations, n_entities, dim, backend, dist)
        from torch.multiprocessing.reductions import reduce_tensor
        dev, G, R = backend.device, self.world, self.R
        eng = backend.eng
        if eng.type_of is None:
            raise ValueError("the backend needs the type tables (type_of, csr_off, csr_ids) for the corruption")
        self.max_batch = backend.max_batch
        self.cap = 3 * self.max_batch
        width = backend.width
        self.shard = torch.zeros((R + self.rows_per, width), dtype=torch.float32, device=dev)
        self.stage = torch.zeros((G, self.cap, width), dtype=torch.float32, device=dev)
        self.relstage = torch.zeros((G, max(R, 1), width), dtype=torch.float32, device=dev)
        self.inbox = torch.zeros((2, G, self.cap), dtype=torch.int32, device=dev)
        self.meta = torch.zeros((2, G, 2), dtype=torch.int32, device=dev)
        self.flags = torch.zeros((2, G), dtype=torch.int32, device=dev)
        self.barrier_err = torch.zeros(1, dtype=torch.int32, device=dev)
        mine = [self.shard, self.stage, self.relstage, self.inbox, self.meta, self.flags]
        if G == 1:
            peers = [mine]
        else:
            everyone = [None] * G
            dist.all_gather_object(everyone, ([reduce_tensor(t) for t in mine], dev.index))
            peers, failure = [], None
            for k, (handles, dev_k) in enumerate(everyone):
                if k == self.my_rank:
                    peers.append(mine)
                    continue
                # Open the peer's allocation in MY device's address space (argument 6 of torch's
                # rebuild is the device the IPC handle is opened on): the tensor then reads as
                # local to torch while its pages live on GPU k -- kernels launched on my device
                # reach it over NVLink.
                try:
                    eng.enable_peer_access(dev_k)
                    peers.append([fn(*(list(a[:6]) + [dev.index] + list(a[7:]))) for fn, a in handles])
                except Exception as e:                       # no P2P / IPC between these two GPUs
                    failure = f"rank {self.my_rank} cannot map rank {k}'s buffers: {e}"
                    break
            # every rank must reach the same verdict, or the first flag wait would spin
            verdicts = [None] * G
            dist.all_gather_object(verdicts, failure)
            bad = [v for v in verdicts if v]
            if bad:
                raise PeerMemoryUnavailable("; ".join(bad))
        self._bind(peers, timeout_s)
        self.check_every = int(check_every)
        self._since_check = 0
        if G > 1:
            torch.cuda.synchronize()
            dist.barrier()

    def _bind(self, peers, timeout_s=0.0):
        self._peers = peers                                         # keeps the mappings alive
        pa = self.be.eng.peer_array
        self.be.eng.shard_init(self.world, self.my_rank, self.R, self.n_ent, self.rows_per, self.max_batch,
                               self.shard, *[pa([p[i] for p in peers]) for i in range(6)], self.barrier_err,
                               timeout_s)

    def load_embeddings(self, E):
        E = torch.as_tensor(E)
        mine = torch.cat([E[: self.R], E[self.begin:self.end]], dim=0)
        self.shard.zero_()
        self.shard[: mine.shape[0]].copy_(self.be.pad_rows(mine))
        return self

    def load_shard(self, rows):
        """rows: device [R + n_mine, row_stride] already in the padded layout (bench: generated on the GPU)."""
        self.shard.zero_()
        self.shard[: rows.shape[0]].copy_(rows)
        return self

    def check_barriers(self):
        """Raises if a peer ever failed to reach a step barrier (host synchronisation)."""
        self._since_check = 0
        if self.be.eng.shard_poll():
            raise RuntimeError("a peer did not reach a step barrier in time; the sharded tables are inconsistent")

    def gather_embeddings(self):
        self.check_barriers()
        return super().gather_embeddings()

    def rank(self, *a, **kw):
        self.check_barriers()
        return super().rank(*a, **kw)

    def _as_slice(self, pos):
        pos = torch.as_tensor(pos)
        if pos.device != self.shard.device or pos.dtype != torch.int32 or not pos.is_contiguous():
            pos = pos.to(self.shard.device, non_blocking=True).to(torch.int32).contiguous()
        return pos

    def train_step(self, pos_local, seed, step, margin, lr, next_pos=None):
        """pos_local: this rank's [B,3] slice of the global batch (device int32 tensors are used as they
        are and must stay alive until the step has run).  next_pos: the slice of step + 1, if known --
        its corruption, routing and plan are then built on the side stream while this step runs."""
        eng = self.be.eng
        pos = self._as_slice(pos_local)
        keep = [pos]
        if getattr(self, "_prepared_key", None) != (pos.data_ptr(), pos.shape[0], int(seed), int(step)):
            eng.shard_prepare(pos, seed, step)              # nothing was built ahead for this step
        self._prepared_key = None
        if next_pos is not None:                            # enqueue before the step, so that it overlaps it
            nxt = self._as_slice(next_pos)
            keep.append(nxt)
            eng.shard_prepare(nxt, seed, step + 1)
            self._prepared_key = (nxt.data_ptr(), nxt.shape[0], int(seed), int(step) + 1)
        loss = eng.shard_step(pos, seed, step, margin, lr)
        self._keep = keep                                   # the side stream may still read them
        self._since_check += 1
        if self.check_every and self._since_check >= self.check_every:
            self.check_barriers()
        return loss

    def train_steps(self, triples_i32, batch_size, seed, first_step, margin, lrs):
        """n_steps consecutive steps on device-resident slices [n_steps*B, 3] without returning to the
        host (hole_shard_steps).  Returns this rank's per-step loss sums (device)."""
        sums = self.be.eng.shard_steps(triples_i32, batch_size, seed, first_step, margin, lrs)
        self._since_check += triples_i32.shape[0] // batch_size
        return sums

    def train_steps_host(self, triples_host, batch_size, seed, first_step, margin, lrs):
        """Same from pinned host triples; returns numpy loss sums (blocks)."""
        out = self.be.eng.shard_steps_host(triples_host, batch_size, seed, first_step, margin, lrs)
        self.check_barriers()
        return out


# --------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1)
# --------------------------------------------------------------------------------------
def bench(args, dist, rank, world, local_rank):
    """Weak scaling: every rank trains `--batch` triples per step on the row-sharded
    BASELINE config 1 table.  Device time, max over ranks; rank 0 prints the JSON line."""
    import bench as B_
    from . import data as D
    from .engine import HOLE_SIDE_TAIL

    Bl, K, W = args.batch, args.steps, args.warmup
    kg = D.make_config(B_.WORKLOAD, n_triples=(K + W) * Bl * world)
    off, ids = D.build_type_csr(kg.type_of)
    be = CudaBackend(kg.n_relations, kg.dim, Bl, local_rank, kg.type_of, off, ids)
    tr = make_trainer(kg.n_relations, kg.n_entities, kg.dim, be, dist,
                      log=lambda m: print(m, file=sys.stderr) if rank == 0 else None).load_embeddings(kg.E)
    cls = type(tr)
    # step s, rank r takes triples [(s*world + r)*Bl, +Bl)
    mine = torch.from_numpy(kg.triples).view(K + W, world, Bl, 3)[:, rank].contiguous()
    dev_tri = mine.cuda()
    host_tri = mine.pin_memory()
    lrs = B_.lr_schedule(3 * (K + W), 0, 30_000_000 // (Bl * world))

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h0 = time.perf_counter()
        fn()
        host_ms[0] = (time.perf_counter() - h0) * 1e3      # host time to enqueue the region
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    host_ms = [0.0]

    # nvidia-smi's start-up (NVML initialisation) disturbs running GPU work for a few hundred
    # ms: start the clock sampler before the warm-up, not inside the timed region
    sampler = B_.ClockSampler(local_rank, period=0.25)
    if rank == 0 and os.environ.get("HOLE_NO_SAMPLER") != "1":
        sampler.start()
        time.sleep(1.0)
    for s in range(W):
        tr.train_step(dev_tri[s], 1, s, B_.MARGIN, float(lrs[s]), next_pos=dev_tri[s + 1])
    be.eng.reset_launch_count()
    stamps = []

    def value_pass():
        for s in range(K):
            stamps.append(time.perf_counter())
            tr.train_step(dev_tri[W + s], 1, W + s, B_.MARGIN, float(lrs[W + s]),
                          next_pos=dev_tri[W + s + 1] if s + 1 < K else None)
        stamps.append(time.perf_counter())

    ms = timed(value_pass)
    if os.environ.get("HOLE_BENCH_TRACE") == "1":
        gaps = np.diff(np.asarray(stamps)) * 1e6
        big = np.flatnonzero(gaps > 1000)
        print(f"[trace rank {rank}] host us/step: median {np.median(gaps):.0f}, mean {gaps.mean():.0f}, "
              f"max {gaps.max():.0f}; steps > 1 ms: {[(int(i), int(gaps[i])) for i in big[:12]]}", flush=True)
    launches = be.eng.launch_count()
    host_enqueue_us = host_ms[0] / K * 1e3
    if hasattr(tr, "check_barriers"):
        tr.check_barriers()
    value = K * Bl * world / (ms * 1e-3)

    losses = []
    stage = [torch.empty_like(dev_tri[0]) for _ in range(2)]

    def e2e_pass():
        # two device staging slots, filled from pinned host memory one step ahead
        stage[0].copy_(host_tri[W], non_blocking=True)
        for s in range(K):
            nxt = None
            if s + 1 < K:
                nxt = stage[(s + 1) & 1]
                nxt.copy_(host_tri[W + s + 1], non_blocking=True)
            loss = tr.train_step(stage[s & 1], 1, K + W + s, B_.MARGIN, float(lrs[K + W + s]), next_pos=nxt)
            losses.append(float(loss.sum().item()))        # device -> host read of the step's result

    ms_e2e = timed(e2e_pass)
    e2e = K * Bl * world / (ms_e2e * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    # sharded ranking: 20k replicated queries x 1.2M candidates split over the ranks
    rng = np.random.default_rng(20170906)
    q = torch.from_numpy(kg.triples[rng.integers(0, len(kg.triples), size=20000)])
    tr.rank(q, HOLE_SIDE_TAIL)
    ms_rank = timed(lambda: tr.rank(q, HOLE_SIDE_TAIL))
    if rank == 0:
        peak, src, _ = B_.peaks()
        alg = (32 * kg.dim + 20)
        print(json.dumps({
            "metric": "HolE train triples/s", "value": value, "unit": "triples/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{B_.WORKLOAD}: BASELINE.json configs[1] table row-sharded over {world} GPUs "
                                   "(relations replicated); requests posted, rows pushed and deltas pulled by our kernels "
                                   "over NVLink peer memory; no NCCL call inside a step" if cls is P2PRowShardedTrainer else
                                   f"{B_.WORKLOAD}: table row-sharded over {world} GPUs, NCCL all-to-all of rows and row deltas",
                       "batch_per_gpu": Bl, "global_batch": Bl * world, "margin": B_.MARGIN, "lr0": B_.LR0,
                       "parallelism": f"rowshard{world}", "l2": "table shard larger than L2; no flush",
                       "host_enqueue_us_per_step": host_enqueue_us,
                       "mean_loss_last_step": losses[-1] / Bl if losses else None},
            "e2e": {"value": e2e, "unit": "triples/s", "h2d_bytes_per_step": 12 * Bl, "d2h_bytes_per_step": 4,
                    "call": "P2PRowShardedTrainer.train_step (pinned host triples in, loss sum out), per rank"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": value * alg / 1e9 / world, "peak": peak, "unit": "GB/s",
                         "frac": value * alg / 1e9 / world / peak, "traffic": None, "peak_source": src,
                         "kernel": "whole sharded step per GPU (NVLink exchange included)"},
            "cpu_baseline": None,
            "ranking": {"workload": f"20000 queries x 1,200,000 candidates sharded over {world} GPUs (tail side)",
                        "ms": ms_rank, "scores_per_s": 20000 * 1.2e6 / (ms_rank * 1e-3)},
        }))
    dist.barrier()
    dist.destroy_process_group()
