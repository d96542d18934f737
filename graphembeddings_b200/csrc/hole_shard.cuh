// Row-sharded training step over NVLink peer memory (no counterpart in the
// reference, which is single-device).  Included by hole_train.cu.
//
// Entity rows are partitioned into contiguous blocks, rank o owning rows
// [R + o*rows_per, R + (o+1)*rows_per); the R relation rows are replicated.  Every rank holds
// shard = [relations | its block] and maps (CUDA IPC) every peer's shard, delta staging buffer,
// relation staging buffer, request inbox and flag array.  One step of rank `me` on its slice of
// the global batch:
//
//   side stream (table independent; hole_shard_prepare works one step ahead):
//     Philox corruption keyed on the GLOBAL triple index -> route (dedup the 3B entity ids into
//     the request list `uniq`, per-owner cut points, triples re-indexed to request-list rows)
//     -> integer update plan
//   compute stream:
//     post    my request lists -> the owners' inboxes                       (peer stores, ids only)
//     K1      waits until every peer's shard is current (flag A), GATHERS the rows of its triples
//             straight from the owners' shards (bulk async copies over NVLink), computes, and
//             stores every unique row's delta into my slice of the owner's staging buffer
//             (peer stores) -- the exchange is fused into the compute kernel, tile by tile
//     K3      ordered combine of rows used more than once -> same staging buffers
//     finish  my relation deltas -> every peer's relation staging; then signals flag B
//             ("my deltas and request lists of this step are delivered") to every peer
//     apply   waits for flag B of every peer; adds the staged deltas to my shard -- a row
//             requested by several ranks gets its deltas in rank order, and the replicated
//             relation block takes every rank's deltas in rank order, so the result is
//             deterministic and the replicas stay bit-identical -- then signals flag A
//
// No NCCL call and no host synchronisation inside a step; two flag waits per step, both at the
// head of a kernel that has work to do anyway.  Flags only grow (epoch = steps completed).
#pragma once

constexpr int SHARD_LOSS_CHUNK = 64;     // steps whose loss rows are kept before they are summed

struct hole_shard_prep {                 // table-independent part of one step, built ahead
  int32_t* uniq = nullptr;               // [cap]      request list: sorted unique global entity rows
  int32_t* cuts = nullptr;               // [world+1]  first request-list slot per owner
  int32_t* pos_w = nullptr;              // [B,3]      triples as request-list rows (R + slot), relation kept
  int32_t* neg_w = nullptr;              // [B]
  int32_t* neg = nullptr;                // [B]        corrupt entity (global row)
  const int32_t* pos = nullptr;          // the triples this was built for
  int64_t B = -1;
  uint64_t seed = 0, step = 0;
  int side = 0;
  int plan_slot = 0;
  bool valid = false, used = false;
  cudaEvent_t freed = nullptr;           // recorded when the step that consumed this slot has been enqueued
};

struct hole_shard_state {
  int world = 0, me = 0;
  int64_t R = 0, n_ent = 0, rows_per = 0, max_batch = 0, cap = 0;
  float* shard = nullptr;
  hole_peer_ptrs p_shard, p_stage, p_relstage, p_inbox, p_meta, p_flags;
  int32_t* err = nullptr;
  unsigned long long timeout_ns = 600ull * 1000000000ull;
  const int32_t* type_of = nullptr;
  const int64_t* csr_off = nullptr;
  const int32_t* csr_ids = nullptr;
  float* Drel = nullptr;                 // [R, stride] my relation deltas of the running step (zero between steps)
  unsigned* done = nullptr;              // [2] last-block-done counters (finish, apply)
  float* loss = nullptr;                 // [SHARD_LOSS_CHUNK, max_batch]
  float* loss_sum = nullptr;             // [SHARD_LOSS_CHUNK]
  int32_t* stage_tri = nullptr;          // device staging of hole_shard_steps_host
  int64_t stage_cap = 0;
  float* loss_pinned = nullptr;          // pinned host + device mirrors of a call's loss sums
  float* sums_dev = nullptr;
  int64_t pinned_cap = 0;
  hole_shard_prep prep[2];
  int prep_toggle = 0;
  int epoch = 0;                         // steps completed
};

// ------------------------------------------------------------------------------------ kernels
// relation deltas of this step -> every peer's relation staging [world][R][stride] (slice `me`),
// my delta rows back to zero; the last block to finish signals flag B = epoch to every peer
template <int GS, int V>
__global__ void __launch_bounds__(256)
hole_shard_finish_kernel(float* __restrict__ Drel, int R, int me, int world, hole_peer_ptrs relstage,
                         hole_peer_ptrs flags, int epoch, unsigned* __restrict__ done, int nvec, int stride) {
  __shared__ bool s_last;
  const int lane = threadIdx.x % GS;
  const int gstride = (gridDim.x * blockDim.x) / GS;
  Row<V> z;
  row_zero(z);
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) / GS; r < R; r += gstride) {
    Row<V> d;
    row_load<GS, V, false>(d, Drel + (size_t)r * stride, lane, nvec);
    row_store<GS, V>(z, Drel + (size_t)r * stride, lane, nvec);
    for (int k = 0; k < world; ++k)
      row_store<GS, V>(d, static_cast<float*>(relstage.p[k]) + ((size_t)me * R + r) * stride, lane, nvec);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) *done = 0;                    // self-reset for the next step
    __threadfence_system();
    if (threadIdx.x < world) {
      int* theirs = static_cast<int*>(flags.p[threadIdx.x]) + world + me;     // flags[1][me] on that rank
      asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    }
  }
}

// first index in ids[0, n) with ids[.] >= x
__device__ __forceinline__ int shard_lower_bound(const int32_t* __restrict__ ids, int n, int x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (ids[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Owner side.  Items [0, R): replicated relation rows, shard[r] += relstage[k][r] for k = 0..world-1 in
// order.  Items R + e: entry e of the concatenated request lists (inbox[k][g], every list ascending);
// the group of the LOWEST rank that lists a row adds the staged deltas of every rank that lists it,
// in rank order; the other groups of that row do nothing.  The last block signals flag A.
template <int GS, int V>
__global__ void __launch_bounds__(256)
hole_shard_apply_kernel(float* __restrict__ shard, int64_t id_offset, const int32_t* __restrict__ inbox,
                        const int32_t* __restrict__ meta, const float* __restrict__ stage,
                        const float* __restrict__ relstage, int R, int world, int me, int64_t cap,
                        const int* __restrict__ my_flags, int wait_epoch, int* __restrict__ err,
                        unsigned long long timeout_ns, hole_peer_ptrs flags, int signal_epoch,
                        unsigned* __restrict__ done, int nvec, int stride) {
  constexpr int CH = (HOLE_MAX_RANKS + GS - 1) / GS;      // ranks per lane
  __shared__ int s_n[HOLE_MAX_RANKS], s_pre[HOLE_MAX_RANKS + 1];
  __shared__ bool s_last;
  hole_flags_wait(my_flags + world, world, wait_epoch, err, timeout_ns);     // every peer's deltas are here
  if (threadIdx.x == 0) {
    int pre = 0;
    for (int k = 0; k < world; ++k) {
      const int n = meta[2 * k];
      s_n[k] = n;
      s_pre[k] = pre;
      pre += n;
    }
    s_pre[world] = pre;
  }
  __syncthreads();
  const int lane = threadIdx.x % GS;
  const int gbase = (threadIdx.x % 32) / GS * GS;
  const unsigned gmask = (GS == 32) ? 0xffffffffu : (((1u << GS) - 1u) << gbase);
  const int total = R + s_pre[world];
  const int gstride = (gridDim.x * blockDim.x) / GS;
  for (int item = (blockIdx.x * blockDim.x + threadIdx.x) / GS; item < total; item += gstride) {
    if (item < R) {
      float* erow = shard + (size_t)item * stride;
      Row<V> x, d;
      row_load<GS, V, false>(x, erow, lane, nvec);
      for (int k = 0; k < world; ++k) {
        row_load<GS, V, false>(d, relstage + ((size_t)k * R + item) * stride, lane, nvec);
        row_add(x, d);
      }
      row_store<GS, V>(x, erow, lane, nvec);
      continue;
    }
    const int e = item - R;
    int k = 0;
    while (e >= s_pre[k + 1]) ++k;
    const int g = e - s_pre[k];
    const int id = inbox[(size_t)k * cap + g];
    // lane l (+ GS, ...) looks the row up in rank l's list
    int pos[CH];
    unsigned present = 0;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const int l = ch * GS + lane;
      int p = -1;
      if (l < world) {
        if (l == k) p = g;
        else {
          const int32_t* ids = inbox + (size_t)l * cap;
          const int q = shard_lower_bound(ids, s_n[l], id);
          if (q < s_n[l] && ids[q] == id) p = q;
        }
      }
      pos[ch] = p;
      const unsigned b = (__ballot_sync(gmask, p >= 0) >> gbase) & ((GS == 32) ? 0xffffffffu : ((1u << GS) - 1u));
      present |= b << (ch * GS);
    }
    if (__ffs(present) - 1 != k) continue;               // a lower rank's group owns this row
    float* erow = shard + (size_t)(id + id_offset) * stride;
    Row<V> x, d;
    row_load<GS, V, false>(x, erow, lane, nvec);
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      for (int j = 0; j < GS; ++j) {
        const int l = ch * GS + j;
        if (l >= world) break;
        const int p = __shfl_sync(gmask, pos[ch], gbase + j);
        if (p < 0) continue;
        row_load<GS, V, false>(d, stage + ((size_t)l * cap + p) * stride, lane, nvec);
        row_add(x, d);
      }
    }
    row_store<GS, V>(x, erow, lane, nvec);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) *done = 0;
    __threadfence_system();
    if (threadIdx.x < world) {
      int* theirs = static_cast<int*>(flags.p[threadIdx.x]) + me;             // flags[0][me] on that rank
      asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(theirs), "r"(signal_epoch) : "memory");
    }
  }
}

// ------------------------------------------------------------------------------------ host side
static void shard_free(hole_ctx* c) {
  hole_shard_state* s = c->shard_state;
  if (s == nullptr) return;
  for (int b = 0; b < 2; ++b) {
    cudaFree(s->prep[b].uniq); cudaFree(s->prep[b].cuts); cudaFree(s->prep[b].pos_w);
    cudaFree(s->prep[b].neg_w); cudaFree(s->prep[b].neg);
    if (s->prep[b].freed) cudaEventDestroy(s->prep[b].freed);
  }
  cudaFree(s->Drel); cudaFree(s->done); cudaFree(s->loss); cudaFree(s->loss_sum); cudaFree(s->stage_tri);
  cudaFree(s->sums_dev);
  if (s->loss_pinned) cudaFreeHost(s->loss_pinned);
  delete s;
  c->shard_state = nullptr;
}

extern "C" int hole_shard_init(hole_ctx* c, int world, int me, int64_t n_relations, int64_t n_entities,
                               int64_t rows_per_rank, int64_t max_batch, float* shard,
                               void* const* peer_shard, void* const* peer_stage, void* const* peer_relstage,
                               void* const* peer_inbox, void* const* peer_meta, void* const* peer_flags,
                               int32_t* err_flag, double timeout_s, const int32_t* type_of,
                               const int64_t* csr_off, const int32_t* csr_ids) {
  HOLE_CHECK_ARG(c && shard && peer_shard && peer_stage && peer_relstage && peer_inbox && peer_meta && peer_flags);
  HOLE_CHECK_ARG(err_flag && type_of && csr_off && csr_ids);
  HOLE_CHECK_ARG(world >= 1 && world <= HOLE_MAX_RANKS && me >= 0 && me < world);
  HOLE_CHECK_ARG(n_relations >= 0 && n_entities > 0 && rows_per_rank > 0 && rows_per_rank * world >= n_entities);
  HOLE_CHECK_ARG(n_relations + rows_per_rank * world < (int64_t(1) << 31));
  HOLE_CHECK_ARG(max_batch > 0 && 4 * max_batch < (int64_t(1) << 31));
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  shard_free(c);
  int rc = hole_ws_reserve(c, max_batch, 1);
  if (rc) return rc;
  hole_shard_state* s = new hole_shard_state();
  c->shard_state = s;
  s->world = world; s->me = me; s->R = n_relations; s->n_ent = n_entities; s->rows_per = rows_per_rank;
  s->max_batch = max_batch; s->cap = 3 * max_batch; s->shard = shard; s->err = err_flag;
  if (timeout_s > 0) s->timeout_ns = (unsigned long long)(timeout_s * 1e9);
  s->type_of = type_of; s->csr_off = csr_off; s->csr_ids = csr_ids;
  if ((rc = peer_ptrs(s->p_shard, peer_shard, world)) || (rc = peer_ptrs(s->p_stage, peer_stage, world)) ||
      (rc = peer_ptrs(s->p_relstage, peer_relstage, world)) || (rc = peer_ptrs(s->p_inbox, peer_inbox, world)) ||
      (rc = peer_ptrs(s->p_meta, peer_meta, world)) || (rc = peer_ptrs(s->p_flags, peer_flags, world))) {
    shard_free(c);
    return rc;
  }
  if (s->p_shard.p[me] != shard) {
    shard_free(c);
    return hole_set_error(HOLE_ERR_ARG, "hole_shard_init: peer_shard[me] must be this rank's shard");
  }
#define SH_ALLOC(ptr, bytes)                                                             \
  do {                                                                                   \
    if (cudaMalloc((void**)&(ptr), (bytes)) != cudaSuccess) {                            \
      cudaGetLastError();                                                                \
      shard_free(c);                                                                     \
      return hole_set_error(HOLE_ERR_ALLOC, "sharded-step allocation of %zu bytes failed", (size_t)(bytes)); \
    }                                                                                    \
  } while (0)
  const size_t rowb = (size_t)c->row_stride * sizeof(float);
  SH_ALLOC(s->Drel, std::max<size_t>(1, (size_t)s->R) * rowb);
  SH_ALLOC(s->done, 2 * sizeof(unsigned));
  SH_ALLOC(s->loss, (size_t)SHARD_LOSS_CHUNK * max_batch * 4);
  SH_ALLOC(s->loss_sum, (size_t)SHARD_LOSS_CHUNK * 4);
  for (int b = 0; b < 2; ++b) {
    SH_ALLOC(s->prep[b].uniq, (size_t)s->cap * 4);
    SH_ALLOC(s->prep[b].cuts, (size_t)(HOLE_MAX_RANKS + 1) * 4);
    SH_ALLOC(s->prep[b].pos_w, (size_t)max_batch * 12);
    SH_ALLOC(s->prep[b].neg_w, (size_t)max_batch * 4);
    SH_ALLOC(s->prep[b].neg, (size_t)max_batch * 4);
    s->prep[b].plan_slot = b;
    HOLE_CUDA_TRY(cudaEventCreateWithFlags(&s->prep[b].freed, cudaEventDisableTiming));
  }
#undef SH_ALLOC
  HOLE_CUDA_TRY(cudaMemset(s->Drel, 0, std::max<size_t>(1, (size_t)s->R) * rowb));
  HOLE_CUDA_TRY(cudaMemset(s->done, 0, 2 * sizeof(unsigned)));
  rc = route_reserve(c, 3 * max_batch);
  if (rc) { shard_free(c); return rc; }
  HOLE_CUDA_TRY(cudaDeviceSynchronize());
  return HOLE_OK;
}

// corruption + routing + update plan of one step on the library's side stream, once `st` has reached
// this point (pos must be complete by then)
static int shard_prepare(hole_ctx* c, const int32_t* pos, int64_t B, uint64_t seed, uint64_t step, cudaStream_t st) {
  hole_shard_state* s = c->shard_state;
  hole_shard_prep& p = s->prep[s->prep_toggle];
  s->prep_toggle ^= 1;
  cudaStream_t ps = c->plan_stream;
  HOLE_CUDA_TRY(cudaEventRecord(c->ev_entry, st));
  HOLE_CUDA_TRY(cudaStreamWaitEvent(ps, c->ev_entry, 0));
  if (p.used) HOLE_CUDA_TRY(cudaStreamWaitEvent(ps, p.freed, 0));      // its last consumer has been enqueued and is done
  p.pos = pos; p.B = B; p.seed = seed; p.step = step;
  p.side = hole_side_coin(seed, step);
  const uint64_t index_base = (uint64_t)s->me * (uint64_t)B;           // my slice of the global batch
  dim3 grid((unsigned)std::min<int64_t>((B + 255) / 256, 65535), 1);
  hole_corrupt_kernel<<<grid, 256, 0, ps>>>(pos, B, 1, s->type_of, s->csr_off, s->csr_ids, seed, step, index_base,
                                            p.neg, nullptr);
  HOLE_LAUNCHED();
  int rc = shard_route(c, pos, p.neg, B, s->R, s->R + s->rows_per * s->world, s->rows_per, s->world, p.uniq, p.cuts,
                       p.pos_w, p.neg_w, ps);
  if (rc) return rc;
  hole_plan& pl = c->plan[p.plan_slot];
  pl.prepared_B = -1;
  pl.prepared_pos = pl.prepared_neg = nullptr;
  rc = plan_steps(c, pl, p.pos_w, B, 1, nullptr, nullptr, nullptr, 0, 0, p.neg_w, ps);
  if (rc) return rc;
  p.valid = true;
  return HOLE_OK;
}

static int shard_check_args(hole_ctx* c, const int32_t* pos, int64_t B) {
  HOLE_CHECK_ARG(c != nullptr);
  if (c->shard_state == nullptr) return hole_set_error(HOLE_ERR_ARG, "hole_shard_init has not been called");
  HOLE_CHECK_ARG(pos != nullptr && B > 0 && B <= c->shard_state->max_batch);
  return HOLE_OK;
}

extern "C" int hole_shard_prepare(hole_ctx* c, const int32_t* pos, int64_t B, uint64_t seed, uint64_t step,
                                  void* stream) {
  int rc = shard_check_args(c, pos, B);
  if (rc) return rc;
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  return shard_prepare(c, pos, B, seed, step, (cudaStream_t)stream);
}

// post + K1 + K3 + finish of the step
extern "C" int hole_shard_step_compute(hole_ctx* c, const int32_t* pos, int64_t B, uint64_t seed, uint64_t step,
                                       float margin, float lr, float* loss_out, void* stream) {
  int rc = shard_check_args(c, pos, B);
  if (rc) return rc;
  HOLE_CHECK_ARG(loss_out != nullptr);
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  hole_shard_state* s = c->shard_state;
  cudaStream_t st = (cudaStream_t)stream;
  int pi = -1;
  for (int b = 0; b < 2; ++b)
    if (s->prep[b].valid && s->prep[b].pos == pos && s->prep[b].B == B && s->prep[b].seed == seed &&
        s->prep[b].step == step)
      pi = b;
  if (pi < 0) {
    pi = s->prep_toggle;
    rc = shard_prepare(c, pos, B, seed, step, st);
    if (rc) return rc;
  }
  hole_shard_prep& p = s->prep[pi];
  hole_plan& pl = c->plan[p.plan_slot];
  HOLE_CUDA_TRY(cudaStreamWaitEvent(st, pl.ready, 0));       // corruption, routing and plan are complete
  const int par = s->epoch & 1;
  const int world = s->world, me = s->me;
  // my request lists -> the owners' inboxes of this step's parity
  hole_peer_ptrs ib, mt;
  for (int k = 0; k < HOLE_MAX_RANKS; ++k) {
    ib.p[k] = k < world ? static_cast<int32_t*>(s->p_inbox.p[k]) + (size_t)par * world * s->cap : nullptr;
    mt.p[k] = k < world ? static_cast<int32_t*>(s->p_meta.p[k]) + (size_t)par * world * 2 : nullptr;
  }
  hole_shard_post_kernel<<<(unsigned)std::min<int64_t>((3 * B + 255) / 256, c->sm_count * 2), 256, 0, st>>>(
      p.uniq, p.cuts, world, me, s->cap, ib, mt);
  HOLE_LAUNCHED();
  hole_k1_shard sh = {};
  sh.tri_w = p.pos_w; sh.neg_w = p.neg_w; sh.cuts = p.cuts;
  sh.flags = static_cast<const int*>(s->p_flags.p[me]);
  sh.err = s->err; sh.wait_epoch = s->epoch; sh.R = (int)s->R; sh.rows_per = (int)s->rows_per;
  sh.me = me; sh.world = world; sh.cap = s->cap; sh.timeout_ns = s->timeout_ns;
  sh.shard = s->p_shard; sh.stage = s->p_stage;
  rc = run_step(c, pl, s->shard, p.pos_w, p.neg_w, p.side, B, margin, lr, loss_out, nullptr, 0, st, s->Drel, false, 0,
                &sh, pos, p.neg);
  if (rc) return rc;
  pl.used = true;
  HOLE_CUDA_TRY(cudaEventRecord(pl.released, st));
  p.used = true;
  p.valid = false;
  HOLE_CUDA_TRY(cudaEventRecord(p.freed, st));
  const unsigned fgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((s->R + (256 / c->gs) - 1) / (256 / c->gs), c->sm_count));
  HOLE_DISPATCH(c, hole_shard_finish_kernel, fgrid, 256, st, s->Drel, (int)s->R, me, world, s->p_relstage, s->p_flags,
                s->epoch + 1, s->done, c->nvec, c->row_stride);
  return HOLE_OK;
}

// apply of the step: every peer's deltas -> my shard
extern "C" int hole_shard_step_apply(hole_ctx* c, void* stream) {
  HOLE_CHECK_ARG(c != nullptr);
  if (c->shard_state == nullptr) return hole_set_error(HOLE_ERR_ARG, "hole_shard_init has not been called");
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  hole_shard_state* s = c->shard_state;
  const int par = s->epoch & 1, world = s->world, me = s->me;
  const int32_t* inbox = static_cast<const int32_t*>(s->p_inbox.p[me]) + (size_t)par * world * s->cap;
  const int32_t* meta = static_cast<const int32_t*>(s->p_meta.p[me]) + (size_t)par * world * 2;
  HOLE_DISPATCH(c, hole_shard_apply_kernel, (unsigned)c->sm_count * 4, 256, (cudaStream_t)stream, s->shard,
                s->R - (s->R + (int64_t)me * s->rows_per), inbox, meta, static_cast<const float*>(s->p_stage.p[me]),
                static_cast<const float*>(s->p_relstage.p[me]), (int)s->R, world, me, s->cap,
                static_cast<const int*>(s->p_flags.p[me]), s->epoch + 1, s->err, s->timeout_ns, s->p_flags,
                s->epoch + 1, s->done + 1, c->nvec, c->row_stride);
  s->epoch += 1;
  return HOLE_OK;
}

extern "C" int hole_shard_step(hole_ctx* c, const int32_t* pos, int64_t B, uint64_t seed, uint64_t step,
                               float margin, float lr, float* loss_out, void* stream) {
  int rc = hole_shard_step_compute(c, pos, B, seed, step, margin, lr, loss_out, stream);
  if (rc) return rc;
  return hole_shard_step_apply(c, stream);
}

// n_steps consecutive steps on device-resident triples [n_steps*B, 3] (this rank's slices), the
// table-independent part of step k+1 built while step k runs.  loss_sum_out: device float32[n_steps].
static int shard_steps(hole_ctx* c, const int32_t* triples, int64_t B, int64_t n_steps, uint64_t seed,
                       uint64_t first_step, float margin, const float* lr, float* loss_sum_out,
                       cudaStream_t st, const cudaEvent_t* chunk_ready, int64_t chunk_steps) {
  hole_shard_state* s = c->shard_state;
  int rc;
  for (int64_t k0 = 0; k0 < n_steps; k0 += SHARD_LOSS_CHUNK) {
    const int64_t n = std::min<int64_t>(SHARD_LOSS_CHUNK, n_steps - k0);
    for (int64_t k = k0; k < k0 + n; ++k) {
      if (chunk_ready && k % chunk_steps == 0) HOLE_CUDA_TRY(cudaStreamWaitEvent(st, chunk_ready[k / chunk_steps], 0));
      if (k == 0) {
        rc = shard_prepare(c, triples, B, seed, first_step, st);
        if (rc) return rc;
      }
      if (k + 1 < n_steps) {
        if (chunk_ready && (k + 1) % chunk_steps == 0)
          HOLE_CUDA_TRY(cudaStreamWaitEvent(st, chunk_ready[(k + 1) / chunk_steps], 0));
        rc = shard_prepare(c, triples + (size_t)(k + 1) * B * 3, B, seed, first_step + (uint64_t)(k + 1), st);
        if (rc) return rc;
      }
      rc = hole_shard_step(c, triples + (size_t)k * B * 3, B, seed, first_step + (uint64_t)k, margin, lr[k],
                           s->loss + (size_t)(k - k0) * B, st);
      if (rc) return rc;
    }
    if (loss_sum_out != nullptr) {
      hole_loss_sum_kernel<<<(unsigned)n, 256, 0, st>>>(s->loss, B, loss_sum_out + k0);
      HOLE_LAUNCHED();
    }
  }
  return HOLE_OK;
}

extern "C" int hole_shard_steps(hole_ctx* c, const int32_t* triples, int64_t B, int64_t n_steps, uint64_t seed,
                                uint64_t first_step, float margin, const float* lr, float* loss_sum_out,
                                void* stream) {
  if (n_steps == 0) return HOLE_OK;
  int rc = shard_check_args(c, triples, B);
  if (rc) return rc;
  HOLE_CHECK_ARG(n_steps > 0 && lr != nullptr);
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  return shard_steps(c, triples, B, n_steps, seed, first_step, margin, lr, loss_sum_out, (cudaStream_t)stream,
                     nullptr, 0);
}

// Same, end to end from HOST triples (pinned or pageable): copied to the device in chunks that overlap
// the steps, per-step loss sums copied back; blocks until they are on the host.
extern "C" int hole_shard_steps_host(hole_ctx* c, const int32_t* triples_host, int64_t B, int64_t n_steps,
                                     uint64_t seed, uint64_t first_step, float margin, const float* lr,
                                     float* loss_sum_host, void* stream) {
  if (n_steps == 0) return HOLE_OK;
  int rc = shard_check_args(c, triples_host, B);
  if (rc) return rc;
  HOLE_CHECK_ARG(n_steps > 0 && lr != nullptr && loss_sum_host != nullptr);
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  hole_shard_state* s = c->shard_state;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t elems = n_steps * B * 3;
  if (elems > s->stage_cap) {
    HOLE_CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(s->stage_tri);
    s->stage_tri = nullptr;
    s->stage_cap = 0;
    if (cudaMalloc((void**)&s->stage_tri, (size_t)elems * 4) != cudaSuccess) {
      cudaGetLastError();
      return hole_set_error(HOLE_ERR_ALLOC, "triple staging allocation failed");
    }
    s->stage_cap = elems;
  }
  if (n_steps > s->pinned_cap) {
    HOLE_CUDA_TRY(cudaDeviceSynchronize());
    if (s->loss_pinned) cudaFreeHost(s->loss_pinned);
    cudaFree(s->sums_dev);
    s->loss_pinned = nullptr;
    s->sums_dev = nullptr;
    s->pinned_cap = 0;
    HOLE_CUDA_TRY(cudaMallocHost((void**)&s->loss_pinned, (size_t)n_steps * sizeof(float)));
    HOLE_CUDA_TRY(cudaMalloc((void**)&s->sums_dev, (size_t)n_steps * sizeof(float)));
    s->pinned_cap = n_steps;
  }
  float* sums = s->sums_dev;
  // H2D in chunks of 16 steps on the copy stream; a step waits for its chunk only
  const int64_t CS = 16, nchunks = (n_steps + CS - 1) / CS;
  std::vector<cudaEvent_t> ready((size_t)nchunks);
  HOLE_CUDA_TRY(cudaEventRecord(c->ev_entry, st));
  HOLE_CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_entry, 0));   // the staging buffer's last readers are done
  for (int64_t ci = 0; ci < nchunks; ++ci) {
    const int64_t k0 = ci * CS, n = std::min(CS, n_steps - k0);
    HOLE_CUDA_TRY(cudaMemcpyAsync(s->stage_tri + (size_t)k0 * B * 3, triples_host + (size_t)k0 * B * 3,
                                  (size_t)n * B * 12, cudaMemcpyHostToDevice, c->copy_stream));
    HOLE_CUDA_TRY(cudaEventCreateWithFlags(&ready[ci], cudaEventDisableTiming));
    HOLE_CUDA_TRY(cudaEventRecord(ready[ci], c->copy_stream));
  }
  rc = shard_steps(c, s->stage_tri, B, n_steps, seed, first_step, margin, lr, sums, st, ready.data(), CS);
  if (rc == HOLE_OK) {
    cudaError_t e = cudaMemcpyAsync(s->loss_pinned, sums, (size_t)n_steps * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = hole_set_error(HOLE_ERR_CUDA, "hole_shard_steps_host: %s", cudaGetErrorString(e));
  } else {
    cudaStreamSynchronize(st);
  }
  for (cudaEvent_t e : ready) cudaEventDestroy(e);
  if (rc) return rc;
  for (int64_t k = 0; k < n_steps; ++k) loss_sum_host[k] = s->loss_pinned[k];
  return HOLE_OK;
}

// Host check of the flag time-out (synchronises the stream): nonzero *timed_out means a peer never
// arrived and the tables are no longer consistent -- stop.
extern "C" int hole_shard_poll(hole_ctx* c, int* timed_out, void* stream) {
  HOLE_CHECK_ARG(c && timed_out);
  if (c->shard_state == nullptr) return hole_set_error(HOLE_ERR_ARG, "hole_shard_init has not been called");
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  int32_t v = 0;
  HOLE_CUDA_TRY(cudaMemcpyAsync(&v, c->shard_state->err, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  HOLE_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  *timed_out = v;
  return HOLE_OK;
}
