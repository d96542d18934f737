// HolE training hot path for sm_100a: Philox type-safe corruption (K2), update-plan sort,
// fused gather + norm-clip + Hermitian score + sigmoid + hinge + backward (K1) and the
// deterministic tree-combined sparse SGD update (K3).
//
// Reference seams: corrupt_batch holE.py:97-158 (+ host subsample 343-347),
// get_embedding 161-168, evaluate_triples 179-202, evaluate_batch 222-234,
// GradientDescentOptimizer.minimize 296.  Arithmetic: SURVEY.md App. A/B.
//
// All kernels are HBM/L2-bandwidth bound integer/fp32 work: no tensor cores here.  A row
// is handled by a group of GS lanes, each owning V float4s of the real half and the
// matching V float4s of the imaginary half, so every complex product is lane-local and
// every global access is a coalesced 16-byte vector.
#include <stdarg.h>

#include <algorithm>

#include "hole_common.cuh"

thread_local std::string g_hole_err;
thread_local int64_t g_hole_launches = 0;

int hole_set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_hole_err = buf;
  return code;
}

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
template <int V>
struct Row {
  float re[4 * V];
  float im[4 * V];
};

__device__ __forceinline__ float4 ld_nc(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float4 ld_cg(const float4* p) { return __ldcg(p); }

template <int GS, int V, bool NC>
__device__ __forceinline__ void row_load(Row<V>& x, const float* base, int lane, int nvec) {
  const float4* p = reinterpret_cast<const float4*>(base);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    int idx = lane + v * GS;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (idx < nvec) {
      a = NC ? ld_nc(p + idx) : ld_cg(p + idx);
      b = NC ? ld_nc(p + nvec + idx) : ld_cg(p + nvec + idx);
    }
    x.re[4 * v + 0] = a.x; x.re[4 * v + 1] = a.y; x.re[4 * v + 2] = a.z; x.re[4 * v + 3] = a.w;
    x.im[4 * v + 0] = b.x; x.im[4 * v + 1] = b.y; x.im[4 * v + 2] = b.z; x.im[4 * v + 3] = b.w;
  }
}

template <int GS, int V>
__device__ __forceinline__ void row_store(const Row<V>& x, float* base, int lane, int nvec) {
  float4* p = reinterpret_cast<float4*>(base);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    int idx = lane + v * GS;
    if (idx < nvec) {
      p[idx] = make_float4(x.re[4 * v], x.re[4 * v + 1], x.re[4 * v + 2], x.re[4 * v + 3]);
      p[nvec + idx] = make_float4(x.im[4 * v], x.im[4 * v + 1], x.im[4 * v + 2], x.im[4 * v + 3]);
    }
  }
}

template <int GS>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = GS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int V>
__device__ __forceinline__ float row_sumsq(const Row<V>& x) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 4 * V; ++k) s = fmaf(x.re[k], x.re[k], s);
#pragma unroll
  for (int k = 0; k < 4 * V; ++k) s = fmaf(x.im[k], x.im[k], s);
  return s;
}

template <int V>
__device__ __forceinline__ float row_dot(const Row<V>& x, const Row<V>& y) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 4 * V; ++k) s = fmaf(x.re[k], y.re[k], s);
#pragma unroll
  for (int k = 0; k < 4 * V; ++k) s = fmaf(x.im[k], y.im[k], s);
  return s;
}

// clip_by_norm(row, 1): y = x * min(rsqrt(sum x^2), 1)  (App. B; holE.py:162 max_norm=1).
// Returns inv = rsqrt(sum x^2); the row is "clipped" (gradient takes the rsqrt branch)
// when inv <= 1.
template <int V>
__device__ __forceinline__ void row_clip(Row<V>& x, float inv) {
  float sc = fminf(inv, 1.0f);
#pragma unroll
  for (int k = 0; k < 4 * V; ++k) { x.re[k] *= sc; x.im[k] *= sc; }
}

__device__ __forceinline__ float sigmoidf_precise(float s) { return 1.0f / (1.0f + expf(-s)); }

// dx = clipped ? (dy - y (y.dy)) * inv : dy      (App. A.3)
template <int V>
__device__ __forceinline__ void clip_backward(Row<V>& dy, const Row<V>& y, float proj, float inv) {
  if (inv <= 1.0f) {
#pragma unroll
    for (int k = 0; k < 4 * V; ++k) {
      dy.re[k] = (dy.re[k] - y.re[k] * proj) * inv;
      dy.im[k] = (dy.im[k] - y.im[k] * proj) * inv;
    }
  }
}

// ---------------------------------------------------------------------------------------
// layout conversion
// ---------------------------------------------------------------------------------------
__global__ void hole_pack_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                      int64_t n, int dim, int H, int Hp) {
  int64_t total = n * 2 * Hp;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = t / (2 * Hp);
    int c = (int)(t % (2 * Hp));
    int half = c / Hp, k = c % Hp;
    dst[t] = (k < H) ? src[row * dim + half * H + k] : 0.0f;
  }
}

__global__ void hole_unpack_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                        int64_t n, int dim, int H, int Hp) {
  int64_t total = n * dim;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = t / dim;
    int c = (int)(t % dim);
    int half = c / H, k = c % H;
    dst[t] = src[row * 2 * Hp + half * Hp + k];
  }
}

// ---------------------------------------------------------------------------------------
// K2: corruption sampler + update-plan keys
// ---------------------------------------------------------------------------------------
// For step s (grid.y) and triple i: neg = csr_ids[off[ty] + draw]; emits the four
// (row, position) sort keys of the step: position = slot*B + i with slots
// [relation, tail-slot, head-slot, corrupt entity].
__global__ void hole_corrupt_kernel(const int32_t* __restrict__ triples, int64_t B, int n_steps,
                                    const int32_t* __restrict__ type_of,
                                    const int64_t* __restrict__ csr_off,
                                    const int32_t* __restrict__ csr_ids, uint64_t seed,
                                    uint64_t first_step, int32_t* __restrict__ neg_out,
                                    int32_t* __restrict__ side_out, uint32_t* __restrict__ keys) {
  int s = blockIdx.y;
  uint64_t step = first_step + (uint64_t)s;
  int side = hole_side_coin(seed, step);
  if (side_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) side_out[s] = side;
  const int32_t* tr = triples + (size_t)s * B * 3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B;
       i += (int64_t)gridDim.x * blockDim.x) {
    int h = tr[3 * i], t = tr[3 * i + 1], r = tr[3 * i + 2];
    int ent = side ? h : t;
    int ty = type_of[ent];
    int64_t lo = csr_off[ty];
    uint32_t cnt = (uint32_t)(csr_off[ty + 1] - lo);
    uint32_t j = hole_entity_draw(seed, step, (uint32_t)i, cnt);
    int n = csr_ids[lo + j];
    neg_out[(size_t)s * B + i] = n;
    if (keys != nullptr) {
      uint32_t* k = keys + (size_t)s * 4 * B;
      k[i] = (uint32_t)r;
      k[B + i] = (uint32_t)t;
      k[2 * B + i] = (uint32_t)h;
      k[3 * B + i] = (uint32_t)n;
    }
  }
}

// Keys for a caller-supplied corruption (hole_train_step).
__global__ void hole_keys_kernel(const int32_t* __restrict__ pos, const int32_t* __restrict__ neg,
                                 int64_t B, uint32_t* __restrict__ k) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B;
       i += (int64_t)gridDim.x * blockDim.x) {
    k[i] = (uint32_t)pos[3 * i + 2];
    k[B + i] = (uint32_t)pos[3 * i + 1];
    k[2 * B + i] = (uint32_t)pos[3 * i];
    k[3 * B + i] = (uint32_t)neg[i];
  }
}

// ---------------------------------------------------------------------------------------
// update plan: one CTA sorts one step's 4B (row, position) pairs by row, stably (LSD
// radix, 8-bit digits), then marks segments.  Integer-only; runs for many steps at once
// ahead of the steps that consume it.
// ---------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 1024;
constexpr int SORT_WARPS = SORT_THREADS / 32;

__global__ void __launch_bounds__(SORT_THREADS, 1)
hole_plan_sort_kernel(uint32_t* keysA, uint32_t* valsA, uint32_t* keysB, uint32_t* valsB,
                      uint32_t* sstart, uint32_t* slen, int M, int passes) {
  __shared__ uint32_t hist[256 * SORT_WARPS];   // [digit][warp]
  __shared__ uint32_t wsum[SORT_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const size_t base = (size_t)blockIdx.x * M;
  uint32_t* kin = keysA + base;
  uint32_t* vin = valsA + base;
  uint32_t* kout = keysB + base;
  uint32_t* vout = valsB + base;
  // each warp owns a contiguous slice (multiple of 32 entries)
  const int per_warp = ((M + SORT_WARPS - 1) / SORT_WARPS + 31) & ~31;
  const int w_lo = min(M, w * per_warp), w_hi = min(M, w_lo + per_warp);

  for (int pass = 0; pass < passes; ++pass) {
    const int shift = 8 * pass;
    for (int i = tid; i < 256 * SORT_WARPS; i += SORT_THREADS) hist[i] = 0;
    __syncthreads();
    // 1. per-warp digit histogram
    for (int j0 = w_lo; j0 < w_hi; j0 += 32) {
      int j = j0 + lane;
      bool ok = j < w_hi;
      uint32_t d = ok ? ((kin[j] >> shift) & 255u) : 256u + lane;   // inactive lanes: unique
      uint32_t peers = __match_any_sync(0xffffffffu, d);
      if (ok && lane == __ffs(peers) - 1) hist[d * SORT_WARPS + w] += __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    // 2. exclusive scan in (digit-major, warp-minor) order: 8 entries per thread
    uint32_t loc[8];
    uint32_t tsum = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { loc[q] = hist[tid * 8 + q]; tsum += loc[q]; }
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += y;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
      uint32_t x = wsum[lane];
      uint32_t xi = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, xi, o);
        if (lane >= o) xi += y;
      }
      wsum[lane] = xi - x;
    }
    __syncthreads();
    uint32_t run = wsum[w] + inc - tsum;
#pragma unroll
    for (int q = 0; q < 8; ++q) { hist[tid * 8 + q] = run; run += loc[q]; }
    __syncthreads();
    // 3. stable scatter
    for (int j0 = w_lo; j0 < w_hi; j0 += 32) {
      int j = j0 + lane;
      bool ok = j < w_hi;
      uint32_t key = ok ? kin[j] : 0u;
      uint32_t val = ok ? (pass == 0 ? (uint32_t)j : vin[j]) : 0u;
      uint32_t d = ok ? ((key >> shift) & 255u) : 256u + lane;
      uint32_t peers = __match_any_sync(0xffffffffu, d);
      uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      uint32_t dst = 0;
      if (ok) {
        dst = hist[d * SORT_WARPS + w] + rank;
        kout[dst] = key;
        vout[dst] = val;
      }
      __syncwarp();
      if (ok && lane == __ffs(peers) - 1) hist[d * SORT_WARPS + w] += __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    uint32_t* t;
    t = kin; kin = kout; kout = t;
    t = vin; vin = vout; vout = t;
    __threadfence_block();
  }
  // sorted pairs are now in (kin, vin).  Host guarantees that is (keysA, valsA) by choosing
  // an even pass count.
  // 4. segment starts: sstart[j] = first sorted index holding the same row as j.
  const int per_thr = (M + SORT_THREADS - 1) / SORT_THREADS;
  const int t_lo = min(M, tid * per_thr), t_hi = min(M, t_lo + per_thr);
  int last_head = -1;   // last segment head inside my chunk
  for (int j = t_lo; j < t_hi; ++j)
    if (j == 0 || kin[j - 1] != kin[j]) last_head = j;
  // block-wide inclusive max-scan of last_head
  int incm = last_head;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, incm, o);
    if (lane >= o) incm = max(incm, y);
  }
  __shared__ int wmax[SORT_WARPS];
  if (lane == 31) wmax[w] = incm;
  __syncthreads();
  if (w == 0) {
    int x = wmax[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x = max(x, y);
    }
    wmax[lane] = x;
  }
  __syncthreads();
  int carry = (w > 0) ? wmax[w - 1] : -1;
  int prev = __shfl_up_sync(0xffffffffu, incm, 1);
  if (lane > 0) carry = max(carry, prev);
  uint32_t* ss = sstart + base;
  uint32_t* sl = slen + base;
  int cur = carry;
  for (int j = t_lo; j < t_hi; ++j) {
    if (j == 0 || kin[j - 1] != kin[j]) cur = j;
    ss[j] = (uint32_t)cur;
  }
  __syncthreads();
  // 5. segment lengths at the heads (written by the last entry of each segment)
  for (int j = t_lo; j < t_hi; ++j)
    if (j == M - 1 || kin[j + 1] != kin[j]) sl[ss[j]] = (uint32_t)(j + 1) - ss[j];
}

// ---------------------------------------------------------------------------------------
// forward only: sigma(score)  (evaluate_triples, holE.py:179-202)
// ---------------------------------------------------------------------------------------
template <int GS, int V>
__global__ void __launch_bounds__(256)
hole_score_kernel(const float* __restrict__ E, const int32_t* __restrict__ triples, int64_t B,
                  int nvec, int stride, float* __restrict__ out) {
  const int lane = threadIdx.x % GS;
  int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / GS;
  const bool valid = g < B;
  const int64_t i = valid ? g : B - 1;
  const int h = triples[3 * i], t = triples[3 * i + 1], r = triples[3 * i + 2];
  Row<V> xh, xt, xr;
  row_load<GS, V, true>(xh, E + (size_t)h * stride, lane, nvec);
  row_load<GS, V, true>(xt, E + (size_t)t * stride, lane, nvec);
  row_load<GS, V, true>(xr, E + (size_t)r * stride, lane, nvec);
  float ih = __frsqrt_rn(group_sum<GS>(row_sumsq(xh)));
  float it = __frsqrt_rn(group_sum<GS>(row_sumsq(xt)));
  float ir = __frsqrt_rn(group_sum<GS>(row_sumsq(xr)));
  row_clip(xh, ih); row_clip(xt, it); row_clip(xr, ir);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 4 * V; ++k) {
    float a = xh.re[k], b = xh.im[k], c = xr.re[k], d = xr.im[k], e = xt.re[k], f = xt.im[k];
    s += (a * c - b * d) * e + (a * d + b * c) * f;
  }
  s = group_sum<GS>(s);
  if (valid && lane == 0) out[i] = sigmoidf_precise(s);
}

// ---------------------------------------------------------------------------------------
// K1: fused forward + backward of one batch.  One group of GS lanes per positive triple.
// Writes the four merged gradient rows of triple i to G[slot*B + i], slots
// [relation, tail-slot, head-slot, corrupt entity] (the rows shared by the positive and
// the negative triple get the sum of both contributions; the clip backward is linear in
// the incoming gradient, so merging before it is exact in real arithmetic).
// ---------------------------------------------------------------------------------------
template <int GS, int V>
__global__ void __launch_bounds__(256)
hole_train_fwd_bwd_kernel(const float* __restrict__ E, const int32_t* __restrict__ pos,
                          const int32_t* __restrict__ neg, int side, int64_t B, int nvec,
                          int stride, float margin, float* __restrict__ G,
                          float* __restrict__ loss, float* __restrict__ sigma) {
  const int lane = threadIdx.x % GS;
  int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / GS;
  const bool valid = g < B;
  const int64_t i = valid ? g : B - 1;
  const int h = pos[3 * i], t = pos[3 * i + 1], r = pos[3 * i + 2], n = neg[i];

  Row<V> yh, yt, yr, yn;
  row_load<GS, V, true>(yh, E + (size_t)h * stride, lane, nvec);
  row_load<GS, V, true>(yt, E + (size_t)t * stride, lane, nvec);
  row_load<GS, V, true>(yr, E + (size_t)r * stride, lane, nvec);
  row_load<GS, V, true>(yn, E + (size_t)n * stride, lane, nvec);

  float ssh = row_sumsq(yh), sst = row_sumsq(yt), ssr = row_sumsq(yr), ssn = row_sumsq(yn);
#pragma unroll
  for (int o = GS / 2; o > 0; o >>= 1) {
    ssh += __shfl_xor_sync(0xffffffffu, ssh, o);
    sst += __shfl_xor_sync(0xffffffffu, sst, o);
    ssr += __shfl_xor_sync(0xffffffffu, ssr, o);
    ssn += __shfl_xor_sync(0xffffffffu, ssn, o);
  }
  const float ih = __frsqrt_rn(ssh), it = __frsqrt_rn(sst), ir = __frsqrt_rn(ssr),
              in_ = __frsqrt_rn(ssn);
  row_clip(yh, ih); row_clip(yt, it); row_clip(yr, ir); row_clip(yn, in_);

  // scores: s = sum (a c - b d) e + (a d + b c) f   with h=(a,b) r=(c,d) t=(e,f)
  float sp = 0.f, sn = 0.f;
#pragma unroll
  for (int k = 0; k < 4 * V; ++k) {
    const float a = yh.re[k], b = yh.im[k], c = yr.re[k], d = yr.im[k], e = yt.re[k], f = yt.im[k];
    const float pr = a * c - b * d, pi = a * d + b * c;
    sp += pr * e + pi * f;
    if (side) {   // negative = (n, t, r)
      const float a2 = yn.re[k], b2 = yn.im[k];
      sn += (a2 * c - b2 * d) * e + (a2 * d + b2 * c) * f;
    } else {      // negative = (h, n, r)
      sn += pr * yn.re[k] + pi * yn.im[k];
    }
  }
#pragma unroll
  for (int o = GS / 2; o > 0; o >>= 1) {
    sp += __shfl_xor_sync(0xffffffffu, sp, o);
    sn += __shfl_xor_sync(0xffffffffu, sn, o);
  }
  const float vp = sigmoidf_precise(sp), vn = sigmoidf_precise(sn);
  const float pre = vp - vn + margin;
  const bool act = pre >= 0.0f;                       // TF Maximum grad: GreaterEqual
  const float gp = act ? vp * (1.0f - vp) : 0.0f;
  const float gn = act ? -(vn * (1.0f - vn)) : 0.0f;
  if (valid && lane == 0) {
    loss[i] = fmaxf(pre, 0.0f);
    if (sigma != nullptr) { sigma[i] = vp; sigma[B + i] = vn; }
  }

  // gradients w.r.t. the clipped rows
  Row<V> dh, dt, dr, dn;
#pragma unroll
  for (int k = 0; k < 4 * V; ++k) {
    const float a = yh.re[k], b = yh.im[k], c = yr.re[k], d = yr.im[k], e = yt.re[k], f = yt.im[k];
    const float nr = yn.re[k], ni = yn.im[k];
    const float p_re = a * c - b * d, p_im = a * d + b * c;     // h * r        -> d/dt
    const float q_re = c * e + d * f, q_im = c * f - d * e;     // r * conj(t)~ -> d/dh
    const float u_re = a * e + b * f, u_im = a * f - b * e;     //              -> d/dr
    if (side) {   // negative (n, t, r): t and r shared, n plays the head
      const float p2_re = nr * c - ni * d, p2_im = nr * d + ni * c;
      const float u2_re = nr * e + ni * f, u2_im = nr * f - ni * e;
      dh.re[k] = gp * q_re;               dh.im[k] = gp * q_im;
      dn.re[k] = gn * q_re;               dn.im[k] = gn * q_im;
      dt.re[k] = gp * p_re + gn * p2_re;  dt.im[k] = gp * p_im + gn * p2_im;
      dr.re[k] = gp * u_re + gn * u2_re;  dr.im[k] = gp * u_im + gn * u2_im;
    } else {      // negative (h, n, r): h and r shared, n plays the tail
      const float q2_re = c * nr + d * ni, q2_im = c * ni - d * nr;
      const float u2_re = a * nr + b * ni, u2_im = a * ni - b * nr;
      dt.re[k] = gp * p_re;               dt.im[k] = gp * p_im;
      dn.re[k] = gn * p_re;               dn.im[k] = gn * p_im;
      dh.re[k] = gp * q_re + gn * q2_re;  dh.im[k] = gp * q_im + gn * q2_im;
      dr.re[k] = gp * u_re + gn * u2_re;  dr.im[k] = gp * u_im + gn * u2_im;
    }
  }
  // through the norm clip
  float ph = row_dot(yh, dh), pt = row_dot(yt, dt), pr_ = row_dot(yr, dr), pn = row_dot(yn, dn);
#pragma unroll
  for (int o = GS / 2; o > 0; o >>= 1) {
    ph += __shfl_xor_sync(0xffffffffu, ph, o);
    pt += __shfl_xor_sync(0xffffffffu, pt, o);
    pr_ += __shfl_xor_sync(0xffffffffu, pr_, o);
    pn += __shfl_xor_sync(0xffffffffu, pn, o);
  }
  clip_backward(dh, yh, ph, ih);
  clip_backward(dt, yt, pt, it);
  clip_backward(dr, yr, pr_, ir);
  clip_backward(dn, yn, pn, in_);
  if (valid) {
    row_store<GS, V>(dr, G + (size_t)(0 * B + i) * stride, lane, nvec);
    row_store<GS, V>(dt, G + (size_t)(1 * B + i) * stride, lane, nvec);
    row_store<GS, V>(dh, G + (size_t)(2 * B + i) * stride, lane, nvec);
    row_store<GS, V>(dn, G + (size_t)(3 * B + i) * stride, lane, nvec);
  }
}

// ---------------------------------------------------------------------------------------
// K3: deterministic sparse SGD update.  One group per sorted entry; only entries that head
// a chunk of C consecutive occurrences of a row do work.  A row with n occurrences is
// reduced by a fixed C-ary tree over its sorted occurrences (order: slot, then batch
// index); the last group to finish a node's children combines them in child order, so the
// result does not depend on scheduling.  Partials reuse the (already consumed) staged
// gradient row of a node's first occurrence.
// ---------------------------------------------------------------------------------------
template <int V>
__device__ __forceinline__ void row_zero(Row<V>& x) {
#pragma unroll
  for (int k = 0; k < 4 * V; ++k) { x.re[k] = 0.f; x.im[k] = 0.f; }
}
template <int V>
__device__ __forceinline__ void row_add(Row<V>& acc, const Row<V>& x) {
#pragma unroll
  for (int k = 0; k < 4 * V; ++k) { acc.re[k] += x.re[k]; acc.im[k] += x.im[k]; }
}

template <int GS, int V>
__global__ void __launch_bounds__(256)
hole_apply_kernel(float* __restrict__ E, float* __restrict__ G, const uint32_t* __restrict__ skey,
                  const uint32_t* __restrict__ spos, const uint32_t* __restrict__ sstart,
                  const uint32_t* __restrict__ slen, int* __restrict__ counters, int M, int nvec,
                  int stride, float lr) {
  constexpr int C = HOLE_TREE_C;
  const int lane = threadIdx.x % GS;
  const unsigned gmask = (GS == 32) ? 0xffffffffu
                                    : (((1u << GS) - 1u) << ((threadIdx.x % 32) / GS * GS));
  const int64_t j64 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / GS;
  if (j64 >= M) return;
  const int j = (int)j64;
  const int s = (int)sstart[j];
  const int rel = j - s;
  if (rel % C != 0) return;
  const int row = (int)skey[j];
  const int n = (int)slen[s];
  const int cnt = min(C, n - rel);

  Row<V> acc, x;
  row_zero(acc);
  for (int q = 0; q < cnt; ++q) {
    const uint32_t p = spos[j + q];
    row_load<GS, V, false>(x, G + (size_t)p * stride, lane, nvec);
    row_add(acc, x);
  }
  int level = 0, idx = rel / C, nl = (n + C - 1) / C;
  int64_t span = C;   // sorted entries covered by one node of this level
  while (true) {
    if (nl == 1) {    // root: apply   E[row] -= lr * acc   (holE.py:296)
      float* erow = E + (size_t)row * stride;
      row_load<GS, V, false>(x, erow, lane, nvec);
#pragma unroll
      for (int k = 0; k < 4 * V; ++k) {
        x.re[k] -= lr * acc.re[k];
        x.im[k] -= lr * acc.im[k];
      }
      row_store<GS, V>(x, erow, lane, nvec);
      return;
    }
    row_store<GS, V>(acc, G + (size_t)spos[s + idx * span] * stride, lane, nvec);
    __threadfence();
    __syncwarp(gmask);
    const int parent = idx / C;
    const int nchild = min(C, nl - parent * C);
    int* ctr = counters + (size_t)level * M + (s + parent * span * C);
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(ctr, 1);
    ticket = __shfl_sync(gmask, ticket, (threadIdx.x % 32) / GS * GS);
    if (ticket != nchild - 1) return;
    if (lane == 0) *ctr = 0;          // self-reset for the next step
    __threadfence();
    row_zero(acc);
    for (int c = 0; c < nchild; ++c) {
      const uint32_t p = spos[s + (parent * (int64_t)C + c) * span];
      row_load<GS, V, false>(x, G + (size_t)p * stride, lane, nvec);
      row_add(acc, x);
    }
    ++level;
    idx = parent;
    nl = (nl + C - 1) / C;
    span *= C;
  }
}

// deterministic per-step loss sum: one CTA per step, fixed-shape tree
__global__ void __launch_bounds__(256)
hole_loss_sum_kernel(const float* __restrict__ loss, int64_t B, float* __restrict__ out) {
  __shared__ float sm[256];
  const float* l = loss + (size_t)blockIdx.x * B;
  float a = 0.f;
  for (int64_t i = threadIdx.x; i < B; i += 256) a += l[i];
  sm[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
#define HOLE_DISPATCH(ctx, KERNEL, grid, block, stream, ...)                                 \
  do {                                                                                       \
    if ((ctx)->gs == 8 && (ctx)->v == 1) KERNEL<8, 1><<<grid, block, 0, stream>>>(__VA_ARGS__);          \
    else if ((ctx)->gs == 16 && (ctx)->v == 1) KERNEL<16, 1><<<grid, block, 0, stream>>>(__VA_ARGS__);   \
    else if ((ctx)->gs == 32 && (ctx)->v == 1) KERNEL<32, 1><<<grid, block, 0, stream>>>(__VA_ARGS__);   \
    else if ((ctx)->gs == 32 && (ctx)->v == 2) KERNEL<32, 2><<<grid, block, 0, stream>>>(__VA_ARGS__);   \
    else if ((ctx)->gs == 32 && (ctx)->v == 3) KERNEL<32, 3><<<grid, block, 0, stream>>>(__VA_ARGS__);   \
    else return hole_set_error(HOLE_ERR_UNSUPPORTED, "no kernel variant for gs=%d v=%d",     \
                               (ctx)->gs, (ctx)->v);                                         \
    HOLE_LAUNCHED();                                                                         \
  } while (0)

static inline unsigned grid_for_groups(int64_t groups, int gs, int block = 256) {
  int64_t per_block = block / gs;
  return (unsigned)std::max<int64_t>(1, (groups + per_block - 1) / per_block);
}

extern "C" int hole_abi_version(void) { return HOLE_ABI_VERSION; }
extern "C" const char* hole_last_error(void) { return g_hole_err.c_str(); }
extern "C" int64_t hole_launch_count(void) { return g_hole_launches; }
extern "C" void hole_launch_count_reset(void) { g_hole_launches = 0; }

extern "C" int hole_row_stride(int dim) {
  if (dim <= 0 || (dim & 1)) return HOLE_ERR_ARG;
  return 2 * (((dim / 2) + 3) / 4 * 4);
}

extern "C" int hole_ctx_create(hole_ctx** out, int device, int64_t n_rows, int dim) {
  HOLE_CHECK_ARG(out != nullptr);
  *out = nullptr;
  HOLE_CHECK_ARG(n_rows > 0 && n_rows < (int64_t(1) << 31));
  if (dim <= 0 || (dim & 1))   // holE.py:164-165 splits the row at dim//2
    return hole_set_error(HOLE_ERR_ARG, "embedding_dim must be a positive even number, got %d", dim);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return hole_set_error(HOLE_ERR_CUDA, "no CUDA device available (%s); libhole_b200 has no CPU fallback",
                          cudaGetErrorString(e));
  HOLE_CHECK_ARG(device >= 0 && device < ndev);
  cudaDeviceProp prop;
  HOLE_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return hole_set_error(HOLE_ERR_CUDA, "device %d is sm_%d%d; libhole_b200 is built for sm_100a only",
                          device, prop.major, prop.minor);
  HOLE_CUDA_TRY(cudaSetDevice(device));
  hole_ctx* c = new hole_ctx();
  c->device = device;
  c->n_rows = n_rows;
  c->dim = dim;
  c->H = dim / 2;
  c->nvec = (c->H + 3) / 4;
  c->row_stride = 8 * c->nvec;
  c->sm_count = prop.multiProcessorCount;
  if (c->nvec <= 8) { c->gs = 8; c->v = 1; }
  else if (c->nvec <= 16) { c->gs = 16; c->v = 1; }
  else { c->gs = 32; c->v = (c->nvec + 31) / 32; }
  if (c->v > 3) {
    delete c;
    return hole_set_error(HOLE_ERR_UNSUPPORTED, "embedding_dim %d > 768 not built", dim);
  }
  int bits = 1;
  while ((int64_t(1) << bits) < n_rows) ++bits;
  c->key_bits = bits;
  HOLE_CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 2; ++k) {
    HOLE_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_copy[k], cudaEventDisableTiming));
    HOLE_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming));
  }
  *out = c;
  return HOLE_OK;
}

static void ws_free(hole_ctx* c) {
  cudaFree(c->G); cudaFree(c->keysA); cudaFree(c->keysB); cudaFree(c->valsA); cudaFree(c->valsB);
  cudaFree(c->sstart); cudaFree(c->slen); cudaFree(c->counters); cudaFree(c->neg);
  cudaFree(c->loss); cudaFree(c->loss_sum);
  c->G = nullptr; c->keysA = c->keysB = c->valsA = c->valsB = c->sstart = c->slen = nullptr;
  c->counters = nullptr; c->neg = nullptr; c->loss = nullptr; c->loss_sum = nullptr;
  c->cap_B = c->cap_S = 0;
}

extern "C" int hole_ctx_destroy(hole_ctx* c) {
  if (c == nullptr) return HOLE_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  ws_free(c);
  hole_rank_ws_free(c);
  cudaFree(c->triples_stage[0]); cudaFree(c->triples_stage[1]);
  if (c->loss_sum_pinned) cudaFreeHost(c->loss_sum_pinned);
  for (int k = 0; k < 2; ++k) {
    if (c->ev_copy[k]) cudaEventDestroy(c->ev_copy[k]);
    if (c->ev_done[k]) cudaEventDestroy(c->ev_done[k]);
  }
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  delete c;
  return HOLE_OK;
}

// Workspace for S steps of batch B.  Reallocation synchronises the device.
int hole_ws_reserve(hole_ctx* c, int64_t B, int64_t S) {
  if (B <= c->cap_B && S <= c->cap_S) return HOLE_OK;
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  HOLE_CUDA_TRY(cudaDeviceSynchronize());
  B = std::max(B, c->cap_B);
  S = std::max(S, c->cap_S);
  ws_free(c);
  const size_t M = (size_t)4 * B;
#define WS_ALLOC(ptr, bytes)                                                             \
  do {                                                                                   \
    if (cudaMalloc((void**)&(ptr), (bytes)) != cudaSuccess) {                            \
      cudaGetLastError();                                                                \
      ws_free(c);                                                                        \
      return hole_set_error(HOLE_ERR_ALLOC, "workspace allocation of %zu bytes failed", (size_t)(bytes)); \
    }                                                                                    \
  } while (0)
  WS_ALLOC(c->G, M * c->row_stride * sizeof(float));
  WS_ALLOC(c->keysA, S * M * 4); WS_ALLOC(c->keysB, S * M * 4);
  WS_ALLOC(c->valsA, S * M * 4); WS_ALLOC(c->valsB, S * M * 4);
  WS_ALLOC(c->sstart, S * M * 4); WS_ALLOC(c->slen, S * M * 4);
  WS_ALLOC(c->counters, (size_t)HOLE_TREE_LEVELS * M * 4);
  WS_ALLOC(c->neg, (size_t)S * B * 4);
  WS_ALLOC(c->loss, (size_t)S * B * 4);
  WS_ALLOC(c->loss_sum, (size_t)S * 4);
#undef WS_ALLOC
  HOLE_CUDA_TRY(cudaMemset(c->counters, 0, (size_t)HOLE_TREE_LEVELS * M * 4));
  c->cap_B = B;
  c->cap_S = S;
  return HOLE_OK;
}

extern "C" int hole_pack_rows(hole_ctx* c, const float* src, float* dst, int64_t n, void* stream) {
  HOLE_CHECK_ARG(c && src && dst && n >= 0);
  if (n == 0) return HOLE_OK;
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  int64_t total = n * c->row_stride;
  unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)c->sm_count * 32);
  hole_pack_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, n, c->dim, c->H, c->row_stride / 2);
  HOLE_LAUNCHED();
  return HOLE_OK;
}

extern "C" int hole_unpack_rows(hole_ctx* c, const float* src, float* dst, int64_t n, void* stream) {
  HOLE_CHECK_ARG(c && src && dst && n >= 0);
  if (n == 0) return HOLE_OK;
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  int64_t total = n * c->dim;
  unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)c->sm_count * 32);
  hole_unpack_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, n, c->dim, c->H, c->row_stride / 2);
  HOLE_LAUNCHED();
  return HOLE_OK;
}

extern "C" int hole_corrupt(hole_ctx* c, const int32_t* triples, int64_t B, const int32_t* type_of,
                            const int64_t* csr_off, const int32_t* csr_ids, uint64_t seed,
                            uint64_t step, int32_t* side_out, int32_t* neg_out, int* side_host,
                            void* stream) {
  HOLE_CHECK_ARG(c && B >= 0);
  if (side_host) *side_host = hole_side_coin(seed, step);
  if (B == 0) return HOLE_OK;
  HOLE_CHECK_ARG(triples && type_of && csr_off && csr_ids && neg_out);
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  dim3 grid((unsigned)std::min<int64_t>((B + 255) / 256, 65535), 1);
  hole_corrupt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(triples, B, 1, type_of, csr_off, csr_ids,
                                                              seed, step, neg_out, side_out, nullptr);
  HOLE_LAUNCHED();
  return HOLE_OK;
}

extern "C" int hole_score(hole_ctx* c, const float* table, const int32_t* triples, int64_t B,
                          float* out_sigma, void* stream) {
  HOLE_CHECK_ARG(c && B >= 0);
  if (B == 0) return HOLE_OK;
  HOLE_CHECK_ARG(table && triples && out_sigma);
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  HOLE_DISPATCH(c, hole_score_kernel, grid_for_groups(B, c->gs), 256, (cudaStream_t)stream, table,
                triples, B, c->nvec, c->row_stride, out_sigma);
  return HOLE_OK;
}

static int sort_passes(const hole_ctx* c) {
  int p = (c->key_bits + 7) / 8;
  return (p + 1) & ~1;   // even, so the sorted pairs end in (keysA, valsA)
}

// K1 + K3 of one step whose plan (sorted keys, segments) is at plan slot `slot`.
static int run_step(hole_ctx* c, float* table, const int32_t* pos, const int32_t* neg, int side,
                    int64_t B, float margin, float lr, float* loss_out, float* sigma_out,
                    int64_t slot, cudaStream_t st) {
  const int M = (int)(4 * B);
  HOLE_DISPATCH(c, hole_train_fwd_bwd_kernel, grid_for_groups(B, c->gs), 256, st, table, pos, neg,
                side, B, c->nvec, c->row_stride, margin, c->G, loss_out, sigma_out);
  const size_t off = (size_t)slot * M;
  HOLE_DISPATCH(c, hole_apply_kernel, grid_for_groups(M, c->gs), 256, st, table, c->G,
                c->keysA + off, c->valsA + off, c->sstart + off, c->slen + off, c->counters, M,
                c->nvec, c->row_stride, lr);
  return HOLE_OK;
}

extern "C" int hole_train_step(hole_ctx* c, float* table, const int32_t* pos, const int32_t* neg_ent,
                               int side, int64_t B, float margin, float lr, float* loss_out,
                               float* sigma_out, void* stream) {
  HOLE_CHECK_ARG(c && B >= 0 && (side == 0 || side == 1));
  if (B == 0) return HOLE_OK;   // an empty batch is a no-op
  HOLE_CHECK_ARG(table && pos && neg_ent && loss_out);
  HOLE_CHECK_ARG(4 * B < (int64_t(1) << 31));
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  int rc = hole_ws_reserve(c, B, 1);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  hole_keys_kernel<<<(unsigned)std::min<int64_t>((B + 255) / 256, 65535), 256, 0, st>>>(pos, neg_ent, B, c->keysA);
  HOLE_LAUNCHED();
  hole_plan_sort_kernel<<<1, SORT_THREADS, 0, st>>>(c->keysA, c->valsA, c->keysB, c->valsB, c->sstart,
                                                   c->slen, (int)(4 * B), sort_passes(c));
  HOLE_LAUNCHED();
  return run_step(c, table, pos, neg_ent, side, B, margin, lr, loss_out, sigma_out, 0, st);
}

// chunk size (steps planned at once): enough CTAs for the sort to fill the GPU, bounded
// workspace
static int64_t plan_chunk(const hole_ctx* c, int64_t B, int64_t n_steps) {
  int64_t by_mem = std::max<int64_t>(1, (int64_t(1) << 26) / (4 * B));   // <= 64M plan entries
  return std::max<int64_t>(1, std::min<int64_t>({n_steps, (int64_t)2 * c->sm_count, by_mem}));
}

static int train_chunk(hole_ctx* c, float* table, const int32_t* triples_dev, int64_t B, int64_t S,
                       const int32_t* type_of, const int64_t* csr_off, const int32_t* csr_ids,
                       uint64_t seed, uint64_t first_step, float margin, const float* lr_host,
                       float* loss_out /*[S*B] device*/, float* loss_sum_dev /*[S] device or null*/,
                       cudaStream_t st) {
  const int M = (int)(4 * B);
  dim3 grid((unsigned)std::min<int64_t>((B + 255) / 256, 4096), (unsigned)S);
  hole_corrupt_kernel<<<grid, 256, 0, st>>>(triples_dev, B, (int)S, type_of, csr_off, csr_ids, seed,
                                            first_step, c->neg, nullptr, c->keysA);
  HOLE_LAUNCHED();
  hole_plan_sort_kernel<<<(unsigned)S, SORT_THREADS, 0, st>>>(c->keysA, c->valsA, c->keysB, c->valsB,
                                                              c->sstart, c->slen, M, sort_passes(c));
  HOLE_LAUNCHED();
  for (int64_t k = 0; k < S; ++k) {
    int side = hole_side_coin(seed, first_step + (uint64_t)k);
    int rc = run_step(c, table, triples_dev + (size_t)k * B * 3, c->neg + (size_t)k * B, side, B,
                      margin, lr_host[k], loss_out + (size_t)k * B, nullptr, k, st);
    if (rc) return rc;
  }
  if (loss_sum_dev != nullptr) {
    hole_loss_sum_kernel<<<(unsigned)S, 256, 0, st>>>(loss_out, B, loss_sum_dev);
    HOLE_LAUNCHED();
  }
  return HOLE_OK;
}

extern "C" int hole_train_steps(hole_ctx* c, float* table, const int32_t* triples, int64_t B,
                                int64_t n_steps, const int32_t* type_of, const int64_t* csr_off,
                                const int32_t* csr_ids, uint64_t seed, uint64_t first_step,
                                float margin, const float* lr, float* loss_out, float* loss_sum_out,
                                void* stream) {
  HOLE_CHECK_ARG(c && B >= 0 && n_steps >= 0);
  if (B == 0 || n_steps == 0) return HOLE_OK;
  HOLE_CHECK_ARG(table && triples && type_of && csr_off && csr_ids && lr);
  HOLE_CHECK_ARG(4 * B < (int64_t(1) << 31));
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  const int64_t S = plan_chunk(c, B, n_steps);
  int rc = hole_ws_reserve(c, B, S);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t k0 = 0; k0 < n_steps; k0 += S) {
    const int64_t s = std::min(S, n_steps - k0);
    float* lo = loss_out ? loss_out + (size_t)k0 * B : c->loss;
    rc = train_chunk(c, table, triples + (size_t)k0 * B * 3, B, s, type_of, csr_off, csr_ids, seed,
                     first_step + (uint64_t)k0, margin, lr + k0, lo,
                     loss_sum_out ? loss_sum_out + k0 : nullptr, st);
    if (rc) return rc;
  }
  return HOLE_OK;
}

extern "C" int hole_train_steps_host(hole_ctx* c, float* table, const int32_t* triples_host, int64_t B,
                                     int64_t n_steps, const int32_t* type_of, const int64_t* csr_off,
                                     const int32_t* csr_ids, uint64_t seed, uint64_t first_step,
                                     float margin, const float* lr, float* loss_sum_host,
                                     void* stream) {
  HOLE_CHECK_ARG(c && B >= 0 && n_steps >= 0);
  if (B == 0 || n_steps == 0) return HOLE_OK;
  HOLE_CHECK_ARG(table && triples_host && type_of && csr_off && csr_ids && lr && loss_sum_host);
  HOLE_CHECK_ARG(4 * B < (int64_t(1) << 31));
  HOLE_CUDA_TRY(cudaSetDevice(c->device));
  const int64_t S = plan_chunk(c, B, n_steps);
  int rc = hole_ws_reserve(c, B, S);
  if (rc) return rc;
  const int64_t stage_elems = S * B * 3;
  if (stage_elems > c->cap_stage) {
    HOLE_CUDA_TRY(cudaDeviceSynchronize());
    for (int k = 0; k < 2; ++k) {
      cudaFree(c->triples_stage[k]);
      c->triples_stage[k] = nullptr;
      if (cudaMalloc((void**)&c->triples_stage[k], stage_elems * 4) != cudaSuccess) {
        cudaGetLastError();
        c->cap_stage = 0;
        return hole_set_error(HOLE_ERR_ALLOC, "triple staging allocation failed");
      }
    }
    c->cap_stage = stage_elems;
  }
  if (n_steps > c->cap_pinned) {
    if (c->loss_sum_pinned) cudaFreeHost(c->loss_sum_pinned);
    c->loss_sum_pinned = nullptr;
    HOLE_CUDA_TRY(cudaMallocHost((void**)&c->loss_sum_pinned, n_steps * sizeof(float)));
    c->cap_pinned = n_steps;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // copy stream must not overtake work already queued on the compute stream that still
  // reads a staging buffer: ev_done[b] guards buffer b.
  int64_t nchunks = (n_steps + S - 1) / S;
  for (int64_t ci = 0; ci < nchunks; ++ci) {
    const int b = (int)(ci & 1);
    const int64_t k0 = ci * S, s = std::min(S, n_steps - k0);
    if (ci >= 2) HOLE_CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_done[b], 0));
    HOLE_CUDA_TRY(cudaMemcpyAsync(c->triples_stage[b], triples_host + (size_t)k0 * B * 3,
                                  (size_t)s * B * 3 * 4, cudaMemcpyHostToDevice, c->copy_stream));
    HOLE_CUDA_TRY(cudaEventRecord(c->ev_copy[b], c->copy_stream));
    HOLE_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_copy[b], 0));
    rc = train_chunk(c, table, c->triples_stage[b], B, s, type_of, csr_off, csr_ids, seed,
                     first_step + (uint64_t)k0, margin, lr + k0, c->loss, c->loss_sum, st);
    if (rc) return rc;
    HOLE_CUDA_TRY(cudaMemcpyAsync(c->loss_sum_pinned + k0, c->loss_sum, (size_t)s * 4,
                                  cudaMemcpyDeviceToHost, st));
    HOLE_CUDA_TRY(cudaEventRecord(c->ev_done[b], st));
  }
  HOLE_CUDA_TRY(cudaStreamSynchronize(st));
  for (int64_t k = 0; k < n_steps; ++k) loss_sum_host[k] = c->loss_sum_pinned[k];
  return HOLE_OK;
}
