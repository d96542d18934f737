// The ARCHIVED score variant of the reference (run directory holE-20170724), as an optional score mode
// (SURVEY.md section 8f item 4): value = tanh(s), s = sum_k Re(m_k) + Im(m_k), m = r * c,
// c = ifft(conj(fft(h)) * fft(t)), i.e. c_k = sum_j conj(h_j) t_{(j+k) mod H}
// [holE-20170724/graph.pbtxt:6221-6521]; loss = max(tanh(s+) - tanh(s-) + margin, 0) [15874-15988].
// h, r, t are the clipped rows as complex vectors, exactly as in the live model.
//
// The correlations are computed directly (O(H^2) per triple, fp32 FMAs on CUDA cores), one warp per
// triple.  With rho = (1 - i) r:  s = Re sum_k rho_k c_k, and
//   d s / d h_j = u_j,        u_j = sum_k rho_k t_{j+k}
//   d s / d t_m = conj(w_m),  w_m = sum_k rho_k conj(h_{m-k})
//   d s / d r_k = [Re c_k + Im c_k ; Re c_k - Im c_k]
// The positive and the negative triple share two of their three rows, so u (or w) of both is ONE
// correlation over the combined vector g+ y + g- y': four correlations per triple in all.
// Row updates follow K1's protocol (plan slots: unique rows in place, others staged in G for K3).
#pragma once

#include "hole_ccorr_core.cuh"

// dy (gradient w.r.t. the clipped row y, held in shared memory) -> through the clip -> in place / staged
template <int NV>
__device__ __forceinline__ void cc_emit(CVec<NV>& dy, const float* __restrict__ y_re,
                                        const float* __restrict__ y_im, float inv, bool act, float lr,
                                        float* __restrict__ erow, float* __restrict__ grow, bool uniq,
                                        int H, int Hp, int lane) {
  if (inv <= 1.0f) {                                       // warp-uniform: the row was clipped
    float proj = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int k = lane + 32 * v;
      if (k < H) proj = fmaf(y_re[k], dy.re[v], fmaf(y_im[k], dy.im[v], proj));
    }
    proj = cc_warp_sum(proj);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int k = lane + 32 * v;
      if (k < H) {
        dy.re[v] = (dy.re[v] - y_re[k] * proj) * inv;
        dy.im[v] = (dy.im[v] - y_im[k] * proj) * inv;
      }
    }
  }
  if (uniq) {
    if (act) {                                             // inactive hinge: zero gradient, row unchanged
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int k = lane + 32 * v;
        if (k < H) {
          erow[k] = erow[k] - lr * dy.re[v];
          erow[Hp + k] = erow[Hp + k] - lr * dy.im[v];
        }
      }
    }
  } else {
#pragma unroll
    for (int v = 0; v < NV; ++v) {                         // the padding of a staged row is summed by K3: zeros
      const int k = lane + 32 * v;
      if (k < Hp) {
        grow[k] = (k < H) ? dy.re[v] : 0.f;
        grow[Hp + k] = (k < H) ? dy.im[v] : 0.f;
      }
    }
  }
}

// floats of shared memory one warp needs: h, t, n doubled (12H), combined vector doubled (4H), rho (2H)
__host__ __device__ constexpr int cc_warp_floats(int H) { return 18 * H; }

template <int NV>
__global__ void __launch_bounds__(256)
hole_ccorr_fwd_bwd_kernel(float* __restrict__ E, const int32_t* __restrict__ tri, const int32_t* __restrict__ neg,
                          const int32_t* __restrict__ perm, const uint32_t* __restrict__ gslot, int side, int B,
                          int T, int H, int stride, float margin, float lr, float* __restrict__ G,
                          float* __restrict__ loss, float* __restrict__ sigma) {
  extern __shared__ float cc_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* sm = cc_smem + (size_t)wib * cc_warp_floats(H);
  float *h_re = sm, *h_im = sm + 2 * H, *t_re = sm + 4 * H, *t_im = sm + 6 * H, *n_re = sm + 8 * H,
        *n_im = sm + 10 * H, *c_re = sm + 12 * H, *c_im = sm + 14 * H, *rho_re = sm + 16 * H, *rho_im = sm + 17 * H;
  const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t g0l = w * T;
  if (g0l >= B) return;
  const int g0 = (int)g0l, g1 = min(B, g0 + T);
  const int Hp = stride / 2;
  pdl_wait();                        // the previous step's K3 has finished updating the table

  CVec<NV> yr, acc;                  // the run's relation row (clipped) and its summed gradient
  int r_cur = -1, run_i = 0;
  float ir = 0.f;
  bool run_act = false;

  auto flush_relation = [&]() {
    const uint32_t sl = gslot[run_i];
    float* yr_re = c_re;             // scratch: the combined-vector buffer is free between triples
    float* yr_im = c_im;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int k = lane + 32 * v;
      if (k < H) { yr_re[k] = yr.re[v]; yr_im[k] = yr.im[v]; }
    }
    __syncwarp();
    cc_emit<NV>(acc, yr_re, yr_im, ir, run_act, lr, E + (size_t)r_cur * stride, G + (size_t)sl * stride,
                sl == HOLE_SLOT_UNIQUE, H, Hp, lane);
    __syncwarp();
  };

  for (int g = g0; g < g1; ++g) {
    const int i = perm[g];
    const int h = tri[3 * i], t = tri[3 * i + 1], r = tri[3 * i + 2], n = neg[i];
    if (r != r_cur) {                // warp-uniform
      if (r_cur >= 0) flush_relation();
      ir = cc_load_row<NV>(E + (size_t)r * stride, H, Hp, lane, c_re, c_im, &yr);   // (c_* only as a landing pad)
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int k = lane + 32 * v;
        if (k < H) { rho_re[k] = yr.re[v] + yr.im[v]; rho_im[k] = yr.im[v] - yr.re[v]; }   // (1 - i)(c + i d)
        acc.re[v] = 0.f; acc.im[v] = 0.f;
      }
      r_cur = r; run_i = i; run_act = false;
    }
    const float ih = cc_load_row<NV>(E + (size_t)h * stride, H, Hp, lane, h_re, h_im, nullptr);
    const float it = cc_load_row<NV>(E + (size_t)t * stride, H, Hp, lane, t_re, t_im, nullptr);
    const float in_ = cc_load_row<NV>(E + (size_t)n * stride, H, Hp, lane, n_re, n_im, nullptr);
    __syncwarp();

    // ---- forward: c+ = ccorr(h, t); c- = ccorr(n, t) (heads corrupted) or ccorr(h, n)
    CVec<NV> cp, cn;
    cc_corr<NV, true, false, true>(cp, h_re, h_im, t_re, t_im, H, lane);
    if (side) cc_corr<NV, true, false, true>(cn, n_re, n_im, t_re, t_im, H, lane);
    else      cc_corr<NV, true, false, true>(cn, h_re, h_im, n_re, n_im, H, lane);
    float sp = 0.f, sn = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int k = lane + 32 * v;
      if (k < H) {                   // Re(rho_k c_k)
        sp += rho_re[k] * cp.re[v] - rho_im[k] * cp.im[v];
        sn += rho_re[k] * cn.re[v] - rho_im[k] * cn.im[v];
      }
    }
    sp = cc_warp_sum(sp);
    sn = cc_warp_sum(sn);
    const float vp = tanhf(sp), vn = tanhf(sn);
    const float pre = vp - vn + margin;
    const bool act = pre >= 0.0f;                          // TF Maximum grad: GreaterEqual
    const float gp = act ? 1.0f - vp * vp : 0.0f;
    const float gn = act ? -(1.0f - vn * vn) : 0.0f;
    if (lane == 0) {
      loss[i] = fmaxf(pre, 0.0f);
      if (sigma != nullptr) { sigma[i] = vp; sigma[B + i] = vn; }
    }
    run_act = run_act || act;
#pragma unroll
    for (int v = 0; v < NV; ++v) {   // d s / d r_k = [Re c + Im c ; Re c - Im c]
      acc.re[v] += gp * (cp.re[v] + cp.im[v]) + gn * (cn.re[v] + cn.im[v]);
      acc.im[v] += gp * (cp.re[v] - cp.im[v]) + gn * (cn.re[v] - cn.im[v]);
    }

    // ---- backward
    const uint32_t sl_t = gslot[B + i], sl_h = gslot[2 * B + i], sl_n = gslot[3 * B + i];
    // combined vector: tails corrupted -> g+ t + g- n (shares h); heads corrupted -> g+ h + g- n (shares t)
    {
      const float* a_re = side ? h_re : t_re;
      const float* a_im = side ? h_im : t_im;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int k = lane + 32 * v;
        if (k < H) {
          const float xr = gp * a_re[k] + gn * n_re[k], xi = gp * a_im[k] + gn * n_im[k];
          c_re[k] = xr; c_re[k + H] = xr;
          c_im[k] = xi; c_im[k + H] = xi;
        }
      }
    }
    __syncwarp();
    CVec<NV> u, wv, d;
    if (side == 0) {
      cc_corr<NV, false, false, true>(u, rho_re, rho_im, c_re, c_im, H, lane);      // u_j = sum_k rho_k (g+ t + g- n)_{j+k}
      cc_corr<NV, false, true, false>(wv, rho_re, rho_im, h_re, h_im, H, lane);     // w_m = sum_k rho_k conj(h_{m-k})
#pragma unroll
      for (int v = 0; v < NV; ++v) { d.re[v] = gp * wv.re[v]; d.im[v] = -gp * wv.im[v]; }
      cc_emit<NV>(d, t_re, t_im, it, act, lr, E + (size_t)t * stride, G + (size_t)sl_t * stride,
                  sl_t == HOLE_SLOT_UNIQUE, H, Hp, lane);
      cc_emit<NV>(u, h_re, h_im, ih, act, lr, E + (size_t)h * stride, G + (size_t)sl_h * stride,
                  sl_h == HOLE_SLOT_UNIQUE, H, Hp, lane);
#pragma unroll
      for (int v = 0; v < NV; ++v) { d.re[v] = gn * wv.re[v]; d.im[v] = -gn * wv.im[v]; }
      cc_emit<NV>(d, n_re, n_im, in_, act, lr, E + (size_t)n * stride, G + (size_t)sl_n * stride,
                  sl_n == HOLE_SLOT_UNIQUE, H, Hp, lane);
    } else {
      cc_corr<NV, false, false, true>(u, rho_re, rho_im, t_re, t_im, H, lane);      // u_j = sum_k rho_k t_{j+k}
      cc_corr<NV, false, true, false>(wv, rho_re, rho_im, c_re, c_im, H, lane);     // w_m = sum_k rho_k conj((g+ h + g- n)_{m-k})
#pragma unroll
      for (int v = 0; v < NV; ++v) { d.re[v] = wv.re[v]; d.im[v] = -wv.im[v]; }
      cc_emit<NV>(d, t_re, t_im, it, act, lr, E + (size_t)t * stride, G + (size_t)sl_t * stride,
                  sl_t == HOLE_SLOT_UNIQUE, H, Hp, lane);
#pragma unroll
      for (int v = 0; v < NV; ++v) { d.re[v] = gp * u.re[v]; d.im[v] = gp * u.im[v]; }
      cc_emit<NV>(d, h_re, h_im, ih, act, lr, E + (size_t)h * stride, G + (size_t)sl_h * stride,
                  sl_h == HOLE_SLOT_UNIQUE, H, Hp, lane);
#pragma unroll
      for (int v = 0; v < NV; ++v) { d.re[v] = gn * u.re[v]; d.im[v] = gn * u.im[v]; }
      cc_emit<NV>(d, n_re, n_im, in_, act, lr, E + (size_t)n * stride, G + (size_t)sl_n * stride,
                  sl_n == HOLE_SLOT_UNIQUE, H, Hp, lane);
    }
    __syncwarp();                    // the next triple overwrites the shared rows
  }
  pdl_launch_dependents();           // K3 of this step may start its prologue
  flush_relation();
}

// floats of shared memory one warp of the score kernel needs: h (2H), t doubled (4H), landing pad (4H)
__host__ __device__ constexpr int cc_score_warp_floats(int H) { return 10 * H; }

// forward only: out[i] = tanh(s_i)
template <int NV>
__global__ void __launch_bounds__(256)
hole_ccorr_score_kernel(const float* __restrict__ E, const int32_t* __restrict__ triples, int64_t B, int H,
                        int stride, float* __restrict__ out) {
  extern __shared__ float cc_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // per warp: h (2H, one copy: it is the broadcast operand), t doubled (4H), landing pad (4H)
  float* sm = cc_smem + (size_t)wib * cc_score_warp_floats(H);
  float *h_re = sm, *h_im = sm + H, *t_re = sm + 2 * H, *t_im = sm + 4 * H, *scratch = sm + 6 * H;
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (i >= B) return;
  const int Hp = stride / 2;
  const int h = triples[3 * i], t = triples[3 * i + 1], r = triples[3 * i + 2];
  CVec<NV> yr;
  {
    const float* p = E + (size_t)h * stride;
    float x_re[NV], x_im[NV], ss = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int k = lane + 32 * v;
      x_re[v] = (k < H) ? p[k] : 0.f;
      x_im[v] = (k < H) ? p[Hp + k] : 0.f;
      ss = fmaf(x_re[v], x_re[v], fmaf(x_im[v], x_im[v], ss));
    }
    const float sc = fminf(__frsqrt_rn(cc_warp_sum(ss)), 1.0f);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int k = lane + 32 * v;
      if (k < H) { h_re[k] = x_re[v] * sc; h_im[k] = x_im[v] * sc; }
    }
  }
  cc_load_row<NV>(E + (size_t)t * stride, H, Hp, lane, t_re, t_im, nullptr);
  cc_load_row<NV>(E + (size_t)r * stride, H, Hp, lane, scratch, scratch + 2 * H, &yr);   // kept in registers
  __syncwarp();
  CVec<NV> c;
  cc_corr<NV, true, false, true>(c, h_re, h_im, t_re, t_im, H, lane);
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int k = lane + 32 * v;
    if (k < H) s += (yr.re[v] + yr.im[v]) * c.re[v] - (yr.im[v] - yr.re[v]) * c.im[v];
  }
  s = cc_warp_sum(s);
  if (lane == 0) out[i] = tanhf(s);
}

