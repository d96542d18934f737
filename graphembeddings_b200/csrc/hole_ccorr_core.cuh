// Building blocks of the archived FFT / tanh score variant (hole_ccorr.cuh, and the ranking query packer
// of hole_rank.cu): direct complex circular correlations out of shared memory, one warp per vector.
#pragma once

template <int NV>
struct CVec { float re[NV], im[NV]; };

__device__ __forceinline__ float cc_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out_k = sum_{j<H} B_j (x) M2[k + j]        (FWD)      k = lane + 32 v
//       = sum_{j<H} B_j (x) M2[k - j + H]    (!FWD)
// M2 = the moving vector stored twice in a row (length 2H), so no index wraps;
// (x) = (b_re + i sb b_im)(m_re + i sm m_im) with sb / sm = -1 for a conjugated operand.
// Terms are added in j order (what oracle/hole_ccorr.py's ccorr_direct does).
template <int NV, bool CONJ_B, bool CONJ_M, bool FWD>
__device__ __forceinline__ void cc_corr(CVec<NV>& out, const float* __restrict__ b_re,
                                        const float* __restrict__ b_im, const float* __restrict__ m_re,
                                        const float* __restrict__ m_im, int H, int lane) {
  int k[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    out.re[v] = 0.f; out.im[v] = 0.f;
    k[v] = min(lane + 32 * v, H - 1) + (FWD ? 0 : H);       // lanes past H compute a value nobody uses
  }
#pragma unroll 4
  for (int j = 0; j < H; ++j) {
    const float br = b_re[j], bi = CONJ_B ? -b_im[j] : b_im[j];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int idx = FWD ? k[v] + j : k[v] - j;
      const float mr = m_re[idx], mi = CONJ_M ? -m_im[idx] : m_im[idx];
      out.re[v] = fmaf(br, mr, fmaf(-bi, mi, out.re[v]));
      out.im[v] = fmaf(br, mi, fmaf(bi, mr, out.im[v]));
    }
  }
}

// clipped row of the table -> the doubled shared-memory copy; returns rsqrt(sum x^2)
template <int NV>
__device__ __forceinline__ float cc_load_row(const float* __restrict__ p, int H, int Hp, int lane,
                                             float* __restrict__ d_re, float* __restrict__ d_im, CVec<NV>* keep) {
  CVec<NV> x;
  float ss = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int k = lane + 32 * v;
    x.re[v] = (k < H) ? p[k] : 0.f;
    x.im[v] = (k < H) ? p[Hp + k] : 0.f;
    ss = fmaf(x.re[v], x.re[v], fmaf(x.im[v], x.im[v], ss));
  }
  const float inv = __frsqrt_rn(cc_warp_sum(ss));
  const float sc = fminf(inv, 1.0f);                       // clip_by_norm(row, 1)  (App. B)
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int k = lane + 32 * v;
    x.re[v] *= sc; x.im[v] *= sc;
    if (k < H) {
      d_re[k] = x.re[v]; d_re[k + H] = x.re[v];
      d_im[k] = x.im[v]; d_im[k + H] = x.im[v];
    }
  }
  if (keep != nullptr) *keep = x;
  return inv;
}

