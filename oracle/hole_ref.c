/* hole_ref.c -- C/OpenMP restatement of holE.py's training step and ranking
 * (TEST INFRASTRUCTURE ONLY: the CPU baseline bench.py times, and a second opinion for
 * the NumPy oracle; never linked into the product).
 *
 * PARITY PINNED MODULO A TENSORFLOW SHIM: TensorFlow 1.2 is not installable here and the
 * reference ships no golden vectors, so the pin is tests/golden/tfshim_step.npz -- the outputs of
 * holE.py's own evaluate_batch / minimize source executed on tests/golden/tfshim.py -- which this
 * port matches to 5e-6 (tests/test_tfshim_golden.py); it is also checked against
 * oracle/hole_oracle.py in tests/test_oracle_c.py.  Cites: get_embedding holE.py:161-168,
 * score holE.py:191-192, sigmoid holE.py:198, hinge holE.py:231, SGD holE.py:296 with the
 * TF conventions of SURVEY.md App. B (clip formula, sum-of-losses seed, >= hinge/clip
 * ties, slice order [r+, r-, t+, t-, h+, h-], sequential ScatterSub).
 *
 * Table layout: float32 [N, D] = [Re | Im] (holE.py:164-165).  Triples (h, t, r).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int hole_ref_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm asks for the cores back */
void hole_ref_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* y = x * min(1/sqrt(sum x^2), 1); returns inv norm */
static inline float clip_row(const float* x, float* y, int D) {
  float ss = 0.f;
  for (int k = 0; k < D; ++k) ss += x[k] * x[k];
  float inv = 1.0f / sqrtf(ss);
  float sc = inv < 1.0f ? inv : 1.0f;
  for (int k = 0; k < D; ++k) y[k] = x[k] * sc;
  return inv;
}

static inline float score_rows(const float* h, const float* t, const float* r, int H) {
  float s = 0.f;
  for (int k = 0; k < H; ++k) {
    float a = h[k], b = h[H + k], c = r[k], d = r[H + k], e = t[k], f = t[H + k];
    s += (a * c - b * d) * e + (a * d + b * c) * f;
  }
  return s;
}

static inline float sigmoidf_(float s) { return 1.0f / (1.0f + expf(-s)); }

/* sigma(score) for B triples */
void hole_ref_score(const float* E, int D, const int32_t* tr, int64_t B, float* out) {
  const int H = D / 2;
#pragma omp parallel
  {
    float* buf = (float*)malloc(sizeof(float) * 3 * D);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < B; ++i) {
      clip_row(E + (size_t)tr[3 * i] * D, buf, D);
      clip_row(E + (size_t)tr[3 * i + 1] * D, buf + D, D);
      clip_row(E + (size_t)tr[3 * i + 2] * D, buf + 2 * D, D);
      out[i] = sigmoidf_(score_rows(buf, buf + D, buf + 2 * D, H));
    }
    free(buf);
  }
}

/* gradient of g * s wrt the three UNCLIPPED rows of one triple -> dh, dt, dr (each D) */
static void side_grads(const float* E, int D, int h, int t, int r, float g, float* y /*3D scratch*/,
                       float* dh, float* dt, float* dr) {
  const int H = D / 2;
  float inv[3];
  const int ids[3] = {h, t, r};
  for (int q = 0; q < 3; ++q) inv[q] = clip_row(E + (size_t)ids[q] * D, y + q * D, D);
  const float *yh = y, *yt = y + D, *yr = y + 2 * D;
  for (int k = 0; k < H; ++k) {
    float a = yh[k], b = yh[H + k], e = yt[k], f = yt[H + k], c = yr[k], d = yr[H + k];
    dh[k] = g * (c * e + d * f);  dh[H + k] = g * (c * f - d * e);
    dt[k] = g * (a * c - b * d);  dt[H + k] = g * (a * d + b * c);
    dr[k] = g * (a * e + b * f);  dr[H + k] = g * (a * f - b * e);
  }
  float* ds[3] = {dh, dt, dr};
  for (int q = 0; q < 3; ++q) {
    if (inv[q] <= 1.0f) {     /* clipped: dx = (dy - y (y.dy)) * inv */
      const float* yy = y + q * D;
      float proj = 0.f;
      for (int k = 0; k < D; ++k) proj += yy[k] * ds[q][k];
      for (int k = 0; k < D; ++k) ds[q][k] = (ds[q][k] - yy[k] * proj) * inv[q];
    }
  }
}

/* One training step in place.  G is scratch of 6*B*D floats, idx scratch of 6*B ints.
 * Returns the sum of the hinge losses; loss_out[B] may be NULL. */
double hole_ref_train_step(float* E, int64_t N, int D, const int32_t* pos, const int32_t* neg_ent,
                           int side, int64_t B, float margin, float lr, float* loss_out, float* G,
                           int32_t* idx) {
  (void)N;
  const int H = D / 2;
  double total = 0.0;
#pragma omp parallel
  {
    float* y = (float*)malloc(sizeof(float) * 3 * D);
#pragma omp for schedule(static) reduction(+ : total)
    for (int64_t i = 0; i < B; ++i) {
      int h = pos[3 * i], t = pos[3 * i + 1], r = pos[3 * i + 2];
      int h2 = side ? neg_ent[i] : h, t2 = side ? t : neg_ent[i];
      clip_row(E + (size_t)h * D, y, D); clip_row(E + (size_t)t * D, y + D, D);
      clip_row(E + (size_t)r * D, y + 2 * D, D);
      float vp = sigmoidf_(score_rows(y, y + D, y + 2 * D, H));
      clip_row(E + (size_t)h2 * D, y, D); clip_row(E + (size_t)t2 * D, y + D, D);
      float vn = sigmoidf_(score_rows(y, y + D, y + 2 * D, H));
      float pre = vp - vn + margin;
      float l = pre > 0.f ? pre : 0.f;
      if (loss_out) loss_out[i] = l;
      total += l;
      float act = pre >= 0.f ? 1.f : 0.f;
      float gp = act * vp * (1.f - vp), gn = -act * vn * (1.f - vn);
      /* slices [r+, r-, t+, t-, h+, h-] (App. B) */
      side_grads(E, D, h, t, r, gp, y, G + (size_t)(4 * B + i) * D, G + (size_t)(2 * B + i) * D,
                 G + (size_t)(0 * B + i) * D);
      side_grads(E, D, h2, t2, r, gn, y, G + (size_t)(5 * B + i) * D, G + (size_t)(3 * B + i) * D,
                 G + (size_t)(1 * B + i) * D);
      idx[0 * B + i] = r;  idx[1 * B + i] = r;
      idx[2 * B + i] = t;  idx[3 * B + i] = t2;
      idx[4 * B + i] = h;  idx[5 * B + i] = h2;
    }
    free(y);
    /* ScatterSub: every (index, row) pair applied in concat order; rows are partitioned
     * over threads so each row still sees its updates in the sequential order. */
#ifdef _OPENMP
    const int nt = omp_get_num_threads(), me = omp_get_thread_num();
#else
    const int nt = 1, me = 0;
#endif
    for (int64_t p = 0; p < 6 * B; ++p) {
      int row = idx[p];
      if (row % nt != me) continue;
      float* e = E + (size_t)row * D;
      const float* g = G + (size_t)p * D;
      for (int k = 0; k < D; ++k) e[k] -= g[k] * lr;
    }
  }
  return total;
}

/* One --log_loss training step in place (holE.py:194-196, 206-220, 296).
 * Loss rows: B positives (label +1), then k corrupt batches (label -1); batch j replaces the
 * head (sides[j] = 1) or the tail of every positive by neg_ent[j*B + i].  Row loss =
 * log(1 + exp(-label * s)); the L2 scalar  l2_loss = sum(E^2)/2  (tf.nn.l2_loss over the whole
 * variable) is returned through *l2_loss_out and is NOT added to loss_out.  Update:
 * E -= lr * (sum of the sparse score gradients taken at the old table + (1+k) B l2 E).
 * G: scratch (1+k)*3*B*D floats, idx: (1+k)*3*B ints.  loss_out: [(1+k)*B]. */
void hole_ref_logloss_step(float* E, int64_t N, int D, const int32_t* pos, const int32_t* neg_ent,
                           const int32_t* sides, int k, int64_t B, float lr, float l2, float* loss_out,
                           double* l2_loss_out, float* G, int32_t* idx) {
  const int H = D / 2;
  const int64_t terms = (int64_t)(1 + k) * B;
  double l2sum = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : l2sum)
  for (int64_t p = 0; p < N * (int64_t)D; ++p) l2sum += (double)E[p] * (double)E[p];
  if (l2_loss_out) *l2_loss_out = 0.5 * l2sum;
#pragma omp parallel
  {
    float* y = (float*)malloc(sizeof(float) * 3 * D);
#pragma omp for schedule(static)
    for (int64_t q = 0; q < terms; ++q) {
      const int64_t j = q / B, i = q % B;           /* j = 0: positives; j >= 1: corrupt batch j-1 */
      int h = pos[3 * i], t = pos[3 * i + 1], r = pos[3 * i + 2];
      float label = 1.f;
      if (j > 0) {
        const int n = neg_ent[(j - 1) * B + i];
        if (sides[j - 1]) h = n; else t = n;
        label = -1.f;
      }
      clip_row(E + (size_t)h * D, y, D); clip_row(E + (size_t)t * D, y + D, D);
      clip_row(E + (size_t)r * D, y + 2 * D, D);
      const float sc = score_rows(y, y + D, y + 2 * D, H);
      loss_out[q] = logf(1.0f + expf(-label * sc));
      const float g = -label * sigmoidf_(-label * sc);     /* d/ds log(1 + exp(-label s)) */
      side_grads(E, D, h, t, r, g, y, G + (size_t)(3 * q) * D, G + (size_t)(3 * q + 1) * D,
                 G + (size_t)(3 * q + 2) * D);
      idx[3 * q] = h; idx[3 * q + 1] = t; idx[3 * q + 2] = r;
    }
    free(y);
  }
  /* dense decay first (it reads the old table), then the sparse sum in term order */
  const float decay = lr * (float)terms * l2;
  if (decay != 0.f) {
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < N * (int64_t)D; ++p) E[p] -= decay * E[p];
  }
#pragma omp parallel
  {
#ifdef _OPENMP
    const int nt = omp_get_num_threads(), me = omp_get_thread_num();
#else
    const int nt = 1, me = 0;
#endif
    for (int64_t p = 0; p < 3 * terms; ++p) {
      const int row = idx[p];
      if (row % nt != me) continue;
      float* e = E + (size_t)row * D;
      const float* g = G + (size_t)p * D;
      for (int c = 0; c < D; ++c) e[c] -= g[c] * lr;
    }
  }
}

/* Rank counts: for each query, # candidates j in [cb, ce) with (s_j, j) < (s_true, true).
 * Yc = pre-clipped candidate rows [ce-cb, D]; qv = query vectors [Q, D] (App. A.4).
 * filter CSR may be NULL. */
void hole_ref_rank(const float* Yc, int64_t cb, int64_t ce, int D, const float* qv,
                   const int32_t* true_id, int64_t Q, const int64_t* foff, const int32_t* fids,
                   int32_t* raw_before, int32_t* filt_before) {
  const int64_t C = ce - cb;
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t q = 0; q < Q; ++q) {
    const float* v = qv + (size_t)q * D;
    const int64_t jt = true_id[q] - cb;
    float thr = 0.f;
    {
      const float* y = Yc + (size_t)jt * D;
      for (int k = 0; k < D; ++k) thr += y[k] * v[k];
    }
    int32_t raw = 0;
    for (int64_t j = 0; j < C; ++j) {
      const float* y = Yc + (size_t)j * D;
      float s = 0.f;
      for (int k = 0; k < D; ++k) s += y[k] * v[k];
      raw += (s < thr) || (s == thr && j < jt);
    }
    int32_t nf = 0;
    if (foff) {
      for (int64_t p = foff[q]; p < foff[q + 1]; ++p) {
        int64_t j = fids[p] - cb;
        if (j < 0 || j >= C) continue;
        const float* y = Yc + (size_t)j * D;
        float s = 0.f;
        for (int k = 0; k < D; ++k) s += y[k] * v[k];
        nf += (s < thr) || (s == thr && j < jt);
      }
    }
    raw_before[q] = raw;
    filt_before[q] = raw - nf;
  }
}

void hole_ref_clip_rows(const float* E, int64_t n, int D, float* Y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) clip_row(E + (size_t)i * D, Y + (size_t)i * D, D);
}
