/*
 * dcn_oracle.c — CPU restatement of the reference's DeformConv2d / TorchDeformConv2d.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is the checker the CUDA engine is compared with;
 * it is never linked into, imported by or called from the product (libdcn_b200.so and the
 * jittor_dcn_b200 package).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build or call it.
 *
 * What it restates (explicit loops, one IEEE-754 binary32 rounding per reference op):
 *   coordinates   deform_conv.py:62-68 + :34-39      train.py:102-113
 *   sampling      nn.grid_sample / F.grid_sample(bilinear, zeros, align_corners=True),
 *                 deform_conv.py:47-52, train.py:121-127.  The arithmetic lives in the
 *                 un-vendored third-party frameworks (torch, unpinned — README.md:16;
 *                 jittor, unpinned — README.md:15).  Published algorithm restated here:
 *                 ATen/native/GridSampler.h grid_sampler_unnormalize ((g+1)/2*(size-1)) and
 *                 within_bounds_2d zero padding; corner weights (1-fy)(1-fx), (1-fy)fx,
 *                 fy(1-fx), fy*fx.
 *   columns       deform_conv.py:54,72-73 (n,c order) / train.py:129-131 (raw reshape)
 *   GEMM + bias   deform_conv.py:74-80 / train.py:133-138
 *   backward      the autograd of the above (train.py:249 / train.py:414), written out
 *                 by hand.
 *
 * Parity status:
 *   DCN_VARIANT_TORCH  — PINNED: checked against outputs of the unmodified reference class
 *                        (AST-loaded from train.py:70-140) in tests/golden/ (made by
 *                        oracle/make_golden.py): corner indices/weights bit-exact, forward
 *                        and the four gradients to fp32 round-off.
 *   DCN_VARIANT_JITTOR — parity unpinned: jittor cannot be installed in the build
 *                        container, the reference holds no golden vectors, and
 *                        jittor.nn.grid_sample's rounding is restated from memory of its
 *                        documented formula (same as torch's).  Checked only against a
 *                        torch transliteration of deform_conv.py:30-81.
 *
 *   DCN_VARIANT_DCNV1  — (not a reference operator) torchvision.ops.deform_conv2d
 *                        semantics restated from its published CPU kernel (deform_conv2d_kernel.cpp:
 *                        bilinear_interpolate, deformable_im2col, one offset group, no mask).
 *                        PINNED to tests/golden/dcnv1_*.npz produced by torchvision 0.26 itself.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off: no FMA contraction, so every
 * float op below rounds exactly once, like the reference's op-by-op tensor chain).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/dcn_b200.h" /* DcnShape POD + variant enums only */

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct Geo {
  int B, C, O, H, W, N, Ho, Wo, K, P; /* P = Ho*Wo*N samples per (b,c) */
  float Dx, Dy;                       /* normalisation divisors */
  float sx, sy;                       /* (W-1)/2, (H-1)/2 */
  int variant;
  int kw, sh, sw, ph, pw;
} Geo;

/* offset channel that moves the ROW / COLUMN coordinate of tap n (see dcn_b200.h variants) */
static inline int off_row_ch(const Geo* g, int n) { return g->variant == DCN_VARIANT_DCNV1 ? 2 * n : n; }
static inline int off_col_ch(const Geo* g, int n) { return g->variant == DCN_VARIANT_DCNV1 ? 2 * n + 1 : g->N + n; }

static int make_geo(const DcnShape* s, Geo* g) {
  if (!s || s->B <= 0 || s->C <= 0 || s->O <= 0 || s->H <= 0 || s->W <= 0 || s->kh <= 0 ||
      s->kw <= 0 || s->sh <= 0 || s->sw <= 0 || s->ph < 0 || s->pw < 0)
    return -1;
  g->B = s->B; g->C = s->C; g->O = s->O; g->H = s->H; g->W = s->W;
  g->N = s->kh * s->kw;
  g->Ho = (s->H + 2 * s->ph - s->kh) / s->sh + 1; /* deform_conv.py:34 */
  g->Wo = (s->W + 2 * s->pw - s->kw) / s->sw + 1; /* deform_conv.py:35 */
  if (g->Ho <= 0 || g->Wo <= 0 || s->H + 2 * s->ph < s->kh || s->W + 2 * s->pw < s->kw) return -1;
  g->K = g->C * g->N;
  g->P = g->Ho * g->Wo * g->N;
  g->variant = s->variant;
  if (s->variant == DCN_VARIANT_TORCH) { /* train.py:111-112 */
    g->Dx = (float)(s->W - 1);
    g->Dy = (float)(s->H - 1);
  } else { /* deform_conv.py:37-38 */
    g->Dx = (float)(g->Wo - 1);
    g->Dy = (float)(g->Ho - 1);
  }
  g->sx = (float)(s->W - 1) / 2.0f;
  g->sy = (float)(s->H - 1) / 2.0f;
  g->kw = s->kw; g->sh = s->sh; g->sw = s->sw; g->ph = s->ph; g->pw = s->pw;
  return 0;
}

/* saturating float -> int that is total (NaN -> very negative) */
static inline int sat_int(float v) {
  if (!(v > -1073741824.0f)) return -1073741824;
  if (v > 1073741824.0f) return 1073741824;
  return (int)v;
}

typedef struct Corner {
  int y0, x0;   /* north-west corner (row, col) */
  float fx, fy; /* fractional parts */
  float w[4];   /* nw, ne, sw, se */
} Corner;

/* One sampling point.  off_x / off_y are the offsets of channels off_row_ch(n) / off_col_ch(n) at
 * (h,w): the reference's "x" offset (channel n) ends up moving the ROW because its grid is
 * [norm_y, norm_x]; DCNv1's dy (channel 2n) moves the row directly. */
static inline void corner_of(const Geo* g, int h, int w, int n, float off_x, float off_y, Corner* c) {
  if (g->variant == DCN_VARIANT_DCNV1) {
    /* torchvision deform_conv2d_kernel.cpp deformable_im2col:
     *   y = (out_y*stride_h - pad_h) + i*dil_h + offset_h ; x likewise; bilinear_interpolate */
    const int ki = n / g->kw, kj = n % g->kw;
    volatile float iy = (float)(h * g->sh - g->ph + ki) + off_x;
    volatile float ix = (float)(w * g->sw - g->pw + kj) + off_y;
    float xf = floorf(ix), yf = floorf(iy);
    c->fx = ix - xf;
    c->fy = iy - yf;
    c->x0 = sat_int(xf);
    c->y0 = sat_int(yf);
    volatile float e = 1.0f - c->fx, s = 1.0f - c->fy;
    c->w[0] = s * e;
    c->w[1] = s * c->fx;
    c->w[2] = c->fy * e;
    c->w[3] = c->fy * c->fx;
    return;
  }
  /* sampling_locs = grid + offset            deform_conv.py:68  train.py:109 */
  volatile float loc_x = (float)w + off_x;
  volatile float loc_y = (float)h + off_y;
  /* norm = loc / D * 2 - 1                   deform_conv.py:37-38  train.py:111-112 */
  volatile float nx = loc_x / g->Dx;
  nx = nx * 2.0f;
  nx = nx - 1.0f;
  volatile float ny = loc_y / g->Dy;
  ny = ny * 2.0f;
  ny = ny - 1.0f;
  /* grid = stack([norm_y, norm_x])           deform_conv.py:39  train.py:113
   * grid_sample reads slot 0 as the WIDTH coordinate, slot 1 as HEIGHT:          */
  volatile float ix = ny + 1.0f; /* GridSampler.h:27-36, align_corners=True */
  ix = ix * g->sx;
  volatile float iy = nx + 1.0f;
  iy = iy * g->sy;
  float xf = floorf(ix), yf = floorf(iy);
  c->fx = ix - xf;
  c->fy = iy - yf;
  c->x0 = sat_int(xf);
  c->y0 = sat_int(yf);
  volatile float e = 1.0f - c->fx; /* distance to east  */
  volatile float s = 1.0f - c->fy; /* distance to south */
  c->w[0] = s * e;
  c->w[1] = s * c->fx;
  c->w[2] = c->fy * e;
  c->w[3] = c->fy * c->fx;
}

static inline int inb(int v, int n) { return v >= 0 && v < n; }

/* zero-padded bilinear sample of one channel plane */
static inline float sample_plane(const float* xp, int H, int W, const Corner* c) {
  float v[4];
  int y0 = c->y0, x0 = c->x0;
  int oky0 = inb(y0, H), oky1 = (y0 >= -1 && y0 + 1 < H), okx0 = inb(x0, W),
      okx1 = (x0 >= -1 && x0 + 1 < W);
  v[0] = (oky0 && okx0) ? xp[(size_t)y0 * W + x0] : 0.0f;
  v[1] = (oky0 && okx1) ? xp[(size_t)y0 * W + x0 + 1] : 0.0f;
  v[2] = (oky1 && okx0) ? xp[(size_t)(y0 + 1) * W + x0] : 0.0f;
  v[3] = (oky1 && okx1) ? xp[(size_t)(y0 + 1) * W + x0 + 1] : 0.0f;
  volatile float acc = v[0] * c->w[0];
  volatile float t = v[1] * c->w[1];
  acc = acc + t;
  t = v[2] * c->w[2];
  acc = acc + t;
  t = v[3] * c->w[3];
  acc = acc + t;
  return acc;
}

/* (row r of the GEMM, column j) -> (channel c, pixel p, tap n) of the sample feeding it.
 * Jittor: A[(h,w), n*C + c]                    deform_conv.py:72-73
 * Torch : A[r, j] = S[b].flat[r*K + j], S[b] laid out (c,h,w,n)      train.py:129-131 */
static inline void col_map(const Geo* g, int r, int j, int* c, int* p, int* n) {
  if (g->variant == DCN_VARIANT_TORCH) {
    long long f = (long long)r * g->K + j;
    *c = (int)(f / g->P);
    int q = (int)(f % g->P);
    *p = q / g->N;
    *n = q % g->N;
  } else if (g->variant == DCN_VARIANT_DCNV1) {
    *c = j / g->N; /* torchvision columns: (c, tap) */
    *n = j % g->N;
    *p = r;
  } else {
    *n = j / g->C;
    *c = j % g->C;
    *p = r;
  }
}

/* ---------------------------------------------------------------------------------- */

int dcn_oracle_corners(const DcnShape* s, const float* off, int32_t* y0, int32_t* x0,
                       float* w4, float* fxy /* nullable [B,N,Ho,Wo,2] */) {
  Geo g;
  if (make_geo(s, &g)) return -1;
  const int HW = g.Ho * g.Wo;
  for (int b = 0; b < g.B; ++b)
    for (int n = 0; n < g.N; ++n)
      for (int p = 0; p < HW; ++p) {
        Corner c;
        const float* ob = off + (size_t)b * 2 * g.N * HW;
        corner_of(&g, p / g.Wo, p % g.Wo, n, ob[(size_t)off_row_ch(&g, n) * HW + p],
                  ob[(size_t)off_col_ch(&g, n) * HW + p], &c);
        size_t i = ((size_t)b * g.N + n) * HW + p;
        y0[i] = c.y0;
        x0[i] = c.x0;
        memcpy(w4 + 4 * i, c.w, sizeof c.w);
        if (fxy) {
          fxy[2 * i] = c.fx;
          fxy[2 * i + 1] = c.fy;
        }
      }
  return 0;
}

/* The [B, C, Ho, Wo, N] sample tensor the reference materialises (train.py:129). */
int dcn_oracle_sample(const DcnShape* s, const float* x, const float* off, float* S) {
  Geo g;
  if (make_geo(s, &g)) return -1;
  const int HW = g.Ho * g.Wo;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < g.B; ++b) {
    const float* ob = off + (size_t)b * 2 * g.N * HW;
    for (int p = 0; p < HW; ++p)
      for (int n = 0; n < g.N; ++n) {
        Corner c;
        corner_of(&g, p / g.Wo, p % g.Wo, n, ob[(size_t)off_row_ch(&g, n) * HW + p],
                  ob[(size_t)off_col_ch(&g, n) * HW + p], &c);
        for (int ch = 0; ch < g.C; ++ch) {
          const float* xp = x + ((size_t)b * g.C + ch) * g.H * g.W;
          S[(((size_t)b * g.C + ch) * HW + p) * g.N + n] = sample_plane(xp, g.H, g.W, &c);
        }
      }
  }
  return 0;
}

/* columns of one batch element as the GEMM sees them: A[HW, K] */
static void build_columns(const Geo* g, const float* S_b /* [C,HW,N] */, float* A) {
  const int HW = g->Ho * g->Wo;
  if (g->variant == DCN_VARIANT_TORCH) {
    memcpy(A, S_b, sizeof(float) * (size_t)HW * g->K); /* raw reshape */
  } else if (g->variant == DCN_VARIANT_DCNV1) {
    for (int p = 0; p < HW; ++p)
      for (int c = 0; c < g->C; ++c)
        for (int n = 0; n < g->N; ++n)
          A[(size_t)p * g->K + c * g->N + n] = S_b[((size_t)c * HW + p) * g->N + n];
  } else {
    for (int p = 0; p < HW; ++p)
      for (int n = 0; n < g->N; ++n)
        for (int c = 0; c < g->C; ++c)
          A[(size_t)p * g->K + n * g->C + c] = S_b[((size_t)c * HW + p) * g->N + n];
  }
}

int dcn_oracle_forward(const DcnShape* s, const float* x, const float* off, const float* wt,
                       const float* bias, float* out) {
  Geo g;
  if (make_geo(s, &g)) return -1;
  const int HW = g.Ho * g.Wo;
  float* S = (float*)malloc(sizeof(float) * (size_t)g.B * g.C * g.P);
  if (!S) return -2;
  dcn_oracle_sample(s, x, off, S);
#pragma omp parallel
  {
    float* A = (float*)malloc(sizeof(float) * (size_t)HW * g.K);
#pragma omp for schedule(static)
    for (int b = 0; b < g.B; ++b) {
      build_columns(&g, S + (size_t)b * g.C * g.P, A);
      for (int o = 0; o < g.O; ++o) {
        const float* wr = wt + (size_t)o * g.K; /* weight.reshape(O,-1)  :74 / :133 */
        for (int r = 0; r < HW; ++r) {
          const float* ar = A + (size_t)r * g.K;
          double acc = 0.0;
          for (int j = 0; j < g.K; ++j) acc += (double)ar[j] * (double)wr[j];
          float v = (float)acc;
          if (bias) v = v + bias[o]; /* :79-80 / :137-138 */
          out[((size_t)b * g.O + o) * HW + r] = v;
        }
      }
    }
    free(A);
  }
  free(S);
  return 0;
}

/* Backward, SURVEY.md Appendix A.4.  All outputs overwritten.  gx/goff/gw/gb may be NULL. */
int dcn_oracle_backward(const DcnShape* s, const float* x, const float* off, const float* wt,
                        const float* gout, float* gx, float* goff, float* gw, float* gb) {
  Geo g;
  if (make_geo(s, &g)) return -1;
  const int HW = g.Ho * g.Wo, N = g.N, C = g.C, K = g.K, O = g.O, H = g.H, W = g.W;
  const size_t xsz = (size_t)g.B * C * H * W;
  double* gxd = gx ? (double*)calloc(xsz, sizeof(double)) : NULL;
  double* gwd = (double*)calloc((size_t)O * K, sizeof(double));
  double* gbd = (double*)calloc((size_t)O, sizeof(double));
  float* S = (float*)malloc(sizeof(float) * (size_t)g.B * C * g.P);
  if ((gx && !gxd) || !gwd || !gbd || !S) return -2;
  dcn_oracle_sample(s, x, off, S);
  /* d(ix)/d(off_y) = 2/Dy * sx ; d(iy)/d(off_x) = 2/Dx * sy   (chain through :37-39) */
  const int v1 = g.variant == DCN_VARIANT_DCNV1; /* pixel coordinates: no chain-rule factor */
  const double mul_offx = v1 ? 1.0 : 2.0 / (double)g.Dx * (double)g.sy; /* row-moving offset */
  const double mul_offy = v1 ? 1.0 : 2.0 / (double)g.Dy * (double)g.sx; /* column-moving offset */

  float* A = (float*)malloc(sizeof(float) * (size_t)HW * K);
  double* gA = (double*)malloc(sizeof(double) * (size_t)HW * K);
  double* gix = (double*)malloc(sizeof(double) * (size_t)N * HW);
  double* giy = (double*)malloc(sizeof(double) * (size_t)N * HW);
  for (int b = 0; b < g.B; ++b) {
    const float* gob = gout + (size_t)b * O * HW; /* g[r,o] = gob[o*HW + r] */
    build_columns(&g, S + (size_t)b * C * g.P, A);
    /* gbias, gW = g^T A, gA = g Wm */
    memset(gA, 0, sizeof(double) * (size_t)HW * K);
    for (int o = 0; o < O; ++o) {
      const float* wr = wt + (size_t)o * K;
      double* gwr = gwd + (size_t)o * K;
      for (int r = 0; r < HW; ++r) {
        const double gv = gob[(size_t)o * HW + r];
        gbd[o] += gv;
        const float* ar = A + (size_t)r * K;
        double* gar = gA + (size_t)r * K;
        for (int j = 0; j < K; ++j) {
          gwr[j] += gv * ar[j];
          gar[j] += gv * wr[j];
        }
      }
    }
    /* scatter + coordinate gradient */
    memset(gix, 0, sizeof(double) * (size_t)N * HW);
    memset(giy, 0, sizeof(double) * (size_t)N * HW);
    const float* ob = off + (size_t)b * 2 * N * HW;
    for (int r = 0; r < HW; ++r)
      for (int j = 0; j < K; ++j) {
        int c, p, n;
        col_map(&g, r, j, &c, &p, &n);
        Corner cr;
        corner_of(&g, p / g.Wo, p % g.Wo, n, ob[(size_t)off_row_ch(&g, n) * HW + p],
                  ob[(size_t)off_col_ch(&g, n) * HW + p], &cr);
        const double gs = gA[(size_t)r * K + j];
        const float* xp = x + ((size_t)b * C + c) * H * W;
        double* gxp = gxd ? gxd + ((size_t)b * C + c) * H * W : NULL;
        double v[4];
        for (int k = 0; k < 4; ++k) {
          int yy = cr.y0 + (k >> 1), xx = cr.x0 + (k & 1);
          int ok = (cr.y0 > -1000000 && cr.x0 > -1000000 && cr.y0 < 1000000 && cr.x0 < 1000000) &&
                   inb(yy, H) && inb(xx, W);
          v[k] = ok ? xp[(size_t)yy * W + xx] : 0.0;
          if (ok && gxp) gxp[(size_t)yy * W + xx] += gs * cr.w[k];
        }
        const double fx = cr.fx, fy = cr.fy;
        gix[(size_t)n * HW + p] += gs * ((v[1] - v[0]) * (1.0 - fy) + (v[3] - v[2]) * fy);
        giy[(size_t)n * HW + p] += gs * ((v[2] - v[0]) * (1.0 - fx) + (v[3] - v[1]) * fx);
      }
    if (goff) {
      float* gob2 = goff + (size_t)b * 2 * N * HW;
      for (int n = 0; n < N; ++n)
        for (int p = 0; p < HW; ++p) {
          double a = giy[(size_t)n * HW + p], c2 = gix[(size_t)n * HW + p];
          /* a non-finite coordinate samples nothing and has zero gradient */
          gob2[(size_t)off_row_ch(&g, n) * HW + p] = (float)(a * mul_offx);
          gob2[(size_t)off_col_ch(&g, n) * HW + p] = (float)(c2 * mul_offy);
        }
    }
  }
  if (gx)
    for (size_t i = 0; i < xsz; ++i) gx[i] = (float)gxd[i];
  if (gw)
    for (size_t i = 0; i < (size_t)O * K; ++i) gw[i] = (float)gwd[i];
  if (gb)
    for (int o = 0; o < O; ++o) gb[o] = (float)gbd[o];
  free(gxd); free(gwd); free(gbd); free(S); free(A); free(gA); free(gix); free(giy);
  return 0;
}

int dcn_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
