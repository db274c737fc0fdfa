/*
 * TEST INFRASTRUCTURE ONLY -- type-generic body of the CPU oracle (see msda3d_oracle.c).
 * Included twice with T = float / double.  Never linked into the product library.
 *
 * Every block cites the reference line range it restates, always relative to
 *   transoar/models/ops/src/cuda/ms_deform_im2col_cuda.cuh   (abbreviated "cuh")
 */

#ifndef T
#error "define T, SUF, FMA_, FLOOR_ before including"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* Pixel coordinate of one axis, cuh:424-426:  `loc * size - 0.5`.
 * contract != 0 -> the single-rounding FFMA/DFMA that nvcc -O3 (fmad on, the default the reference's
 *                  setup.py builds with) emits for that expression; SASS evidence in
 *                  profiles/r01_reference_fwd_sass.txt (FFMA R, size, loc, -0.5).
 * contract == 0 -> the C source semantics: round the product to T, then subtract. */
static inline T FN(pix_)(T loc, int size, int contract)
{
  if (contract) return FMA_(loc, (T)size, (T)-0.5);
  volatile T prod = loc * (T)size;
  return prod - (T)0.5;
}

/* Everything that depends only on (loc, shape): cuh:36-46 and the range test cuh:428. */
typedef struct {
  int in_range;
  int d_low, h_low, w_low;
  T ld, lh, lw;
} FN(sample_);

static inline FN(sample_) FN(locate_)(T loc_w, T loc_h, T loc_d, int D, int H, int W, int contract)
{
  FN(sample_) s;
  const T d_im = FN(pix_)(loc_d, D, contract);
  const T h_im = FN(pix_)(loc_h, H, contract);
  const T w_im = FN(pix_)(loc_w, W, contract);
  s.in_range = (d_im > -1 && h_im > -1 && w_im > -1 && d_im < D && h_im < H && w_im < W); /* cuh:428 */
  s.d_low = (int)FLOOR_(d_im);   /* cuh:36-38 */
  s.h_low = (int)FLOOR_(h_im);
  s.w_low = (int)FLOOR_(w_im);
  s.ld = d_im - (T)s.d_low;      /* cuh:43-45 */
  s.lh = h_im - (T)s.h_low;
  s.lw = w_im - (T)s.w_low;
  return s;
}

/* Corner c (0..7) in the reference's v1..v8 order, cuh:60-107:
 *   bit2 -> d high, bit1 -> h high, bit0 -> w high.  Returns voxel index or -1 when zero-padded. */
static inline long FN(corner_)(const FN(sample_) *s, int c, int D, int H, int W)
{
  const int d = s->d_low + ((c >> 2) & 1), h = s->h_low + ((c >> 1) & 1), w = s->w_low + (c & 1);
  if (d < 0 || h < 0 || w < 0 || d > D - 1 || h > H - 1 || w > W - 1) return -1;
  return ((long)d * H + h) * W + w;
}

/* The eight trilinear weights, cuh:109-110, products evaluated left to right. */
static inline void FN(weights_)(const FN(sample_) *s, T w[8])
{
  const T ld = s->ld, lh = s->lh, lw = s->lw;
  const T hd = 1 - ld, hh = 1 - lh, hw = 1 - lw;   /* cuh:46 */
  w[0] = hd * hh * hw; w[1] = hd * hh * lw; w[2] = hd * lh * hw; w[3] = hd * lh * lw;
  w[4] = ld * hh * hw; w[5] = ld * hh * lw; w[6] = ld * lh * hw; w[7] = ld * lh * lw;
}

/* cuh:112.  contract: FMUL(w2,v2) -> FFMA(w1,v1,.) -> FFMA(w3,v3,.) ... -> FFMA(w8,v8,.), the order in
 * the compiled reference; otherwise plain left-to-right sums of rounded products. */
static inline T FN(blend_)(const T w[8], const T v[8], int contract)
{
  if (contract) {
    T acc = w[1] * v[1];
    acc = FMA_(w[0], v[0], acc);
    for (int i = 2; i < 8; ++i) acc = FMA_(w[i], v[i], acc);
    return acc;
  }
  volatile T acc = w[0] * v[0];
  for (int i = 1; i < 8; ++i) { volatile T p = w[i] * v[i]; acc = acc + p; }
  return acc;
}

/* Forward: cuh:370-439 (kernel), shapes per ms_deform_attn_cuda.cu:40-60.
 *   value [N,S,M,C], shapes int64 [L,3] = (D,H,W), starts int64 [L], loc [N,Lq,M,L,P,3] = (x->W,y->H,z->D),
 *   aw [N,Lq,M,L,P]  ->  out [N,Lq,M*C]. */
void FN(msda3d_oracle_forward_)(const T *value, const int64_t *shapes, const int64_t *starts,
                                const T *loc, const T *aw, int N, int S, int M, int C, int L, int Lq, int P,
                                T *out, int contract)
{
  const long units = (long)N * Lq * M;
#pragma omp parallel for schedule(static)
  for (long u = 0; u < units; ++u) {
    const int m = (int)(u % M);
    const int b = (int)(u / ((long)M * Lq));
    const T *loc_u = loc + u * L * P * 3;
    const T *aw_u = aw + u * L * P;
    T *out_u = out + u * C;
    for (int c = 0; c < C; ++c) out_u[c] = 0;
    for (int l = 0; l < L; ++l) {
      const int D = (int)shapes[l * 3], H = (int)shapes[l * 3 + 1], W = (int)shapes[l * 3 + 2];
      const T *val_l = value + ((long)b * S + starts[l]) * M * C;   /* cuh:412 */
      for (int p = 0; p < P; ++p) {
        const T *xyz = loc_u + (l * P + p) * 3;
        const T weight = aw_u[l * P + p];
        const FN(sample_) s = FN(locate_)(xyz[0], xyz[1], xyz[2], D, H, W, contract);
        if (!s.in_range) continue;
        T w[8];
        long vox[8];
        FN(weights_)(&s, w);
        for (int k = 0; k < 8; ++k) vox[k] = FN(corner_)(&s, k, D, H, W);
        for (int c = 0; c < C; ++c) {
          T v[8];
          for (int k = 0; k < 8; ++k) v[k] = vox[k] < 0 ? (T)0 : val_l[vox[k] * M * C + m * C + c];
          const T val = FN(blend_)(w, v, contract);
          if (contract) out_u[c] = FMA_(weight, val, out_u[c]);       /* cuh:430 */
          else { volatile T pr = val * weight; out_u[c] = out_u[c] + pr; }
        }
      }
    }
  }
}

/* Backward: cuh:116-241 (per sample/channel), kernel loop cuh:551-661, zero-init ms_deform_attn_cuda.cu:122-124.
 * grad_value is ACCUMULATED into (caller zero-fills, as at::zeros_like does in the reference);
 * grad_loc / grad_aw are overwritten.  The per-(b,m) slices of grad_value are disjoint, so the loop is
 * parallel over (b,m) and deterministic.  Channel sums for grad_loc/grad_aw run c = 0..C-1 in order
 * (the reference tree-reduces in shared memory, cuh:632-643 -- tolerance-level difference only). */
void FN(msda3d_oracle_backward_)(const T *grad_out, const T *value, const int64_t *shapes, const int64_t *starts,
                                 const T *loc, const T *aw, int N, int S, int M, int C, int L, int Lq, int P,
                                 T *grad_value, T *grad_loc, T *grad_aw, int contract)
{
  const int NM = N * M;
#pragma omp parallel for schedule(dynamic, 1)
  for (int bm = 0; bm < NM; ++bm) {
    const int b = bm / M, m = bm % M;
    for (int q = 0; q < Lq; ++q) {
      const long u = ((long)b * Lq + q) * M + m;
      const T *go = grad_out + u * C;
      for (int l = 0; l < L; ++l) {
        const int D = (int)shapes[l * 3], H = (int)shapes[l * 3 + 1], W = (int)shapes[l * 3 + 2];
        const long lvl_off = ((long)b * S + starts[l]) * M * C;
        const T *val_l = value + lvl_off;
        T *gval_l = grad_value + lvl_off;
        for (int p = 0; p < P; ++p) {
          const long si = u * L * P + l * P + p;
          const T *xyz = loc + si * 3;
          const T weight = aw[si];
          T g_w = 0, g_h = 0, g_d = 0, g_a = 0;
          const FN(sample_) s = FN(locate_)(xyz[0], xyz[1], xyz[2], D, H, W, contract);
          if (s.in_range) {
            T w[8];
            long vox[8];
            FN(weights_)(&s, w);
            for (int k = 0; k < 8; ++k) vox[k] = FN(corner_)(&s, k, D, H, W);
            const T ld = s.ld, lh = s.lh, lw = s.lw, hd = 1 - ld, hh = 1 - lh, hw = 1 - lw;
            /* d(weight_k)/d(ld|lh|lw): cuh:159-231, same corner order as `corner_` */
            const T dd[8] = {-(hh * hw), -(hh * lw), -(lh * hw), -(lh * lw), hh * hw, hh * lw, lh * hw, lh * lw};
            const T dh[8] = {-(hd * hw), -(hd * lw), hd * hw, hd * lw, -(ld * hw), -(ld * lw), ld * hw, ld * lw};
            const T dw[8] = {-(hd * hh), hd * hh, -(hd * lh), hd * lh, -(ld * hh), ld * hh, -(ld * lh), ld * lh};
            for (int c = 0; c < C; ++c) {
              const T top = go[c];
              const T tgv = top * weight;              /* cuh:151 */
              T v[8], gd = 0, gh = 0, gw = 0;
              for (int k = 0; k < 8; ++k) {
                if (vox[k] < 0) { v[k] = 0; continue; }
                const long a = vox[k] * M * C + m * C + c;
                v[k] = val_l[a];
                gd += dd[k] * v[k]; gh += dh[k] * v[k]; gw += dw[k] * v[k];
                gval_l[a] += w[k] * tgv;                /* cuh:166 ... 231 (atomicAdd) */
              }
              g_a += top * FN(blend_)(w, v, contract);  /* cuh:236-237 */
              g_w += (T)W * gw * tgv;                   /* cuh:238-240 */
              g_h += (T)H * gh * tgv;
              g_d += (T)D * gd * tgv;
            }
          }
          grad_loc[si * 3 + 0] = g_w;
          grad_loc[si * 3 + 1] = g_h;
          grad_loc[si * 3 + 2] = g_d;
          grad_aw[si] = g_a;
        }
      }
    }
  }
}

/* Sampling-index arithmetic only, for the bit-exact index parity test.  Per sample (N*Lq*M*L*P):
 *   idx[4] = {in_range, d_low, h_low, w_low}  (int32),  frac[3] = {ld, lh, lw}. */
void FN(msda3d_oracle_indices_)(const int64_t *shapes, const T *loc, int N, int M, int L, int Lq, int P,
                                int32_t *idx, T *frac, int contract)
{
  const long T_ = (long)N * Lq * M * L * P;
#pragma omp parallel for schedule(static)
  for (long si = 0; si < T_; ++si) {
    const int l = (int)((si / P) % L);
    const int D = (int)shapes[l * 3], H = (int)shapes[l * 3 + 1], W = (int)shapes[l * 3 + 2];
    const FN(sample_) s = FN(locate_)(loc[si * 3], loc[si * 3 + 1], loc[si * 3 + 2], D, H, W, contract);
    idx[si * 4 + 0] = s.in_range; idx[si * 4 + 1] = s.d_low; idx[si * 4 + 2] = s.h_low; idx[si * 4 + 3] = s.w_low;
    frac[si * 3 + 0] = s.ld; frac[si * 3 + 1] = s.lh; frac[si * 3 + 2] = s.lw;
  }
}

#undef FN
#undef CAT
#undef CAT_
