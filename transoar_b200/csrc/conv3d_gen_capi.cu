// conv3d_gen_capi.cu -- C ABI of the general tcgen05 3x3x3 convolution (include/conv3d_gen.h): tile-box choice, tensor maps (incl. the
// eight parity-class views that express stride 2), step tables, launch.
#include "conv3d_gen_kernels.cuh"

#include <atomic>
#include <cstdlib>
#include <mutex>

#include "../../include/conv3d_gen.h"
#include "../../include/msda3d.h"

extern std::atomic<unsigned long long> g_msda3d_launches;

namespace {

using EncodeTiled = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled needs a context current on the calling thread (autograd's backward thread may have none yet)
void ensure_context_on_this_thread()
{
  static thread_local bool bound = false;
  if (!bound) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaSetDevice(dev);
    bound = true;
  }
}

EncodeTiled encode_fn()
{
  static EncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiled>(p);
  });
  return fn;
}

int sm_count()
{
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms = n;
  }
  return sms;
}

struct Vol { int N, D, H, W, C; };

// Channels-last volume [N, D, H, W, C], or -- step == 2 -- its parity class (pd, ph, pw): the voxels with index % 2 == p per axis, a plain
// 5-D tensor with doubled strides.  Box = 32 channels x (bw, bh, bd) voxels.  `load`: TF32-rounding loads, else fp32 stores.
int make_vol_map(CUtensorMap *map, const float *ptr, const Vol &v, int step, int pd, int ph, int pw, int bw, int bh, int bd, bool mn_major, bool load)
{
  EncodeTiled enc = encode_fn();
  if (enc == nullptr) return MSDA3D_ENODEV;
  const cuuint64_t C = (cuuint64_t)v.C, W = (cuuint64_t)v.W, H = (cuuint64_t)v.H, D = (cuuint64_t)v.D;
  const cuuint64_t vw = (W - pw + step - 1) / step, vh = (H - ph + step - 1) / step, vd = (D - pd + step - 1) / step;
  if (vw == 0 || vh == 0 || vd == 0) return MSDA3D_EINVAL;
  const cuuint64_t gdim[5] = {C, vw, vh, vd, (cuuint64_t)v.N};
  const cuuint64_t gstride[4] = {step * C * 4, step * W * C * 4, step * H * W * C * 4, D * H * W * C * 4};
  const cuuint32_t box[5] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const float *base = ptr + (((size_t)pd * H + ph) * W + pw) * C;
  const CUresult r = enc(map, load ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float *>(base), gdim, gstride,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         load ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MSDA3D_EINVAL;
}

// tap-major weights [27][CO][CI] as the 3-D tensor (ci, co, tap); box = 32 ci x rows co x 1 tap: the rows of one (tap, chunk) block are
// CI * 4 bytes apart (with the channels-last layout [CO][27][CI] they are 27 * CI * 4 bytes apart and every 128-byte row is its own L2 request)
int make_weight_map(CUtensorMap *map, const float *w, int CI, int CO, int rows, bool mn_major, int taps = 1)
{
  EncodeTiled enc = encode_fn();
  if (enc == nullptr) return MSDA3D_ENODEV;
  const cuuint64_t gdim[3] = {(cuuint64_t)CI, (cuuint64_t)CO, 27};
  const cuuint64_t gstride[2] = {(cuuint64_t)CI * 4, (cuuint64_t)CI * CO * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)rows, (cuuint32_t)taps};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, const_cast<float *>(w), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MSDA3D_EINVAL;
}

// box of `vox` (128 or 32) voxels over a W x H x D grid with the least padding; power-of-two edges, bw <= 32
void choose_box(int vox, int W, int H, int D, int *bw, int *bh, int *bd)
{
  long long best = -1;
  for (int w = 32; w >= 1; w >>= 1) {
    if (w > vox) continue;
    for (int h = vox / w; h >= 1; h >>= 1) {
      const int d = vox / (w * h);
      if (d > 64) continue;
      const long long pad = (long long)((W + w - 1) / w * w) * ((H + h - 1) / h * h) * ((D + d - 1) / d * d);
      if (best < 0 || pad < best) { best = pad; *bw = w; *bh = h; *bd = d; }
    }
  }
}

int choose_bn(int N)
{
  for (int c : {32, 64, 96, 128, 192, 256}) if (N <= c) return c;
  int bn = 128, best_tiles = 1 << 30, best_pad = 1 << 30;
  for (int c : {128, 192, 256}) {
    const int tiles = (N + c - 1) / c, pad = tiles * c - N;
    if (tiles < best_tiles || (tiles == best_tiles && pad < best_pad)) { best_tiles = tiles; best_pad = pad; bn = c; }
  }
  return bn;
}

template <int BN, bool B_MN> int launch_k(cudaStream_t st, const convgen::Problem &p, const float *bias)
{
  using C = tcgemm::Cfg<BN, 1>;
  auto kern = convgen::conv_kmajor_kernel<BN, B_MN>;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [&] { err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES); });
  if (err != cudaSuccess) return (int)err;
  const long long work = (long long)p.nclass * p.batch * p.td * p.th * p.tw * ((p.N + BN - 1) / BN) * p.ksplit;
  const int grid = (int)(work < sm_count() ? work : sm_count());
  kern<<<grid, tcgemm::kThreads, C::SMEM_BYTES, st>>>(p, bias);
  ++g_msda3d_launches;
  return (int)cudaGetLastError();
}

template <bool B_MN> int dispatch_k(cudaStream_t st, int bn, const convgen::Problem &p, const float *bias)
{
  switch (bn) {
    case 32: return launch_k<32, B_MN>(st, p, bias);
    case 64: return launch_k<64, B_MN>(st, p, bias);
    case 96: return launch_k<96, B_MN>(st, p, bias);
    case 128: return launch_k<128, B_MN>(st, p, bias);
    case 192: return launch_k<192, B_MN>(st, p, bias);
    default: return launch_k<256, B_MN>(st, p, bias);
  }
}

template <int BN, bool B_MN, int CH, int REG, int T = 1> int launch_h(cudaStream_t st, const convgen::HProblem &p, const float *bias, int smem)
{
  auto kern = convgen::conv_halo_kernel<BN, B_MN, CH, REG, T>;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [&] { err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  if (err != cudaSuccess) return (int)err;
  const long long work = (long long)p.nclass * p.batch * p.td * p.th * p.tw * ((p.N + BN - 1) / BN);
  const int grid = (int)(work < sm_count() ? work : sm_count());
  kern<<<grid, tcgemm::kThreads, smem, st>>>(p, bias);
  ++g_msda3d_launches;
  return (int)cudaGetLastError();
}

std::atomic<int> g_hdbg{0};
std::atomic<int> g_chains{0};      // 0 = default (two accumulation chains where TMEM allows), 1 = one chain (conv3d_gen_set_path bit 2: experiments)

// reg: 0 = table-driven taps (stride 2), 1 = stride-1 forward, 2 = stride-1 input gradient (unrolled constant taps)
template <bool B_MN, int REG> int dispatch_h2(cudaStream_t st, int bn, const convgen::HProblem &p, const float *bias, int smem)
{
  switch (bn) {
    case 32: return launch_h<32, B_MN, 2, REG>(st, p, bias, smem);
    case 64: return launch_h<64, B_MN, 2, REG>(st, p, bias, smem);
    case 96: return launch_h<96, B_MN, 2, REG>(st, p, bias, smem);
    case 128: return launch_h<128, B_MN, 2, REG>(st, p, bias, smem);
    case 192: return launch_h<192, B_MN, 1, REG>(st, p, bias, smem);
    default: return launch_h<256, B_MN, 1, REG>(st, p, bias, smem);
  }
}
template <bool B_MN, int REG> int dispatch_pair(cudaStream_t st, int bn, const convgen::HProblem &p, const float *bias, int smem)
{
  switch (bn) {
    case 32: return launch_h<32, B_MN, 1, REG, 2>(st, p, bias, smem);
    case 64: return launch_h<64, B_MN, 1, REG, 2>(st, p, bias, smem);
    case 96: return launch_h<96, B_MN, 1, REG, 2>(st, p, bias, smem);
    default: return launch_h<128, B_MN, 1, REG, 2>(st, p, bias, smem);
  }
}
template <bool B_MN> int dispatch_h(cudaStream_t st, int bn, const convgen::HProblem &p, const float *bias, int smem, int reg)
{
  if (p.pair) return B_MN ? dispatch_pair<B_MN, 2>(st, bn, p, bias, smem) : dispatch_pair<B_MN, 1>(st, bn, p, bias, smem);
  if (reg == 0 || g_chains.load() == 1) return dispatch_h2<B_MN, 0>(st, bn, p, bias, smem);     // (set_path bit 2: force the table-driven loop, experiments)
  if (reg == 3) return dispatch_h2<false, 3>(st, bn, p, bias, smem);
  return B_MN ? dispatch_h2<B_MN, 2>(st, bn, p, bias, smem) : dispatch_h2<B_MN, 1>(st, bn, p, bias, smem);
}

// 0 = choose per problem, 1 = always the per-tap kernel, 2 = the halo kernel wherever it fits (conv3d_gen_set_path; CONV3D_GEN_PATH=tap|halo)
std::atomic<int> g_path{-1};
int forced_path()
{
  int v = g_path.load();
  if (v < 0) {
    const char *e = getenv("CONV3D_GEN_PATH");
    v = e == nullptr ? 0 : e[0] == 't' ? 1 : e[0] == 'h' ? 2 : 0;
    g_path.store(v);
  }
  return v;
}

// the halo kernel pays for whole 8 x 16 tiles: use it where the tile rows are mostly real voxels
bool halo_wanted(int stride, int rows_h, int rows_w, long long planes, bool forward = false)
{
  const int f = forced_path();
  if (f == 1) return false;
  if (f == 2) return true;
  // measured per layer (profiles/r02_experiments.md): the halo kernel wins where its unrolled constant-tap loop applies (stride 1) and
  // the 16-row tiles are mostly real voxels; stride 2 runs its table-driven loop, which the per-tap kernel matches or beats
  if (stride != 1 && !forward) return false;
  const int th = (rows_h + 15) / 16, tw = (rows_w + 7) / 8;
  if (planes * th * tw < sm_count() / 2) return false;             // few tiles: the per-tap kernel splits the K loop over the idle SMs
  return (double)rows_h * rows_w >= 0.8 * (th * 16.0) * (tw * 8.0);
}

// few-tile layers (the coarse pyramid levels): split the (tap, chunk) loop so that every SM has a work item
int choose_ksplit(long long items, int ksteps)
{
  if (items >= sm_count()) return 1;
  // the split with the shortest makespan: rounds of work items on the persistent CTAs x (K-steps per item + its fixed cost, ~6 K-steps
  // for the accumulator hand-over and the stores)
  const long long sms = sm_count();
  const int cap = ksteps / 8 > 0 ? ksteps / 8 : 1;
  long long best = -1;
  int best_ks = 1;
  for (int ks = 1; ks <= cap; ++ks) {
    const long long per = (ksteps + ks - 1) / ks, rounds = (items * ks + sms - 1) / sms, span = rounds * (per + 6);
    if (best < 0 || span < best) { best = span; best_ks = ks; }
  }
  return best_ks;
}

// tile pairs (two 8 x 16 tiles side by side in w sharing the weight blocks): stride 1, narrow column tiles, W a multiple of 16
bool pair_wanted(int stride, int bn, int rows_w)
{
  return stride == 1 && bn <= 128 && rows_w % 16 == 0 && g_chains.load() != 1;
}

void h_add_box(convgen::HProblem &p, int b, int cls_hw, int ow, int oh, int lw, int lh)
{
  convgen::HBox &bx = p.boxes[b];
  bx.cls_hw = cls_hw; bx.ow = ow; bx.oh = oh; bx.lw = lw; bx.bytes = lw * lh * 128;
  bx.off = b == 0 ? 0 : (p.boxes[b - 1].off + p.boxes[b - 1].bytes + 1023) / 1024 * 1024;
  p.a_bytes += bx.bytes;
  p.a_stage_bytes = (bx.off + bx.bytes + 1023) / 1024 * 1024;
  p.nbox = b + 1;
}

// ring depths that fit 227 KB next to the epilogue staging: three plane stages (each feeds up to nine taps), the rest goes to the ring
// of weight blocks -- those are small (BN x 128 bytes) and the ring must hold a load latency's worth of them.  Returns the dynamic
// shared-memory size, or 0 if the problem does not fit.
int h_plan_smem(convgen::HProblem &p, int bn)
{
  // taps per weight box: every stage costs a barrier round trip in the single producer / issuer threads (~500 clocks measured), which
  // four MMAs of a narrow tile (N <= 128) do not cover
  p.tps = bn <= 32 ? 9 : bn <= 128 ? 3 : 1;
  p.dbg = g_hdbg.load();
  const int avail = 227 * 1024 - 1024 - 512 - convgen::kHEpiBytes;
  const int b_bytes = p.tps * bn * 128;
  for (int as : {3, 2}) {
    int bs = (avail - as * p.a_stage_bytes) / b_bytes;
    if (bs > convgen::kHMaxBStages) bs = convgen::kHMaxBStages;
    if (bs >= 3 || (bs >= 2 && (as == 2 || p.pair))) {
      p.a_stages = as; p.b_stages = bs;
      // weight-box bookkeeping of every tap: box start (a multiple of tps: the tap index is kd * 9 + kh * 3 + kw), slot, first-of-box flag
      for (int c = 0; c < p.nclass; ++c)
        for (int pl = 0; pl < p.cls[c].nplanes; ++pl) {
          convgen::HPlane &hp = p.cls[c].planes[pl];
          for (int t = 0; t < hp.ntaps; ++t) {
            convgen::HTap &tp = hp.taps[t];
            const int wtap = tp.wtap0;                               // the builders leave the tap index here
            tp.slot = wtap % p.tps;
            tp.wtap0 = wtap - tp.slot;
          }
          for (int t = 0; t < hp.ntaps; ++t) hp.taps[t].newgrp = (t == 0 || hp.taps[t].wtap0 != hp.taps[t - 1].wtap0) ? 1 : 0;
        }
      return 1024 + as * p.a_stage_bytes + bs * b_bytes + convgen::kHEpiBytes + 512;
    }
  }
  return 0;
}

template <int BHL, int GEO> int launch_w(cudaStream_t st, const convgen::WProblem &p, int grid, int smem)
{
  auto kern = convgen::conv_wgrad_kernel<BHL, GEO>;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [&] { err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  if (err != cudaSuccess) return (int)err;
  kern<<<grid, convgen::kThreadsW, smem, st>>>(p);
  ++g_msda3d_launches;
  return (int)cudaGetLastError();
}

bool shape_ok(int batch, int D, int H, int W, int ci, int co, int stride)
{
  if (batch <= 0 || D <= 0 || H <= 0 || W <= 0 || ci <= 0 || co <= 0 || ci % 4 || co % 4) return false;
  if (stride != 1 && stride != 2) return false;
  if (stride == 2 && (D < 2 || H < 2 || W < 2)) return false;
  return true;
}

// per-axis tap geometry under stride 2, seen from the output grid: tap k reads parity class par(k) of the input at index o + off(k)
inline int s2_par(int k) { return k == 1 ? 0 : 1; }
inline int s2_off(int k) { return k == 0 ? -1 : 0; }

}  // namespace

extern "C" void conv3d_gen_set_path(int path)
{
  g_path.store((path & 3) == 1 || (path & 3) == 2 ? (path & 3) : 0);
  g_chains.store((path & 4) ? 1 : 0);
  g_hdbg.store((path >> 3) & 15);
}

extern "C" int conv3d_gen_supported(int in_channels, int out_channels, int stride)
{
  return in_channels > 0 && out_channels > 0 && in_channels % 4 == 0 && out_channels % 4 == 0 && (stride == 1 || stride == 2);
}

extern "C" int conv3d_gen_forward(void *stream, const float *x, const float *w, const float *bias, int batch, int depth, int height, int width,
                                  int in_channels, int out_channels, int stride, float *y)
{
  if (!x || !w || !y || !shape_ok(batch, depth, height, width, in_channels, out_channels, stride)) return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w)) & 15) return MSDA3D_EALIGN;
  ensure_context_on_this_thread();
  const int OD = (depth + stride - 1) / stride, OH = (height + stride - 1) / stride, OW = (width + stride - 1) / stride;
  const Vol vx = {batch, depth, height, width, in_channels}, vy = {batch, OD, OH, OW, out_channels};
  int rc;
  if (halo_wanted(stride, OH, OW, (long long)batch * OD, true)) {
    convgen::HProblem h = {};
    h.batch = batch; h.tw = (OW + 7) / 8; h.th = (OH + 15) / 16; h.td = OD;
    h.N = out_channels; h.chunks = (in_channels + 31) / 32; h.nclass = 1;
    convgen::HClass &hc = h.cls[0];
    hc.nplanes = 3;
    h.pair = pair_wanted(stride, choose_bn(out_channels), OW) ? 1 : 0;
    if (h.pair) h.tw = (OW + 15) / 16;
    if (stride == 1) {
      h_add_box(h, 0, 0, -1, -1, h.pair ? 18 : 10, 18);
      if ((rc = make_vol_map(&h.tmA[0], x, vx, 1, 0, 0, 0, h.pair ? 18 : 10, 18, 1, false, true))) return rc;
    } else {
      h_add_box(h, 0, 3, -1, -1, 9, 17);             // odd h, odd w
      h_add_box(h, 1, 2, 0, -1, 8, 17);              // odd h, even w
      h_add_box(h, 2, 1, -1, 0, 9, 16);              // even h, odd w
      h_add_box(h, 3, 0, 0, 0, 8, 16);               // even h, even w
      for (int c = 0; c < 8; ++c)
        if ((rc = make_vol_map(&h.tmA[c], x, vx, 2, c >> 2, (c >> 1) & 1, c & 1, 8 + (c & 1), 16 + ((c >> 1) & 1), 1, false, true))) return rc;
    }
    for (int kd = 0; kd < 3; ++kd) {
      convgen::HPlane &hp = hc.planes[kd];
      hp.cls_d = stride == 1 ? 0 : s2_par(kd);
      hp.od = stride == 1 ? kd - 1 : s2_off(kd);
      hp.ntaps = 9;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          convgen::HTap &t = hp.taps[kh * 3 + kw];
          t.wtap0 = (kd * 3 + kh) * 3 + kw;
          if (stride == 1) { t.a_off = (kh * 10 + kw) * 128; t.sbo = 10 * 128; }
          else {
            const convgen::HBox &bx = h.boxes[(s2_par(kh) ? 0 : 2) + (s2_par(kw) ? 0 : 1)];
            t.a_off = bx.off + ((kh == 2 ? bx.lw : 0) + (kw == 2 ? 1 : 0)) * 128; t.sbo = bx.lw * 128;
          }
        }
    }
    const int bn = choose_bn(out_channels);
    const int smem = h_plan_smem(h, bn);
    if (smem > 0) {
      if ((rc = make_vol_map(&h.tmD[0], y, vy, 1, 0, 0, 0, 8, 4, 1, false, false))) return rc;
      if ((rc = make_weight_map(&h.tmB, w, in_channels, out_channels, bn, false, h.tps))) return rc;
      return dispatch_h<false>(reinterpret_cast<cudaStream_t>(stream), bn, h, bias, smem, stride == 1 ? 1 : 3);
    }
  }
  convgen::Problem p = {};
  choose_box(128, OW, OH, OD, &p.BW, &p.BH, &p.BD);
  p.qh = p.BH < 32 / p.BW ? p.BH : 32 / p.BW;
  p.qd = 32 / (p.BW * p.qh);
  p.batch = batch; p.tw = (OW + p.BW - 1) / p.BW; p.th = (OH + p.BH - 1) / p.BH; p.td = (OD + p.BD - 1) / p.BD;
  p.N = out_channels; p.chunks = (in_channels + 31) / 32; p.nclass = 1; p.nsteps[0] = 27;
  if (stride == 1) {
    if ((rc = make_vol_map(&p.tmA[0], x, vx, 1, 0, 0, 0, p.BW, p.BH, p.BD, false, true))) return rc;
  } else {
    for (int c = 0; c < 8; ++c)
      if ((rc = make_vol_map(&p.tmA[c], x, vx, 2, c >> 2, (c >> 1) & 1, c & 1, p.BW, p.BH, p.BD, false, true))) return rc;
  }
  for (int kd = 0; kd < 3; ++kd)
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        convgen::Step &s = p.steps[0][(kd * 3 + kh) * 3 + kw];
        s.tap = (kd * 3 + kh) * 3 + kw;
        if (stride == 1) { s.amap = 0; s.dw = (signed char)(kw - 1); s.dh = (signed char)(kh - 1); s.dd = (signed char)(kd - 1); }
        else { s.amap = (signed char)(s2_par(kd) * 4 + s2_par(kh) * 2 + s2_par(kw)); s.dw = (signed char)s2_off(kw); s.dh = (signed char)s2_off(kh); s.dd = (signed char)s2_off(kd); }
      }
  if ((rc = make_vol_map(&p.tmD[0], y, vy, 1, 0, 0, 0, p.BW, p.qh, p.qd, false, false))) return rc;
  const int bn = choose_bn(out_channels);
  if ((rc = make_weight_map(&p.tmB, w, in_channels, out_channels, bn, false))) return rc;
  p.ksplit = choose_ksplit((long long)batch * p.td * p.th * p.tw * ((out_channels + bn - 1) / bn), 27 * p.chunks);
  if (p.ksplit > 1) {
    const cudaError_t e = cudaMemsetAsync(y, 0, (size_t)batch * OD * OH * OW * out_channels * sizeof(float), reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return (int)e;
  }
  return dispatch_k<false>(reinterpret_cast<cudaStream_t>(stream), bn, p, bias);
}

extern "C" int conv3d_gen_dgrad(void *stream, const float *dy, const float *w, int batch, int depth, int height, int width, int in_channels,
                                int out_channels, int stride, float *dx)
{
  if (!dy || !w || !dx || !shape_ok(batch, depth, height, width, in_channels, out_channels, stride)) return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(w)) & 15) return MSDA3D_EALIGN;
  ensure_context_on_this_thread();
  const int OD = (depth + stride - 1) / stride, OH = (height + stride - 1) / stride, OW = (width + stride - 1) / stride;
  const Vol vdy = {batch, OD, OH, OW, out_channels}, vdx = {batch, depth, height, width, in_channels};
  int rc;
  if (halo_wanted(stride, OH, OW, (long long)batch * OD * (stride == 1 ? 1 : 8))) {
    convgen::HProblem h = {};
    h.batch = batch; h.tw = (OW + 7) / 8; h.th = (OH + 15) / 16; h.td = OD;
    h.N = in_channels; h.chunks = (out_channels + 31) / 32;
    h.pair = pair_wanted(stride, choose_bn(in_channels), OW) ? 1 : 0;
    if (h.pair) h.tw = (OW + 15) / 16;
    if (stride == 1) {
      h.nclass = 1;
      h_add_box(h, 0, 0, -1, -1, h.pair ? 18 : 10, 18);
      if ((rc = make_vol_map(&h.tmA[0], dy, vdy, 1, 0, 0, 0, h.pair ? 18 : 10, 18, 1, false, true))) return rc;
      if ((rc = make_vol_map(&h.tmD[0], dx, vdx, 1, 0, 0, 0, 8, 4, 1, false, false))) return rc;
      convgen::HClass &hc = h.cls[0];
      hc.nplanes = 3;
      for (int jd = 0; jd < 3; ++jd) {                     // box plane jd = dy plane i + jd - 1 = i + 1 - kd: kd = 2 - jd (same for h, w)
        convgen::HPlane &hp = hc.planes[jd];
        hp.cls_d = 0; hp.od = jd - 1; hp.ntaps = 9;
        for (int jh = 0; jh < 3; ++jh)
          for (int jw = 0; jw < 3; ++jw) {
            convgen::HTap &t = hp.taps[jh * 3 + jw];
            t.a_off = (jh * 10 + jw) * 128; t.sbo = 10 * 128; t.wtap0 = ((2 - jd) * 3 + (2 - jh)) * 3 + (2 - jw);
          }
      }
    } else {
      h.nclass = 8;
      h_add_box(h, 0, 0, 0, 0, 9, 17);
      if ((rc = make_vol_map(&h.tmA[0], dy, vdy, 1, 0, 0, 0, 9, 17, 1, false, true))) return rc;
      for (int c = 0; c < 8; ++c) {
        const int par[3] = {c >> 2, (c >> 1) & 1, c & 1};
        if ((rc = make_vol_map(&h.tmD[c], dx, vdx, 2, par[0], par[1], par[2], 8, 4, 1, false, false))) return rc;
        convgen::HClass &hc = h.cls[c];
        hc.nplanes = 0;
        for (int kd = 0; kd < 3; ++kd) {
          if (par[0] == 0 ? kd != 1 : kd == 1) continue;
          convgen::HPlane &hp = hc.planes[hc.nplanes++];
          hp.cls_d = 0; hp.od = kd == 0 ? 1 : 0; hp.ntaps = 0;
          for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
              if ((par[1] == 0 ? kh != 1 : kh == 1) || (par[2] == 0 ? kw != 1 : kw == 1)) continue;
              convgen::HTap &t = hp.taps[hp.ntaps++];
              t.a_off = ((kh == 0 ? 9 : 0) + (kw == 0 ? 1 : 0)) * 128; t.sbo = 9 * 128; t.wtap0 = (kd * 3 + kh) * 3 + kw;
            }
        }
      }
    }
    const int bn = choose_bn(in_channels);
    const int smem = h_plan_smem(h, bn);
    if (smem > 0) {
      if ((rc = make_weight_map(&h.tmB, w, in_channels, out_channels, 32, true))) return rc;
      return dispatch_h<true>(reinterpret_cast<cudaStream_t>(stream), bn, h, nullptr, smem, stride == 1 ? 2 : 0);
    }
  }
  convgen::Problem p = {};
  // stride 1: tiles over dx; stride 2: every parity class of dx is tiled over the dy grid (its own extent is that or one less)
  choose_box(128, OW, OH, OD, &p.BW, &p.BH, &p.BD);
  p.qh = p.BH < 32 / p.BW ? p.BH : 32 / p.BW;
  p.qd = 32 / (p.BW * p.qh);
  p.batch = batch; p.tw = (OW + p.BW - 1) / p.BW; p.th = (OH + p.BH - 1) / p.BH; p.td = (OD + p.BD - 1) / p.BD;
  p.N = in_channels; p.chunks = (out_channels + 31) / 32;
  if ((rc = make_vol_map(&p.tmA[0], dy, vdy, 1, 0, 0, 0, p.BW, p.BH, p.BD, false, true))) return rc;
  if (stride == 1) {
    p.nclass = 1; p.nsteps[0] = 27;
    for (int kd = 0; kd < 3; ++kd)
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          convgen::Step &s = p.steps[0][(kd * 3 + kh) * 3 + kw];
          s.tap = (kd * 3 + kh) * 3 + kw; s.amap = 0;
          s.dw = (signed char)(1 - kw); s.dh = (signed char)(1 - kh); s.dd = (signed char)(1 - kd);      // dx[i] = sum_k dy[i + 1 - k] W[k]
        }
    if ((rc = make_vol_map(&p.tmD[0], dx, vdx, 1, 0, 0, 0, p.BW, p.qh, p.qd, false, false))) return rc;
  } else {
    p.nclass = 8;
    for (int c = 0; c < 8; ++c) {
      const int par[3] = {c >> 2, (c >> 1) & 1, c & 1};                  // (pd, ph, pw)
      int n = 0;
      for (int kd = 0; kd < 3; ++kd)
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            const int k[3] = {kd, kh, kw};
            bool ok = true;
            for (int a = 0; a < 3; ++a) ok = ok && (par[a] == 0 ? k[a] == 1 : k[a] != 1);
            if (!ok) continue;
            convgen::Step &s = p.steps[c][n++];
            s.tap = (kd * 3 + kh) * 3 + kw; s.amap = 0;
            s.dd = (signed char)(kd == 0); s.dh = (signed char)(kh == 0); s.dw = (signed char)(kw == 0);   // dx[2 j + 1] takes dy[j + 1] W[0] + dy[j] W[2]
          }
      p.nsteps[c] = n;
      if ((rc = make_vol_map(&p.tmD[c], dx, vdx, 2, par[0], par[1], par[2], p.BW, p.qh, p.qd, false, false))) return rc;
    }
  }
  const int bn = choose_bn(in_channels);
  if ((rc = make_weight_map(&p.tmB, w, in_channels, out_channels, 32, true))) return rc;
  p.ksplit = choose_ksplit((long long)p.nclass * batch * p.td * p.th * p.tw * ((in_channels + bn - 1) / bn), (stride == 1 ? 27 : 8) * p.chunks);
  if (p.ksplit > 1) {
    const cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)batch * depth * height * width * in_channels * sizeof(float), reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return (int)e;
  }
  return dispatch_k<true>(reinterpret_cast<cudaStream_t>(stream), bn, p, nullptr);
}

extern "C" int conv3d_gen_wgrad(void *stream, const float *x, const float *dy, int batch, int depth, int height, int width, int in_channels,
                                int out_channels, int stride, float *dw)
{
  if (!x || !dy || !dw || !shape_ok(batch, depth, height, width, in_channels, out_channels, stride)) return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) return MSDA3D_EALIGN;
  ensure_context_on_this_thread();
  const int OD = (depth + stride - 1) / stride, OH = (height + stride - 1) / stride, OW = (width + stride - 1) / stride;
  convgen::WProblem p = {};
  // K-block = 8 x BH x BD output voxels (16 lines): least padded volume, then the smallest halo
  {
    long long best = -1;
    for (int bh : {16, 8, 4}) {
      const int bd = 16 / bh;
      const long long pad = (long long)((OH + bh - 1) / bh * bh) * ((OD + bd - 1) / bd * bd) * 1000 + (bh + 2) * bd;
      if (best < 0 || pad < best) { best = pad; p.BH = bh; p.BD = bd; }
    }
  }
  p.batch = batch; p.tw = (OW + 7) / 8; p.th = (OH + p.BH - 1) / p.BH; p.td = (OD + p.BD - 1) / p.BD;
  p.CI = in_channels; p.CO = out_channels; p.chunks = (in_channels + 31) / 32; p.dw = dw;
  const Vol vx = {batch, depth, height, width, in_channels}, vdy = {batch, OD, OH, OW, out_channels};
  int rc;
  auto add_box = [&](int b, int cls_hw, int ow, int oh, int lw, int lh) {
    convgen::WBox &bx = p.boxes[b];
    bx.cls_hw = cls_hw; bx.ow = ow; bx.oh = oh; bx.lw = lw; bx.lh = lh;
    bx.off = b == 0 ? 0 : (p.boxes[b - 1].off + p.boxes[b - 1].lw * p.boxes[b - 1].lh * p.BD * 128 + 1023) / 1024 * 1024;
    p.x_bytes += lw * lh * p.BD * 128;
  };
  auto add_group = [&](int g, int box, int oh, int t0, int t1, int t2) {
    convgen::WGroup &gr = p.groups[g];
    gr.box = box; gr.oh = oh; gr.tap[0] = t0; gr.tap[1] = t1; gr.tap[2] = t2; gr.tap[3] = -1;
  };
  int max_acc;
  if (stride == 1) {
    p.nbox = 1; p.ngroups = 3; max_acc = 3;
    add_box(0, 0, -1, -1, 10, p.BH + 2);
    for (int kh = 0; kh < 3; ++kh) add_group(kh, 0, kh, kh * 3, kh * 3 + 1, kh * 3 + 2);
    for (int kd = 0; kd < 3; ++kd) { p.kd_cls[kd] = 0; p.kd_off[kd] = kd - 1; }
    if ((rc = make_vol_map(&p.tmX[0], x, vx, 1, 0, 0, 0, 10, p.BH + 2, p.BD, true, true))) return rc;
  } else {
    p.nbox = 4; p.ngroups = 6; max_acc = 6;
    add_box(0, 3, -1, -1, 9, p.BH + 1);            // odd h, odd w: (kh, kw) in {0, 2} x {0, 2}
    add_box(1, 2, 0, -1, 8, p.BH + 1);             // odd h, even w: kh in {0, 2}, kw = 1
    add_box(2, 1, -1, 0, 9, p.BH);                 // even h, odd w: kh = 1, kw in {0, 2}
    add_box(3, 0, 0, 0, 8, p.BH);                  // even h, even w: the centre tap of the plane
    add_group(0, 0, 0, 0, 2, -1); add_group(1, 1, 0, 1, -1, -1);        // kh = 0
    add_group(2, 0, 1, 6, 8, -1); add_group(3, 1, 1, 7, -1, -1);        // kh = 2: one line further in the odd-h boxes
    add_group(4, 2, 0, 3, 5, -1); add_group(5, 3, 0, 4, -1, -1);        // kh = 1
    for (int kd = 0; kd < 3; ++kd) { p.kd_cls[kd] = s2_par(kd); p.kd_off[kd] = s2_off(kd); }
    for (int c = 0; c < 8; ++c) {
      const int ph = (c >> 1) & 1, pw = c & 1;
      if ((rc = make_vol_map(&p.tmX[c], x, vx, 2, c >> 2, ph, pw, 8 + pw, p.BH + ph, p.BD, true, true))) return rc;
    }
  }
  if ((rc = make_vol_map(&p.tmDy, dy, vdy, 1, 0, 0, 0, 8, p.BH, p.BD, true, true))) return rc;
  const int last = p.nbox - 1;
  p.dy_off = (p.boxes[last].off + p.boxes[last].lw * p.boxes[last].lh * p.BD * 128 + 1023) / 1024 * 1024;
  // co tile: as wide as the accumulators (512 TMEM columns / max_acc) and the shared memory allow, least padding first
  {
    const int cap = 512 / max_acc;
    long long best = -1;
    for (int bn : {128, 96, 64, 32}) {
      if (bn > cap) continue;
      if (2 * (p.dy_off + bn / 32 * convgen::kWDyChunkBytes) + 1024 + 256 > 227 * 1024) continue;
      const int tiles = (out_channels + bn - 1) / bn;
      const long long cost = (long long)tiles * bn * 16 + tiles;          // padded columns, then fewer tiles
      if (best < 0 || cost < best) { best = cost; p.BN = bn; p.n_tiles = tiles; }
    }
    if (best < 0) return MSDA3D_EINVAL;
  }
  p.stage_bytes = p.dy_off + p.BN / 32 * convgen::kWDyChunkBytes;
  p.dbg = g_hdbg.load();
  p.stages = (227 * 1024 - 1024 - 256) / p.stage_bytes;
  if (p.stages > convgen::kWMaxStages) p.stages = convgen::kWMaxStages;
  const int smem = p.stages * p.stage_bytes + 1024 + 256;
  const long long kblocks = (long long)batch * p.td * p.th * p.tw;
  const long long per_split = 3LL * p.chunks * p.n_tiles;
  // voxel-range splits: the persistent CTAs take work items round robin, so the kernel lasts ceil(items / SMs) rounds of kb_per_split
  // K-blocks -- pick the split count with the shortest makespan (300 items on 148 SMs is three rounds for four CTAs: 33 % tail)
  {
    const long long sms = sm_count();
    long long best_span = -1, best_s = 1;
    const long long s_max = kblocks < (8 * sms) / per_split + 1 ? kblocks : (8 * sms) / per_split + 1;
    for (long long s = 1; s <= s_max; ++s) {
      const long long kbps = (kblocks + s - 1) / s, actual = (kblocks + kbps - 1) / kbps;
      const long long rounds = (per_split * actual + sms - 1) / sms;
      const long long span = rounds * (kbps + 2);                          // + the epilogue of a work item, in K-block units
      if (best_span < 0 || span < best_span) { best_span = span; best_s = s; }
    }
    p.kb_per_split = (kblocks + best_s - 1) / best_s;
    p.splits = (int)((kblocks + p.kb_per_split - 1) / p.kb_per_split);
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)out_channels * 27 * in_channels * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  const long long work = per_split * p.splits;
  const int grid = (int)(work < sm_count() ? work : sm_count());
  // GEO 1 / 2: the stage geometry as compile-time constants (same numbers as the tables above; WGeo in the kernel header)
  if (stride == 1) {
    switch (p.BH) {
      case 16: return launch_w<4, 1>(st, p, grid, smem);
      case 8: return launch_w<3, 1>(st, p, grid, smem);
      default: return launch_w<2, 1>(st, p, grid, smem);
    }
  }
  switch (p.BH) {                       // stride 2: the table-driven loop (the 96-MMA constant-geometry unroll measured 30 % slower: 180 registers)
    case 16: return launch_w<4, 0>(st, p, grid, smem);
    case 8: return launch_w<3, 0>(st, p, grid, smem);
    default: return launch_w<2, 0>(st, p, grid, smem);
  }
}

// Stride-2 input gradient for narrow layers, with the eight parity classes of dx folded into the N extent of ONE problem.
// Every class reads the same 2 x 2 x 2 neighbourhood of dy: dx[2 j + p] = sum over delta in {0, 1}^3 of dy[j + delta] * W[k(p, delta)], with
// k(0, 0) = 1, k(1, 0) = 2, k(1, 1) = 0 per axis and no tap for (p, delta) = (0, 1).  So with w_fold[delta][class * CIP + ci][co] =
// W[k(class, delta)][co][ci] (zero where there is no tap; CIP = CI rounded up to 32) the whole input gradient is 8 K-steps per 32 output
// channels of 128 x (8 * CIP) MMAs instead of 27 K-steps of 128 x CIP ones -- the per-tap kernel is bound by its per-K-step issue overhead,
// not by the tensor pipe, so wide K-steps are what makes it fast (1.9 -> 0.5 ms on 24 <- 48 channels at 160 x 160 x 256).
extern "C" int conv3d_gen_dgrad_s2_folded(void *stream, const float *dy, const float *w_fold, int batch, int depth, int height, int width,
                                          int in_channels, int out_channels, float *dx)
{
  if (!dy || !w_fold || !dx || !shape_ok(batch, depth, height, width, in_channels, out_channels, 2) || in_channels > 64) return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(w_fold)) & 15) return MSDA3D_EALIGN;
  ensure_context_on_this_thread();
  const int OD = (depth + 1) / 2, OH = (height + 1) / 2, OW = (width + 1) / 2;
  const int cip = (in_channels + 31) / 32 * 32;
  convgen::Problem p = {};
  choose_box(128, OW, OH, OD, &p.BW, &p.BH, &p.BD);
  p.qh = p.BH < 32 / p.BW ? p.BH : 32 / p.BW;
  p.qd = 32 / (p.BW * p.qh);
  p.batch = batch; p.tw = (OW + p.BW - 1) / p.BW; p.th = (OH + p.BH - 1) / p.BH; p.td = (OD + p.BD - 1) / p.BD;
  p.N = 8 * cip; p.chunks = (out_channels + 31) / 32; p.nclass = 1; p.nsteps[0] = 8; p.fold_cip = cip; p.ksplit = 1;
  const Vol vdy = {batch, OD, OH, OW, out_channels}, vdx = {batch, depth, height, width, in_channels};
  int rc;
  if ((rc = make_vol_map(&p.tmA[0], dy, vdy, 1, 0, 0, 0, p.BW, p.BH, p.BD, false, true))) return rc;
  for (int d = 0; d < 8; ++d) {
    convgen::Step &s = p.steps[0][d];
    s.amap = 0; s.dd = (signed char)(d >> 2); s.dh = (signed char)((d >> 1) & 1); s.dw = (signed char)(d & 1); s.tap = d;
  }
  for (int c = 0; c < 8; ++c)
    if ((rc = make_vol_map(&p.tmD[c], dx, vdx, 2, c >> 2, (c >> 1) & 1, c & 1, p.BW, p.qh, p.qd, false, false))) return rc;
  // w_fold [8][8 * CIP][CO] as the tensor (co, n, delta): K-major rows of 32 output channels
  {
    EncodeTiled enc = encode_fn();
    if (enc == nullptr) return MSDA3D_ENODEV;
    const cuuint64_t gdim[3] = {(cuuint64_t)out_channels, (cuuint64_t)(8 * cip), 8};
    const cuuint64_t gstride[2] = {(cuuint64_t)out_channels * 4, (cuuint64_t)out_channels * 8 * cip * 4};
    const cuuint32_t box[3] = {32, 256, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, const_cast<float *>(w_fold), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MSDA3D_EINVAL;
  }
  return dispatch_k<false>(reinterpret_cast<cudaStream_t>(stream), 256, p, nullptr);
}

// experiments only (tools/probe_kshift.py): see k_sw128_probe_kernel
extern "C" int conv3d_gen_debug_k_probe(void *stream, const float *X, const float *Y, float *D, int row0, int group_stride_rows, int mode)
{
  if (!X || !Y || !D || row0 < 0 || group_stride_rows < 1 || row0 + 15 * group_stride_rows + 8 > convgen::kProbeRows) return MSDA3D_EINVAL;
  convgen::k_sw128_probe_kernel<<<1, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(X, Y, D, row0, group_stride_rows, mode);
  ++g_msda3d_launches;
  return (int)cudaGetLastError();
}

// experiments only (tools/probe_mma_rate.py): see mma_rate_probe_kernel
extern "C" int conv3d_gen_debug_mma_rate(void *stream, int layout, int n, int iters, long long *out_clocks)
{
  // layout bits 0-1: operand layout; bits 4-7: rotate over this many different operand tiles (0 = always the same tile)
  const int rotate = ((layout >> 4) & 15) >= 2 ? 4 : 1;
  layout &= 15;
  if (!out_clocks || layout < 0 || layout > 2 || n < 16 || n > 256 || n % 16 || iters < 1) return MSDA3D_EINVAL;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [&] { err = cudaFuncSetAttribute(convgen::mma_rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024); });
  if (err != cudaSuccess) return (int)err;
  convgen::mma_rate_probe_kernel<<<1, 128, 66 * 1024, reinterpret_cast<cudaStream_t>(stream)>>>(layout, n, iters, out_clocks, rotate);
  ++g_msda3d_launches;
  return (int)cudaGetLastError();
}
