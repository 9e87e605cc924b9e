// Training-path convolution of the head (north_star (4), SURVEY K-p): forward AND backward of one `ConvModule`
// (conv + bias + ReLU, reference tps_pp.py:126-131,149-154,538-548) as native kernels, so that autograd of the rectifier's
// convolutions -- 88 % of its FLOPs -- no longer goes through cuDNN.
//
//   forward   y = relu(conv(x, w) + b)                        the inference engine (conv_tma_kernel / conv_ts_kernel)
//   backward  g  = gy * [y > 0],  gb = sum_{b,p} g            relu_mask_bias_kernel (+ finalize)
//             gx = conv_transpose(g, w)                       the SAME tcgen05 engine: a stride-1 convolution over g with the
//                                                             weight image transposed and mirrored (wprep dg_*), 64 input
//                                                             channels per launch; a strided forward becomes a convolution
//                                                             over the ZERO-INSERTED gradient (ConvArgs::zi)
//             gw[co][ci][t] = sum_{b,p} g[b,co,p] x[b,ci,p+t]  wgrad_kernel: fp32 FMA outer products (64 x 64 tile per tap and
//                                                             pixel split, deterministic two-stage reduction)
// Geometry: NCHW fp32, Cout = 64, Cin a multiple of 32, 1x1 (stride 1) or 3x3 (pad 1; stride 1, 2 or (2,1)).
#include "head.cuh"

#include <string.h>

namespace tpspp {

// ---- g = gy * [y > 0] and per-channel partial sums (deterministic: fixed split of the batch, fixed tree) ----
constexpr int MB_SPLITS = 16;
__global__ void __launch_bounds__(256) relu_mask_bias_kernel(const float* __restrict__ y, const float* __restrict__ gy,
                                                             float* __restrict__ g, float* __restrict__ part, int B, int HW, int relu) {
  const int c = blockIdx.x, sp = blockIdx.y;
  const int b0 = (int)((long long)B * sp / MB_SPLITS), b1 = (int)((long long)B * (sp + 1) / MB_SPLITS);
  float acc = 0.f;
  for (int b = b0; b < b1; ++b) {
    const size_t base = ((size_t)b * 64 + c) * HW;
    for (int i = threadIdx.x * 4; i < HW; i += 256 * 4) {        // HW is a multiple of 4 for every layer of the head
      const float4 gv = *reinterpret_cast<const float4*>(gy + base + i);
      float4 r = gv;
      if (relu) {
        const float4 yv = *reinterpret_cast<const float4*>(y + base + i);
        r.x = yv.x > 0.f ? gv.x : 0.f; r.y = yv.y > 0.f ? gv.y : 0.f; r.z = yv.z > 0.f ? gv.z : 0.f; r.w = yv.w > 0.f ? gv.w : 0.f;
      }
      *reinterpret_cast<float4*>(g + base + i) = r;
      acc += (r.x + r.y) + (r.z + r.w);
    }
  }
  __shared__ float red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[sp * 64 + c] = red[0];
}
__global__ void bias_finalize_kernel(const float* __restrict__ part, float* __restrict__ gb) {
  const int c = threadIdx.x;
  float acc = 0.f;
  for (int s = 0; s < MB_SPLITS; ++s) acc += part[s * 64 + c];
  gb[c] = acc;
}

// ---- weight gradient: one block = one (tap, 64-input-channel block, pixel split): a 64 x 64 tile of outer products over its
//      32-pixel chunks.  Shared tiles are stored pixel-major ([k][channel], pitch 68) so that a thread reads the four output
//      and the four input channels of its 4 x 4 register tile with one LDS.128 each per pixel: 2 loads per 16 FMAs. ----
struct WgradArgs {
  const float *g, *x;          // g [B,64,Ho,Wo], x [B,Cin,H,W]
  float* part;                 // [splits][T * Cin * 64] as [tap][ci][co]
  int B, Cin, H, W, Ho, Wo, KS, sh, sw, splits;
};
constexpr int WG_PITCH = 68;
__global__ void __launch_bounds__(256) wgrad_kernel(WgradArgs a) {
  __shared__ __align__(16) float Gs[32 * WG_PITCH];
  __shared__ __align__(16) float Xs[32 * WG_PITCH];
  const int T = a.KS * a.KS;
  const int tap = blockIdx.x % T, cib = blockIdx.x / T, sp = blockIdx.y;
  const int dy = tap / a.KS, dx = tap - dy * a.KS, pad = a.KS / 2;
  const int HoWo = a.Ho * a.Wo, HW = a.H * a.W;
  const long long nchunks = (long long)a.B * HoWo / 32;
  const long long c0 = nchunks * sp / a.splits, c1 = nchunks * (sp + 1) / a.splits;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int co4 = (tid >> 4) * 4, ci4 = (tid & 15) * 4;
  const int cw = a.Cin - cib * 64 < 64 ? a.Cin - cib * 64 : 64;     // input channels of this block (32 for the 32-channel layers)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long ch = c0; ch < c1; ++ch) {
    const long long p0 = ch * 32;
    const int b = (int)(p0 / HoWo);
    const int p = (int)(p0 - (long long)b * HoWo) + lane;       // this lane's output pixel
    const int oy = p / a.Wo, ox = p - oy * a.Wo;
    const int iy = oy * a.sh + dy - pad, ix = ox * a.sw + dx - pad;
    const bool ok = iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
    const float* gp = a.g + (size_t)b * 64 * HoWo + p;
    const float* xp = a.x + ((size_t)b * a.Cin + cib * 64) * HW + (ok ? iy * a.W + ix : 0);
    __syncthreads();                                             // the previous chunk's tiles are no longer read
#pragma unroll
    for (int r = 0; r < 8; ++r) {                                // warp w stages channels w, w + 8, ...: lanes = pixels, coalesced
      const int c = warp + 8 * r;
      Gs[lane * WG_PITCH + c] = __ldg(gp + (size_t)c * HoWo);
      Xs[lane * WG_PITCH + c] = (ok && c < cw) ? __ldg(xp + (size_t)c * HW) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float4 gv = *reinterpret_cast<const float4*>(Gs + k * WG_PITCH + co4);
      const float4 xv = *reinterpret_cast<const float4*>(Xs + k * WG_PITCH + ci4);
      const float gg[4] = {gv.x, gv.y, gv.z, gv.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gg[i], xx[j], acc[i][j]);
    }
  }
  float* o = a.part + ((size_t)sp * T + tap) * a.Cin * 64 + (size_t)(cib * 64) * 64;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (ci4 + j < cw) *reinterpret_cast<float4*>(o + (size_t)(ci4 + j) * 64 + co4) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
}
// gw[co][ci][tap] = sum over splits of part[split][tap][ci][co]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ gw, int Cin, int T, int splits) {
  const int total = 64 * Cin * T;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    const int co = i / (Cin * T), r = i - co * (Cin * T), ci = r / T, tap = r - ci * T;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += __ldg(part + ((size_t)s * T + tap) * Cin * 64 + (size_t)ci * 64 + co);
    gw[i] = acc;
  }
}

struct ConvDims { int B, Cin, H, W, KS, sh, sw, Ho, Wo, relu, T, slices, splits; };
static int conv_dims(const tpspp_conv_cfg* c, ConvDims* d) {
  TPSPP_REQUIRE(c != nullptr, "conv cfg is NULL");
  TPSPP_REQUIRE(c->batch >= 0, "batch must be >= 0");
  TPSPP_REQUIRE(c->cin > 0 && c->cin % 32 == 0 && (c->cin <= 64 || c->cin % 64 == 0), "cin must be 32 or a multiple of 64 (got %d)", c->cin);
  TPSPP_REQUIRE(c->ksize == 1 || c->ksize == 3, "kernel size must be 1 or 3");
  TPSPP_REQUIRE((c->stride_h == 1 && c->stride_w == 1) || (c->ksize == 3 && c->stride_h == 2 && (c->stride_w == 1 || c->stride_w == 2)),
                "stride must be 1, or (2,2) / (2,1) for a 3x3 kernel (got %d,%d)", c->stride_h, c->stride_w);
  TPSPP_REQUIRE(c->height % c->stride_h == 0 && c->width % c->stride_w == 0, "input size must be a multiple of the stride");
  d->B = c->batch; d->Cin = c->cin; d->H = c->height; d->W = c->width; d->KS = c->ksize; d->sh = c->stride_h; d->sw = c->stride_w;
  d->Ho = d->H / d->sh; d->Wo = d->W / d->sw; d->relu = c->relu != 0; d->T = d->KS * d->KS;
  TPSPP_REQUIRE(d->B == 0 || ((long long)d->B * d->Ho * d->Wo) % 128 == 0, "batch * output pixels must be a multiple of 128");
  TPSPP_REQUIRE((d->Ho * d->Wo) % 32 == 0 && (d->H * d->W) % 4 == 0, "output plane must be a multiple of 32 pixels");
  d->slices = (d->Cin + 63) / 64;
  // pixel splits of the weight gradient: enough blocks for ~2 waves, at least 8 chunks of 32 pixels per block
  long long chunks = (long long)d->B * d->Ho * d->Wo / 32;
  long long blocks_per_split = (long long)d->T * d->slices;
  long long sp = (2LL * sm_count() * 2 + blocks_per_split - 1) / blocks_per_split;
  if (sp > chunks / 8) sp = chunks / 8;
  if (sp < 1) sp = 1;
  if (sp > 64) sp = 64;
  d->splits = (int)sp;
  return TPSPP_OK;
}
enum { CW_WFWD = 0, CW_WDG, CW_G, CW_BPART, CW_WPART, CW_COUNT };
static void conv_offsets(const ConvDims& d, size_t* off, size_t* total) {
  size_t sz[CW_COUNT];
  sz[CW_WFWD] = conv_tc_wprep_floats(d.Cin, d.KS, 64);
  sz[CW_WDG] = (size_t)d.slices * conv_tc_wprep_floats(64, d.KS, 64);
  sz[CW_G] = (size_t)d.B * 64 * d.Ho * d.Wo;
  sz[CW_BPART] = MB_SPLITS * 64;
  sz[CW_WPART] = (size_t)d.splits * d.T * d.Cin * 64;
  size_t cur = 0;
  for (int i = 0; i < CW_COUNT; ++i) {
    off[i] = cur;
    cur += (sz[i] * sizeof(float) + 255) / 256 * 256;
  }
  *total = cur + 256;
}

}  // namespace tpspp

using namespace tpspp;

extern "C" size_t tpspp_conv_workspace_bytes(const tpspp_conv_cfg* cfg) {
  ConvDims d;
  if (conv_dims(cfg, &d) != TPSPP_OK) return 0;
  size_t off[CW_COUNT], total;
  conv_offsets(d, off, &total);
  return total;
}

extern "C" int tpspp_conv_fwd(const tpspp_conv_cfg* cfg, const float* x, const float* w, const float* bias, float* y,
                              void* workspace, tpspp_stream_t stream) {
  reset_launch_count();
  ConvDims d;
  int rc = conv_dims(cfg, &d);
  if (rc != TPSPP_OK) return rc;
  if (d.B == 0) return TPSPP_OK;
  TPSPP_REQUIRE(x && w && y && workspace, "tpspp_conv_fwd: null pointer");
  TPSPP_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)workspace) & 15) == 0, "tpspp_conv_fwd: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  size_t off[CW_COUNT], total;
  conv_offsets(d, off, &total);
  float* wimg = reinterpret_cast<float*>((char*)workspace + off[CW_WFWD]);
  const int mode = d.KS == 3 ? CM_MIX : CM_TF32X3;
  WPrepLayer L;
  memset(&L, 0, sizeof(L));
  L.w = w; L.out = wimg; L.Ctot = d.Cin; L.taps = d.T; L.N = 64; L.NT = 64; L.bf16 = mode;
  rc = conv_tc_prepare_weights(&L, 1, st);
  if (rc != TPSPP_OK) return rc;
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.src[0].ptr = x; a.src[0].C = d.Cin; a.src[0].H = d.H; a.src[0].W = d.W; a.src[0].uh = 1; a.src[0].uw = 1;
  a.src[1].H = a.src[1].W = a.src[1].uh = a.src[1].uw = 1; a.src[2] = a.src[1];
  a.weight = w; a.bias = bias; a.out = y; a.B = d.B; a.Ho = d.Ho; a.Wo = d.Wo; a.Ctot = d.Cin; a.sh = d.sh; a.sw = d.sw;
  a.pad = d.KS / 2; a.act = d.relu ? CONV_ACT_RELU : CONV_ACT_NONE; a.act_scale = 1.f; a.Cout = 64;
  return run_conv_tc(d.KS, a, wimg, 64, st, mode);
}

extern "C" int tpspp_conv_bwd(const tpspp_conv_cfg* cfg, const float* x, const float* w, const float* y, const float* gy,
                              float* gx, float* gw, float* gb, void* workspace, tpspp_stream_t stream) {
  reset_launch_count();
  ConvDims d;
  int rc = conv_dims(cfg, &d);
  if (rc != TPSPP_OK) return rc;
  if (d.B == 0) return TPSPP_OK;
  TPSPP_REQUIRE(x && w && y && gy && workspace, "tpspp_conv_bwd: null pointer");
  TPSPP_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gy | (uintptr_t)gx | (uintptr_t)workspace) & 15) == 0,
                "tpspp_conv_bwd: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  size_t off[CW_COUNT], total;
  conv_offsets(d, off, &total);
  auto W = [&](int i) { return reinterpret_cast<float*>((char*)workspace + off[i]); };
  const int HoWo = d.Ho * d.Wo;
  // 1. g = gy * [y > 0], bias gradient
  relu_mask_bias_kernel<<<dim3(64, MB_SPLITS), 256, 0, st>>>(y, gy, W(CW_G), W(CW_BPART), d.B, HoWo, d.relu);
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());
  if (gb != nullptr) {
    bias_finalize_kernel<<<1, 64, 0, st>>>(W(CW_BPART), gb);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
  }
  // 2. data gradient: one stride-1 convolution over g (zero-inserted when the forward was strided) per 64 input channels
  if (gx != nullptr) {
    const int mode = d.KS == 3 ? CM_MIX : CM_TF32X3;
    WPrepLayer L[8];
    TPSPP_REQUIRE(d.slices <= 8, "too many input-channel slices");
    memset(L, 0, sizeof(L));
    const size_t per = conv_tc_wprep_floats(64, d.KS, 64);
    for (int s = 0; s < d.slices; ++s) {
      const int n = d.Cin - s * 64 < 64 ? d.Cin - s * 64 : 64;
      L[s].w = w; L[s].out = W(CW_WDG) + (size_t)s * per; L[s].Ctot = 64; L[s].taps = d.T; L[s].N = n; L[s].NT = 64; L[s].bf16 = mode;
      L[s].dg_cin = d.Cin; L[s].dg_ci0 = s * 64;
    }
    rc = conv_tc_prepare_weights(L, d.slices, st);
    if (rc != TPSPP_OK) return rc;
    for (int s = 0; s < d.slices; ++s) {
      const int n = d.Cin - s * 64 < 64 ? d.Cin - s * 64 : 64;
      ConvArgs a;
      memset(&a, 0, sizeof(a));
      a.src[0].ptr = W(CW_G); a.src[0].C = 64; a.src[0].H = d.Ho; a.src[0].W = d.Wo; a.src[0].uh = d.sh; a.src[0].uw = d.sw;
      a.src[1].H = a.src[1].W = a.src[1].uh = a.src[1].uw = 1; a.src[2] = a.src[1];
      a.zi = (d.sh == 2 || d.sw == 2) ? 1 : 0;
      a.out = gx + (size_t)s * 64 * d.H * d.W; a.out_cstride = d.Cin; a.Cout = n;
      a.B = d.B; a.Ho = d.H; a.Wo = d.W; a.Ctot = 64; a.sh = 1; a.sw = 1; a.pad = d.KS / 2;
      a.act = CONV_ACT_NONE; a.act_scale = 1.f;
      rc = run_conv_tc(d.KS, a, W(CW_WDG) + (size_t)s * per, 64, st, mode);
      if (rc != TPSPP_OK) return rc;
    }
  }
  // 3. weight gradient
  if (gw != nullptr) {
    WgradArgs wa{W(CW_G), x, W(CW_WPART), d.B, d.Cin, d.H, d.W, d.Ho, d.Wo, d.KS, d.sh, d.sw, d.splits};
    wgrad_kernel<<<dim3(d.T * d.slices, d.splits), 256, 0, st>>>(wa);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
    wgrad_reduce_kernel<<<64, 256, 0, st>>>(W(CW_WPART), gw, d.Cin, d.T, d.splits);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
  }
  return TPSPP_OK;
}
