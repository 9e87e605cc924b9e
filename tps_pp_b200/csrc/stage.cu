// Backbone stage in front of the rectifier: what ResNetABI_v2_large.forward computes before it
// calls tpsnet(x, outs) -- reference backbones/resnet_v2_large.py:131-135 (stem), :109-129 (_make_layer), :176-191 (forward),
// block = layers/conv_layer.py:12-33 on top of mmcv's BasicBlock: relu(bn1(conv1x1 x)) -> bn2(conv3x3 .) -> + identity -> relu,
// identity through `downsample` (1x1 stride-s convolution + BN) in the first block of layer2.  Inference only (BatchNorm
// in eval mode): every BN is folded into the convolution in front of it,
//     y = conv(x) * s + (beta + (conv_bias - mean) * s),   s = gamma / sqrt(var + eps),
// s goes into the tensor-core weight image (wprep_kernel's per-row scale), the rest is the epilogue's bias.  With the
// stage native the host boundary of the rectifier moves from three fp32 feature maps (1.31 MB per image) to the image
// itself (49 KB): o0, o1 and x are written straight into the buffers tpspp_head_fwd / tpspp_warp_fwd read.
//   stem          3 -> 32, 3x3: 27 inputs per pixel, CUDA cores in fp32 (stem_kernel; 7 MFLOP per image)
//   15 more convs 1x1 / 3x3, 32 or 64 channels: the head's tcgen05 engine (conv_tma_kernel / conv_ts_kernel, head_tc.cu):
//                 split-fp32 operands (tf32 main term + bf16 corrections), 32-channel layers on 32-column MMAs (half the weight
//                 bytes per chunk), residual added before the ReLU in the epilogue (ConvArgs::skip_pre)
#include "head.cuh"

#include <string.h>

namespace tpspp {

// ---- BatchNorm folding: one block per convolution ----
struct FoldItem {
  const float *conv_bias, *gamma, *beta, *mean, *var;   // conv_bias may be null
  float *scale, *bias;                                  // [C] each
  int C;
};
constexpr int STAGE_CONVS = 16;
struct FoldArgs { FoldItem it[STAGE_CONVS]; };
__global__ void __launch_bounds__(64) bnfold_kernel(FoldArgs a) {
  const FoldItem f = a.it[blockIdx.x];
  const int n = threadIdx.x;
  if (n >= f.C) return;
  // fp64: the fold itself must not add an error the unfolded fp32 reference does not have
  const double s = (double)__ldg(f.gamma + n) / sqrt((double)__ldg(f.var + n) + 1e-5);     // nn.BatchNorm2d eps
  const double cb = f.conv_bias != nullptr ? (double)__ldg(f.conv_bias + n) : 0.0;
  f.scale[n] = (float)s;
  f.bias[n] = (float)((double)__ldg(f.beta + n) + (cb - (double)__ldg(f.mean + n)) * s);
}

// ---- stem: conv 3x3 (3 -> 32, pad 1) + folded BN + ReLU.  Thread = output pixel (lanes walk the row: coalesced loads and
//      stores), all 32 output channels in registers, the 864 weights (pre-multiplied by the BN scale) in shared memory ----
struct StemArgs {
  const float *img, *w, *scale, *bias;
  float* out;
  int B, H, W;
};
__global__ void __launch_bounds__(256) stem_kernel(StemArgs a) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float ws[27 * 32];      // [k = cin*9 + tap][cout]
  __shared__ __align__(16) float bs[32];
  for (int i = threadIdx.x; i < 27 * 32; i += 256) {
    const int k = i >> 5, n = i & 31;
    ws[i] = __ldg(a.w + n * 27 + k) * __ldg(a.scale + n);
  }
  if (threadIdx.x < 32) bs[threadIdx.x] = __ldg(a.bias + threadIdx.x);
  __syncthreads();
  const int HW = a.H * a.W;
  const long long total = (long long)a.B * HW;
  for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < total; p += (long long)gridDim.x * 256) {
    const int b = (int)(p / HW), r = (int)(p - (long long)b * HW);
    const int y = r / a.W, x = r - y * a.W;
    float v[27];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int yy = y + dy - 1, xx = x + dx - 1;
          const bool ok = yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
          v[c * 9 + dy * 3 + dx] = ok ? __ldg(a.img + ((size_t)b * 3 + c) * HW + (size_t)yy * a.W + xx) : 0.f;
        }
    float* po = a.out + (size_t)b * 32 * HW + r;
    // four output channels per shared-memory load (warp-uniform LDS.128): with one LDS per FMA the kernel was bound by the
    // shared-memory pipe (97 us at B = 256 against a 23 us HBM floor)
    const float4* ws4 = reinterpret_cast<const float4*>(ws);
    const float4* bs4 = reinterpret_cast<const float4*>(bs);
#pragma unroll 1
    for (int n4 = 0; n4 < 8; ++n4) {
      float4 acc = bs4[n4];
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        const float4 w4 = ws4[k * 8 + n4];
        acc.x = fmaf(v[k], w4.x, acc.x); acc.y = fmaf(v[k], w4.y, acc.y);
        acc.z = fmaf(v[k], w4.z, acc.z); acc.w = fmaf(v[k], w4.w, acc.w);
      }
      po[(size_t)(4 * n4) * HW] = fmaxf(acc.x, 0.f);
      po[(size_t)(4 * n4 + 1) * HW] = fmaxf(acc.y, 0.f);
      po[(size_t)(4 * n4 + 2) * HW] = fmaxf(acc.z, 0.f);
      po[(size_t)(4 * n4 + 3) * HW] = fmaxf(acc.w, 0.f);
    }
  }
}

// ---- the 15 tensor-core convolutions, in launch order ----
struct StageConv {
  int w_idx;        // index of the convolution weight in the params table; its BN follows at w_idx + 1 .. + 4
  int Cin, Cout, KS, stride;
};
// params table (include/tpspp.h TPSPP_SP_*): conv1.weight, conv1.bias, bn1 x4, then per block conv1.weight, bn1 x4,
// conv2.weight, bn2 x4 (+ downsample.0.weight, downsample.1 x4 in layer2.0)
static const StageConv kStage[STAGE_CONVS - 1] = {
    {6, 32, 32, 1, 1},  {11, 32, 32, 3, 1},                       // layer1.0
    {16, 32, 32, 1, 1}, {21, 32, 32, 3, 1},                       // layer1.1
    {26, 32, 32, 1, 1}, {31, 32, 32, 3, 1},                       // layer1.2
    {36, 32, 64, 1, 1}, {41, 64, 64, 3, 2}, {46, 32, 64, 1, 2},   // layer2.0: conv1, conv2 (stride 2), downsample (stride 2)
    {51, 64, 64, 1, 1}, {56, 64, 64, 3, 1},                       // layer2.1
    {61, 64, 64, 1, 1}, {66, 64, 64, 3, 1},                       // layer2.2
    {71, 64, 64, 1, 1}, {76, 64, 64, 3, 1}};                      // layer2.3
static_assert(TPSPP_SP_COUNT == 81, "stage parameter table");

struct StageDims { int B, H, W, h, w; };
static int stage_dims(const tpspp_stage_cfg* c, StageDims* d) {
  TPSPP_REQUIRE(c != nullptr, "stage cfg is NULL");
  TPSPP_REQUIRE(c->batch >= 0, "batch must be >= 0");
  TPSPP_REQUIRE(c->height >= 4 && c->height % 4 == 0, "image height must be a multiple of 4 (got %d)", c->height);
  TPSPP_REQUIRE(c->width == 128, "the stage is built for 128-pixel-wide images (TPS_PP needs 64-wide feature maps, SURVEY F4); got %d",
                c->width);
  TPSPP_REQUIRE(c->precision == TPSPP_HEAD_TC, "the backbone stage runs on the tensor-core engine only (precision TPSPP_HEAD_TC)");
  d->B = c->batch; d->H = c->height; d->W = c->width; d->h = c->height / 2; d->w = c->width / 2;
  return TPSPP_OK;
}

enum { SW_FOLD = 0, SW_WPREP, SW_T1, SW_YA, SW_YB, SW_T2, SW_IDN, SW_T3, SW_ZA, SW_ZB, SW_COUNT };
static void stage_offsets(const StageDims& d, size_t* off, size_t* total) {
  const size_t B = d.B, big32 = B * 32 * d.H * d.W, big64 = B * 64 * d.H * d.W, small64 = B * 64 * d.h * d.w;
  size_t sz[SW_COUNT];
  sz[SW_FOLD] = STAGE_CONVS * 128;
  size_t wp = 0;
  for (int i = 0; i < STAGE_CONVS - 1; ++i) wp += conv_tc_wprep_floats(kStage[i].Cin, kStage[i].KS, kStage[i].Cout);
  sz[SW_WPREP] = wp;
  sz[SW_T1] = big32; sz[SW_YA] = big32; sz[SW_YB] = big32; sz[SW_T2] = big64;
  sz[SW_IDN] = small64; sz[SW_T3] = small64; sz[SW_ZA] = small64; sz[SW_ZB] = small64;
  size_t cur = 0;
  for (int i = 0; i < SW_COUNT; ++i) {
    off[i] = cur;
    cur += (sz[i] * sizeof(float) + 255) / 256 * 256;
  }
  *total = cur + 256;
}

}  // namespace tpspp

using namespace tpspp;

extern "C" size_t tpspp_stage_workspace_bytes(const tpspp_stage_cfg* cfg) {
  StageDims d;
  if (stage_dims(cfg, &d) != TPSPP_OK) return 0;
  size_t off[SW_COUNT], total;
  stage_offsets(d, off, &total);
  return total;
}

extern "C" int tpspp_stage_fwd(const tpspp_stage_cfg* cfg, const float* img, const float* const* P, float* o0, float* o1,
                               float* x, void* workspace, tpspp_stream_t stream) {
  reset_launch_count();
  prof_begin((cudaStream_t)stream);
  StageDims d;
  int rc = stage_dims(cfg, &d);
  if (rc != TPSPP_OK) return rc;
  if (d.B == 0) return TPSPP_OK;
  TPSPP_REQUIRE(img && P && o0 && o1 && x && workspace, "tpspp_stage_fwd: null pointer");
  TPSPP_REQUIRE(((uintptr_t)workspace & 255) == 0, "tpspp_stage_fwd: workspace must be 256-byte aligned");
  TPSPP_REQUIRE((((uintptr_t)img | (uintptr_t)o0 | (uintptr_t)o1 | (uintptr_t)x) & 15) == 0,
                "tpspp_stage_fwd: image and output buffers must be 16-byte aligned");
  for (int i = 0; i < TPSPP_SP_COUNT; ++i) {
    TPSPP_REQUIRE(P[i] != nullptr, "tpspp_stage_fwd: params[%d] is NULL", i);
    TPSPP_REQUIRE(((uintptr_t)P[i] & 3) == 0, "tpspp_stage_fwd: params[%d] is not a float pointer", i);
  }
  cudaStream_t st = (cudaStream_t)stream;
  struct PdlScope { PdlScope() { pdl_scope(true); } ~PdlScope() { pdl_scope(false); } } pdl_guard;
  size_t off[SW_COUNT], total;
  stage_offsets(d, off, &total);
  auto W = [&](int i) { return reinterpret_cast<float*>((char*)workspace + off[i]); };
  float* fold = W(SW_FOLD);                                   // conv i: scale at fold + 128 i, bias at + 64
  const float* wp[STAGE_CONVS - 1];
  {
    float* cur = W(SW_WPREP);
    for (int i = 0; i < STAGE_CONVS - 1; ++i) {
      wp[i] = cur;
      cur += conv_tc_wprep_floats(kStage[i].Cin, kStage[i].KS, kStage[i].Cout);
    }
  }
  if (!(cfg->flags & TPSPP_HEAD_FLAG_WEIGHTS_CACHED)) {
    FoldArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.it[0] = FoldItem{P[TPSPP_SP_CONV1_B], P[TPSPP_SP_BN1_W], P[TPSPP_SP_BN1_B], P[TPSPP_SP_BN1_MEAN], P[TPSPP_SP_BN1_VAR],
                        fold, fold + 64, 32};
    for (int i = 0; i < STAGE_CONVS - 1; ++i) {
      const int w = kStage[i].w_idx;
      fa.it[i + 1] = FoldItem{nullptr, P[w + 1], P[w + 2], P[w + 3], P[w + 4], fold + 128 * (i + 1), fold + 128 * (i + 1) + 64,
                              kStage[i].Cout};
    }
    bnfold_kernel<<<STAGE_CONVS, 64, 0, st>>>(fa);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
    WPrepLayer L[STAGE_CONVS - 1];
    for (int i = 0; i < STAGE_CONVS - 1; ++i) {
      L[i].w = P[kStage[i].w_idx]; L[i].out = const_cast<float*>(wp[i]); L[i].Ctot = kStage[i].Cin;
      L[i].taps = kStage[i].KS * kStage[i].KS; L[i].N = kStage[i].Cout; L[i].NT = kStage[i].Cout;     // 32-row images for the 32-channel layers
      L[i].bf16 = kStage[i].KS == 3 ? CM_MIX : CM_TF32X3;
      L[i].scale = fold + 128 * (i + 1); L[i].dg_cin = 0; L[i].dg_ci0 = 0;
    }
    rc = conv_tc_prepare_weights(L, STAGE_CONVS - 1, st);
    if (rc != TPSPP_OK) return rc;
  }
  // stem (resnet_v2_large.py:176-178)
  {
    StemArgs sa{img, P[TPSPP_SP_CONV1_W], fold, fold + 64, o0, d.B, d.H, d.W};
    const long long total_px = (long long)d.B * d.H * d.W;
    long long blocks = (total_px + 255) / 256;
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    launch_k(stem_kernel, dim3((unsigned)blocks), dim3(256), 0, st, sa);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
  }
  // one tensor-core convolution of the table: relu(conv(src) * s + b [+ identity])
  auto conv = [&](int i, const float* src, int Hin, int Win, const float* identity, float* out, bool relu) -> int {
    const StageConv& c = kStage[i];
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.src[0].ptr = src; a.src[0].C = c.Cin; a.src[0].H = Hin; a.src[0].W = Win; a.src[0].uh = 1; a.src[0].uw = 1; a.src[0].nhwc = 0;
    a.src[1].H = a.src[1].W = a.src[1].uh = a.src[1].uw = 1; a.src[2] = a.src[1];
    a.weight = P[c.w_idx]; a.bias = fold + 128 * (i + 1) + 64; a.skip = identity; a.skip_pre = 1; a.out = out;
    a.B = d.B; a.Ho = Hin / c.stride; a.Wo = Win / c.stride; a.Ctot = c.Cin; a.sh = c.stride; a.sw = c.stride;
    a.pad = c.KS == 3 ? 1 : 0; a.out_nhwc = 0; a.act = relu ? CONV_ACT_RELU : CONV_ACT_NONE; a.act_scale = 1.f;
    a.Cout = c.Cout; a.wimg_stride = 0;
    return run_conv_tc(c.KS, a, wp[i], c.Cout, st, c.KS == 3 ? CM_MIX : CM_TF32X3);
  };
#define CONV(...) do { rc = conv(__VA_ARGS__); if (rc != TPSPP_OK) return rc; } while (0)
  const int H = d.H, Wd = d.W, h = d.h, w = d.w;
  // layer1 (three blocks, 32 channels, full resolution); its output is outs[1]
  CONV(0, o0, H, Wd, nullptr, W(SW_T1), true);        CONV(1, W(SW_T1), H, Wd, o0, W(SW_YA), true);
  CONV(2, W(SW_YA), H, Wd, nullptr, W(SW_T1), true);  CONV(3, W(SW_T1), H, Wd, W(SW_YA), W(SW_YB), true);
  CONV(4, W(SW_YB), H, Wd, nullptr, W(SW_T1), true);  CONV(5, W(SW_T1), H, Wd, W(SW_YB), o1, true);
  // layer2.0: 1x1 to 64 channels, identity through the stride-2 1x1 convolution + BN, 3x3 stride 2
  CONV(6, o1, H, Wd, nullptr, W(SW_T2), true);
  CONV(8, o1, H, Wd, nullptr, W(SW_IDN), false);
  CONV(7, W(SW_T2), H, Wd, W(SW_IDN), W(SW_ZA), true);
  // layer2.1 - layer2.3 at half resolution; the last block writes x
  CONV(9, W(SW_ZA), h, w, nullptr, W(SW_T3), true);   CONV(10, W(SW_T3), h, w, W(SW_ZA), W(SW_ZB), true);
  CONV(11, W(SW_ZB), h, w, nullptr, W(SW_T3), true);  CONV(12, W(SW_T3), h, w, W(SW_ZB), W(SW_ZA), true);
  CONV(13, W(SW_ZA), h, w, nullptr, W(SW_T3), true);  CONV(14, W(SW_T3), h, w, W(SW_ZA), x, true);
#undef CONV
  return TPSPP_OK;
}
