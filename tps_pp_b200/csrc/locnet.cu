// Classical (RARE) localisation network on native kernels: what
// `TPSPreprocessor.LocalizationNetwork.forward` computes (reference preprocessor/tps_preprocessor.py:96-156) -- four
// conv3x3 + BatchNorm + ReLU blocks (C -> 64 -> 128 -> 256 -> 512; MaxPool 2x2 after the first three, AdaptiveAvgPool(1) after
// the last), localization_fc1 (512 -> 256, ReLU) and localization_fc2 (256 -> 2F) -> C' [B, F, 2].  It is 97 % of the
// classical preprocessor's time (1.87 GFLOP per 64x256 image against 0.4 MB of warp traffic).  Inference only (BatchNorm in
// eval mode, folded into the convolutions as in stage.cu).
//   block 1   C = 1 or 3 input channels: 27 (9) inputs per pixel -- CUDA cores in fp32, conv + BN + ReLU + 2x2 max-pool fused
//             (locnet_stem_kernel: thread = pooled pixel, 4 x 4 input patch and 4 pixels x 4 channels of accumulators in registers)
//   blocks 2-4 the head's tcgen05 engine (conv_tma_kernel, tf32 main term + bf16 corrections = fp32-level), 64 output channels
//             per launch written as a channel slice of the layer's output (ConvArgs::out_cstride)
//   pools     maxpool2_kernel, locnet_avgpool_kernel (one warp per plane); the two dense layers in locnet_fc_kernel (8 images per CTA)
#include "head.cuh"

#include <string.h>

namespace tpspp {

constexpr int LN_LAYERS = 4;
static const int kLocC[LN_LAYERS + 1] = {0, 64, 128, 256, 512};

// ---- BatchNorm folding for all four blocks: scale / bias per output channel ----
struct LocFoldArgs {
  const float *gamma[LN_LAYERS], *beta[LN_LAYERS], *mean[LN_LAYERS], *var[LN_LAYERS];
  float* fold;      // layer l: scale at fold + 2 * off[l], bias right behind it
};
__global__ void __launch_bounds__(256) locnet_fold_kernel(LocFoldArgs a) {
  const int l = blockIdx.y;
  const int C = l == 0 ? 64 : l == 1 ? 128 : l == 2 ? 256 : 512;
  const int off = l == 0 ? 0 : l == 1 ? 64 : l == 2 ? 192 : 448;
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= C) return;
  const double s = (double)__ldg(a.gamma[l] + n) / sqrt((double)__ldg(a.var[l] + n) + 1e-5);      // nn.BatchNorm2d eps
  a.fold[2 * off + n] = (float)s;
  a.fold[2 * off + C + n] = (float)((double)__ldg(a.beta[l] + n) - (double)__ldg(a.mean[l] + n) * s);
}

// ---- block 1: conv 3x3 (CIN -> 64, pad 1) + folded BN + ReLU + MaxPool 2x2.  Thread = pooled output pixel: its 4 x 4 input
//      patch per channel lives in registers, the 9 CIN x 64 weights (pre-multiplied by the BN scale) in shared memory as
//      [k][64]; four output channels per warp-uniform LDS.128 feed 16 FMAs (4 pixels x 4 channels) ----
struct LocStemArgs {
  const float *img, *w, *scale, *bias;
  float* out;       // [B, 64, H/2, W/2]
  int B, H, W;
};
template <int CIN>
__global__ void __launch_bounds__(128) locnet_stem_kernel(LocStemArgs a) {
  constexpr int KT = CIN * 9;
  __shared__ __align__(16) float ws[KT * 64];
  __shared__ __align__(16) float bs[64];
  for (int i = threadIdx.x; i < KT * 64; i += 128) {
    const int k = i >> 6, n = i & 63;
    ws[i] = __ldg(a.w + n * KT + k) * __ldg(a.scale + n);
  }
  if (threadIdx.x < 64) bs[threadIdx.x] = __ldg(a.bias + threadIdx.x);
  __syncthreads();
  const int Hp = a.H >> 1, Wp = a.W >> 1, HWp = Hp * Wp, HW = a.H * a.W;
  const long long total = (long long)a.B * HWp;
  for (long long p = (long long)blockIdx.x * 128 + threadIdx.x; p < total; p += (long long)gridDim.x * 128) {
    const int b = (int)(p / HWp), r = (int)(p - (long long)b * HWp);
    const int py = r / Wp, px = r - py * Wp;
    float v[CIN][4][4];
#pragma unroll
    for (int c = 0; c < CIN; ++c)
#pragma unroll
      for (int dy = 0; dy < 4; ++dy) {
        const int yy = 2 * py + dy - 1;
        const float* row = a.img + ((size_t)b * CIN + c) * HW + (size_t)yy * a.W;
#pragma unroll
        for (int dx = 0; dx < 4; ++dx) {
          const int xx = 2 * px + dx - 1;
          v[c][dy][dx] = (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) ? __ldg(row + xx) : 0.f;
        }
      }
    float* po = a.out + (size_t)b * 64 * HWp + r;
    const float4* ws4 = reinterpret_cast<const float4*>(ws);
    const float4* bs4 = reinterpret_cast<const float4*>(bs);
#pragma unroll 1
    for (int n4 = 0; n4 < 16; ++n4) {
      const float4 b4 = bs4[n4];
      float4 acc[4] = {b4, b4, b4, b4};                       // the 2 x 2 conv outputs under this pooled pixel
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4 w4 = ws4[(c * 9 + ky * 3 + kx) * 16 + n4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float x = v[c][(q >> 1) + ky][(q & 1) + kx];
              acc[q].x = fmaf(x, w4.x, acc[q].x); acc[q].y = fmaf(x, w4.y, acc[q].y);
              acc[q].z = fmaf(x, w4.z, acc[q].z); acc[q].w = fmaf(x, w4.w, acc[q].w);
            }
          }
      // relu(max) == max(relu): the reference applies ReLU, then the pool
      const float m0 = fmaxf(fmaxf(fmaxf(acc[0].x, acc[1].x), fmaxf(acc[2].x, acc[3].x)), 0.f);
      const float m1 = fmaxf(fmaxf(fmaxf(acc[0].y, acc[1].y), fmaxf(acc[2].y, acc[3].y)), 0.f);
      const float m2 = fmaxf(fmaxf(fmaxf(acc[0].z, acc[1].z), fmaxf(acc[2].z, acc[3].z)), 0.f);
      const float m3 = fmaxf(fmaxf(fmaxf(acc[0].w, acc[1].w), fmaxf(acc[2].w, acc[3].w)), 0.f);
      po[(size_t)(4 * n4) * HWp] = m0;
      po[(size_t)(4 * n4 + 1) * HWp] = m1;
      po[(size_t)(4 * n4 + 2) * HWp] = m2;
      po[(size_t)(4 * n4 + 3) * HWp] = m3;
    }
  }
}

// ---- MaxPool2d(2, 2) over [planes, H, W] -> [planes, H/2, W/2]: one thread per two pooled pixels (16-byte row reads) ----
__global__ void __launch_bounds__(256) maxpool2_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int H,
                                                       int W) {
  const int Hp = H >> 1, Wp2 = W >> 2;             // pairs of pooled pixels per pooled row
  const long long total = planes * Hp * Wp2;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long pl = i / ((long long)Hp * Wp2);
    const int r = (int)(i - pl * Hp * Wp2), y = r / Wp2, x2 = r - y * Wp2;
    const float4* q = reinterpret_cast<const float4*>(in + (pl * H + 2 * y) * W) + x2;
    const float4 t = __ldg(q), u = __ldg(q + (W >> 2));
    *reinterpret_cast<float2*>(out + (pl * Hp + y) * (W >> 1) + 2 * x2) =
        make_float2(fmaxf(fmaxf(t.x, t.y), fmaxf(u.x, u.y)), fmaxf(fmaxf(t.z, t.w), fmaxf(u.z, u.w)));
  }
}

// ---- AdaptiveAvgPool2d(1) + localization_fc1 (512 -> 256, ReLU) + localization_fc2 (256 -> 2F): one CTA per image ----
struct LocFcArgs {
  const float *feat;                    // pooled features [B, 512]
  const float *w1, *b1, *w2, *b2;       // [256, 512], [256], [2F, 256], [2F]
  float* c_prime;                       // [B, 2F]
  int hw, nout;
};
// AdaptiveAvgPool2d(1): one warp per (image, channel) plane, fixed lane assignment + shuffle tree (deterministic)
__global__ void __launch_bounds__(256) locnet_avgpool_kernel(const float* __restrict__ feat, float* __restrict__ pooled, long long planes,
                                                             int hw) {
  const int lane = threadIdx.x & 31;
  const long long pl = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pl >= planes) return;
  const float4* q = reinterpret_cast<const float4*>(feat + pl * hw);
  float s = 0.f;
  for (int k = lane; k < (hw >> 2); k += 32) {
    const float4 v = __ldg(q + k);
    s += (v.x + v.y) + (v.z + v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) pooled[pl] = s / (float)hw;
}
constexpr int FC_IMGS = 8;        // images per CTA: every weight row read from L2 feeds eight dot products
__global__ void __launch_bounds__(256) locnet_fc_kernel(LocFcArgs a, int B) {
  __shared__ __align__(16) float pooled[FC_IMGS][512];
  __shared__ float hid[FC_IMGS][256];
  const int b0 = blockIdx.x * FC_IMGS, tid = threadIdx.x;
  const int nb = min(FC_IMGS, B - b0);
  for (int u = tid; u < FC_IMGS * 512; u += 256) {
    const int i = u >> 9;
    pooled[i][u & 511] = i < nb ? __ldg(a.feat + (size_t)b0 * 512 + u) : 0.f;      // a.feat = pooled features [B, 512]
  }
  __syncthreads();
  {
    const float4* w = reinterpret_cast<const float4*>(a.w1 + (size_t)tid * 512);
    float acc[FC_IMGS];
#pragma unroll
    for (int i = 0; i < FC_IMGS; ++i) acc[i] = 0.f;
#pragma unroll 2
    for (int k = 0; k < 128; ++k) {
      const float4 w4 = __ldg(w + k);
#pragma unroll
      for (int i = 0; i < FC_IMGS; ++i) {
        const float4 p4 = *reinterpret_cast<const float4*>(&pooled[i][4 * k]);     // warp-uniform: broadcast
        acc[i] = fmaf(w4.x, p4.x, acc[i]); acc[i] = fmaf(w4.y, p4.y, acc[i]);
        acc[i] = fmaf(w4.z, p4.z, acc[i]); acc[i] = fmaf(w4.w, p4.w, acc[i]);
      }
    }
    const float bb = __ldg(a.b1 + tid);
#pragma unroll
    for (int i = 0; i < FC_IMGS; ++i) hid[i][tid] = fmaxf(acc[i] + bb, 0.f);
  }
  __syncthreads();
  for (int o = tid; o < a.nout * FC_IMGS; o += 256) {
    const int i = o / a.nout, n = o - i * a.nout;
    if (i >= nb) continue;
    const float* w = a.w2 + (size_t)n * 256;
    float acc = 0.f;
    for (int k = 0; k < 256; ++k) acc = fmaf(__ldg(w + k), hid[i][k], acc);
    a.c_prime[(size_t)(b0 + i) * a.nout + n] = acc + __ldg(a.b2 + n);
  }
}

struct LocDims { int B, C, H, W, F; };
// a feature map the tensor-core convolution can tile into 128-pixel rectangles of one image (conv_tma_plan)
static bool loc_level_ok(int h, int w) {
  if (!(w == 16 || w == 32 || (w >= 64 && w % 64 == 0))) return false;
  const int tw = w < 64 ? w : 64;
  return h % (128 / tw) == 0;
}
static int loc_dims(const tpspp_locnet_cfg* c, LocDims* d) {
  TPSPP_REQUIRE(c != nullptr, "locnet cfg is NULL");
  TPSPP_REQUIRE(c->batch >= 0, "batch must be >= 0");
  TPSPP_REQUIRE(c->channels == 1 || c->channels == 3, "the localisation network takes 1 or 3 image channels (got %d)", c->channels);
  TPSPP_REQUIRE(c->num_fiducial > 0 && c->num_fiducial <= 128, "num_fiducial must be in 1..128");
  TPSPP_REQUIRE(c->height >= 8 && c->height % 8 == 0 && c->width >= 8 && c->width % 8 == 0, "image size must be a multiple of 8 (got %d x %d)",
                c->height, c->width);
  // the tensor-core convolutions tile every pooled feature map into 128-pixel rectangles of one image: widths 16, 32 or a
  // multiple of 64 with matching heights (64x256, 64x128, 32x256, 128x512, ...; not the 32x100 default of the recogniser configs)
  TPSPP_REQUIRE(loc_level_ok(c->height / 2, c->width / 2) && loc_level_ok(c->height / 4, c->width / 4) &&
                    loc_level_ok(c->height / 8, c->width / 8),
                "image size %d x %d is not supported by the native localisation network", c->height, c->width);
  d->B = c->batch; d->C = c->channels; d->H = c->height; d->W = c->width; d->F = c->num_fiducial;
  return TPSPP_OK;
}
enum { LW_FOLD = 0, LW_WPREP, LW_P1, LW_C2, LW_P2, LW_C3, LW_P3, LW_C4, LW_COUNT };
static void loc_offsets(const LocDims& d, size_t* off, size_t* total) {
  const size_t B = d.B, s2 = (size_t)(d.H / 2) * (d.W / 2), s4 = s2 / 4, s8 = s4 / 4;
  size_t sz[LW_COUNT];
  sz[LW_FOLD] = 2 * 960;
  sz[LW_WPREP] = conv_tc_wprep_floats(64, 3, 128) + conv_tc_wprep_floats(128, 3, 256) + conv_tc_wprep_floats(256, 3, 512);
  sz[LW_P1] = B * 64 * s2; sz[LW_C2] = B * 128 * s2; sz[LW_P2] = B * 128 * s4; sz[LW_C3] = B * 256 * s4;
  sz[LW_P3] = B * 256 * s8; sz[LW_C4] = B * 512 * s8;
  size_t cur = 0;
  for (int i = 0; i < LW_COUNT; ++i) {
    off[i] = cur;
    cur += (sz[i] * sizeof(float) + 255) / 256 * 256;
  }
  *total = cur + 256;
}

}  // namespace tpspp

using namespace tpspp;

extern "C" size_t tpspp_locnet_workspace_bytes(const tpspp_locnet_cfg* cfg) {
  LocDims d;
  if (loc_dims(cfg, &d) != TPSPP_OK) return 0;
  size_t off[LW_COUNT], total;
  loc_offsets(d, off, &total);
  return total;
}

extern "C" int tpspp_locnet_fwd(const tpspp_locnet_cfg* cfg, const float* img, const float* const* P, float* c_prime,
                                void* workspace, tpspp_stream_t stream) {
  reset_launch_count();
  LocDims d;
  int rc = loc_dims(cfg, &d);
  if (rc != TPSPP_OK) return rc;
  if (d.B == 0) return TPSPP_OK;
  TPSPP_REQUIRE(img && P && c_prime && workspace, "tpspp_locnet_fwd: null pointer");
  TPSPP_REQUIRE(((uintptr_t)workspace & 255) == 0, "tpspp_locnet_fwd: workspace must be 256-byte aligned");
  for (int i = 0; i < TPSPP_LP_COUNT; ++i) {
    TPSPP_REQUIRE(P[i] != nullptr, "tpspp_locnet_fwd: params[%d] is NULL", i);
    TPSPP_REQUIRE(((uintptr_t)P[i] & 3) == 0, "tpspp_locnet_fwd: params[%d] is not a float pointer", i);
  }
  TPSPP_REQUIRE(((uintptr_t)P[TPSPP_LP_FC1_W] & 15) == 0, "tpspp_locnet_fwd: localization_fc1 weight must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  size_t off[LW_COUNT], total;
  loc_offsets(d, off, &total);
  auto W = [&](int i) { return reinterpret_cast<float*>((char*)workspace + off[i]); };
  float* fold = W(LW_FOLD);
  static const int foff[LN_LAYERS] = {0, 64, 192, 448};
  auto scale_of = [&](int l) { return fold + 2 * foff[l]; };
  auto bias_of = [&](int l) { return fold + 2 * foff[l] + kLocC[l + 1]; };
  const float* wimg[LN_LAYERS] = {nullptr, nullptr, nullptr, nullptr};
  {
    float* cur = W(LW_WPREP);
    for (int l = 1; l < LN_LAYERS; ++l) {
      wimg[l] = cur;
      cur += conv_tc_wprep_floats(kLocC[l], 3, kLocC[l + 1]);
    }
  }
  if (!(cfg->flags & TPSPP_HEAD_FLAG_WEIGHTS_CACHED)) {
    LocFoldArgs fa;
    for (int l = 0; l < LN_LAYERS; ++l) {
      const int base = 5 * l;         // conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var
      fa.gamma[l] = P[base + 1]; fa.beta[l] = P[base + 2]; fa.mean[l] = P[base + 3]; fa.var[l] = P[base + 4];
    }
    fa.fold = fold;
    locnet_fold_kernel<<<dim3(2, LN_LAYERS), 256, 0, st>>>(fa);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
    // one 64-row weight image per 64 output channels of blocks 2-4 (the engine's column tile), BN scale folded in
    WPrepLayer L[WPREP_MAX_LAYERS];
    int nl = 0;
    for (int l = 1; l < LN_LAYERS; ++l) {
      const int Cin = kLocC[l], Cout = kLocC[l + 1];
      const size_t per = conv_tc_wprep_floats(Cin, 3, 64);
      for (int s = 0; s < Cout / 64; ++s) {
        memset(&L[nl], 0, sizeof(L[nl]));
        L[nl].w = P[5 * l] + (size_t)s * 64 * Cin * 9; L[nl].out = const_cast<float*>(wimg[l]) + (size_t)s * per;
        L[nl].Ctot = Cin; L[nl].taps = 9; L[nl].N = 64; L[nl].NT = 64; L[nl].bf16 = CM_MIX; L[nl].scale = scale_of(l) + s * 64;
        ++nl;
      }
    }
    rc = conv_tc_prepare_weights(L, nl, st);          // 2 + 4 + 8 = 14 images
    if (rc != TPSPP_OK) return rc;
  }
  // block 1
  {
    LocStemArgs sa{img, P[0], scale_of(0), bias_of(0), W(LW_P1), d.B, d.H, d.W};
    const long long px = (long long)d.B * (d.H / 2) * (d.W / 2);
    long long blocks = (px + 127) / 128;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    if (d.C == 1) locnet_stem_kernel<1><<<(unsigned)blocks, 128, 0, st>>>(sa);
    else locnet_stem_kernel<3><<<(unsigned)blocks, 128, 0, st>>>(sa);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
  }
  // blocks 2-4: relu(conv(src) * s + b), 64 output channels per launch
  auto conv_block = [&](int l, const float* src, int Hin, int Win, float* out) -> int {
    const int Cin = kLocC[l], Cout = kLocC[l + 1];
    const size_t per = conv_tc_wprep_floats(Cin, 3, 64);
    for (int s = 0; s < Cout / 64; ++s) {
      ConvArgs a;
      memset(&a, 0, sizeof(a));
      a.src[0].ptr = src; a.src[0].C = Cin; a.src[0].H = Hin; a.src[0].W = Win; a.src[0].uh = 1; a.src[0].uw = 1;
      a.src[1].H = a.src[1].W = a.src[1].uh = a.src[1].uw = 1; a.src[2] = a.src[1];
      a.weight = P[5 * l]; a.bias = bias_of(l) + s * 64; a.out = out + (size_t)s * 64 * Hin * Win; a.out_cstride = Cout;
      a.B = d.B; a.Ho = Hin; a.Wo = Win; a.Ctot = Cin; a.sh = 1; a.sw = 1; a.pad = 1;
      a.act = CONV_ACT_RELU; a.act_scale = 1.f; a.Cout = 64;
      const int r = run_conv_tc(3, a, wimg[l] + (size_t)s * per, 64, st, CM_MIX);
      if (r != TPSPP_OK) return r;
    }
    return TPSPP_OK;
  };
  auto pool = [&](const float* in, float* out, int C, int Hin, int Win) -> int {
    const long long planes = (long long)d.B * C;
    const long long tot = planes * (Hin / 2) * (Win / 4);
    maxpool2_kernel<<<(unsigned)min((tot + 255) / 256, 32LL * sm_count()), 256, 0, st>>>(in, out, planes, Hin, Win);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
    return TPSPP_OK;
  };
  const int H2 = d.H / 2, W2 = d.W / 2, H4 = d.H / 4, W4 = d.W / 4, H8 = d.H / 8, W8 = d.W / 8;
  rc = conv_block(1, W(LW_P1), H2, W2, W(LW_C2)); if (rc != TPSPP_OK) return rc;
  rc = pool(W(LW_C2), W(LW_P2), 128, H2, W2);     if (rc != TPSPP_OK) return rc;
  rc = conv_block(2, W(LW_P2), H4, W4, W(LW_C3)); if (rc != TPSPP_OK) return rc;
  rc = pool(W(LW_C3), W(LW_P3), 256, H4, W4);     if (rc != TPSPP_OK) return rc;
  rc = conv_block(3, W(LW_P3), H8, W8, W(LW_C4)); if (rc != TPSPP_OK) return rc;
  {
    // the pooled [B, 512] features reuse the (dead by now) P3 buffer
    const long long planes = (long long)d.B * 512;
    locnet_avgpool_kernel<<<(unsigned)((planes + 7) / 8), 256, 0, st>>>(W(LW_C4), W(LW_P3), planes, H8 * W8);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
    LocFcArgs fa{W(LW_P3), P[TPSPP_LP_FC1_W], P[TPSPP_LP_FC1_B], P[TPSPP_LP_FC2_W], P[TPSPP_LP_FC2_B], c_prime, H8 * W8, 2 * d.F};
    locnet_fc_kernel<<<(d.B + FC_IMGS - 1) / FC_IMGS, 256, 0, st>>>(fa, d.B);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
  }
  return TPSPP_OK;
}
