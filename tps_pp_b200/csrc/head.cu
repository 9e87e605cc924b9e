// Control-point attention head of TPS_PP, forward, fp32 on CUDA cores (parity mode).
//
// Reference being replaced (backbones/tps_pp/tps_pp.py unless noted):
//   down0/1/2, down0_1/1_1, grid()/down_feat        :538-548,560-562,581-585   -> conv_ffma_kernel
//   Encoder_Decoder_Feature_Extractor.forward       :156-169                   -> conv_ffma_kernel (+skip)
//   CBAM / ChannelAttention / SpatialAttention      :27-82                     -> cbam_kernel
//   DGAB.forward / DGAB_Block.forward (DGAB.py:39-55,74-77), LayerNorm(H,W)    -> dgab_warp_kernel
//   Mlp.forward (DGAB.py:17-23) + residual                                     -> dgab_mlp_kernel
//   localization_fc1/fc2 -> C' (:321-323), p_linear (:305)                     -> loc_p1_kernel
//   feat_linear + atten_score = tanh(f p1^T * 64^-0.5) (:293-312)              -> score_kernel
//
// All tensors NCHW fp32 exactly as the reference holds them.  Every conv here has 64 output
// channels; conv_ffma_kernel is an implicit GEMM (M = B*Ho*Wo pixels, N = 64, K = Cin*KH*KW) with
// up to three concatenated input tensors, optional nearest up-sampling of each input
// (nn.Upsample in the decoder / grid()), stride, zero padding, bias + ReLU and the decoder's
// skip add fused -- torch.cat / Upsample never materialise.
#include "head.cuh"

#include <string.h>

namespace tpspp {

// =====================================================================================
// implicit-GEMM convolution, 64 output channels, fp32 FFMA
// =====================================================================================
constexpr int CV_TM = 128;   // output pixels per CTA
constexpr int CV_KC = 32;    // K chunk
constexpr int CV_WLD = 68;   // padded leading dim of the weight chunk (floats)
constexpr int CV_SMEM = (2 * CV_KC * CV_TM + 2 * CV_KC * CV_WLD) * 4;

template <int KS>
__global__ void __launch_bounds__(256, 2) conv_ffma_kernel(ConvArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                           // [2][CV_KC][CV_TM]
  float* Ws = smem + 2 * CV_KC * CV_TM;       // [2][CV_KC][CV_WLD]
  constexpr int TAPS = KS * KS;
  const int tid = threadIdx.x;
  const int HoWo = a.Ho * a.Wo;
  const long long Mtot = (long long)a.B * HoWo;
  const long long m_base = (long long)blockIdx.x * CV_TM;
  const int Ktot = a.Ctot * TAPS;
  const int nchunks = (Ktot + CV_KC - 1) / CV_KC;

  // loader role: one output pixel per thread (lp), k rows lk0, lk0+2, ...
  const int lp = tid & (CV_TM - 1), lk0 = tid >> 7;
  const long long lm = m_base + lp;
  const bool lvalid = lm < Mtot;
  int lb = 0, iy0 = 0, ix0 = 0;
  if (lvalid) {
    lb = (int)(lm / HoWo);
    const int r = (int)(lm - (long long)lb * HoWo);
    const int oy = r / a.Wo, ox = r - oy * a.Wo;
    iy0 = oy * a.sh - a.pad;
    ix0 = ox * a.sw - a.pad;
  }
  const int wk = tid & 31, wc0 = tid >> 5;

  float ra[16], rw[8];
  auto load_chunk = [&](int chunk) {
    const int k0 = chunk * CV_KC;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int k = k0 + lk0 + 2 * i;
      float v = 0.f;
      if (lvalid && k < Ktot) {
        const int cin = k / TAPS, tap = k - cin * TAPS;
        const int dy = tap / KS, dx = tap - dy * KS;
        int s = 0, c = cin;
        if (c >= a.src[0].C) {
          c -= a.src[0].C; s = 1;
          if (c >= a.src[1].C) { c -= a.src[1].C; s = 2; }
        }
        // select by value (no dynamic indexing of the kernel parameter struct -> no local-memory copy)
        const float* sp = s == 0 ? a.src[0].ptr : (s == 1 ? a.src[1].ptr : a.src[2].ptr);
        const int SC = s == 0 ? a.src[0].C : (s == 1 ? a.src[1].C : a.src[2].C);
        const int SH = s == 0 ? a.src[0].H : (s == 1 ? a.src[1].H : a.src[2].H);
        const int SW = s == 0 ? a.src[0].W : (s == 1 ? a.src[1].W : a.src[2].W);
        const int uh = s == 0 ? a.src[0].uh : (s == 1 ? a.src[1].uh : a.src[2].uh);
        const int uw = s == 0 ? a.src[0].uw : (s == 1 ? a.src[1].uw : a.src[2].uw);
        const int nhwc = s == 0 ? a.src[0].nhwc : (s == 1 ? a.src[1].nhwc : a.src[2].nhwc);
        const int iy = iy0 + dy, ix = ix0 + dx;
        if (iy >= 0 && ix >= 0 && iy < SH * uh && ix < SW * uw) {
          const int sy = (uh == 2) ? (iy >> 1) : iy, sx = (uw == 2) ? (ix >> 1) : ix;
          v = nhwc ? __ldg(sp + (((size_t)lb * SH + sy) * SW + sx) * SC + c)
                   : __ldg(sp + (((size_t)lb * SC + c) * SH + sy) * SW + sx);
        }
      }
      ra[i] = v;
    }
    const int k = k0 + wk;
#pragma unroll
    for (int i = 0; i < 8; ++i) rw[i] = (k < Ktot) ? __ldg(a.weight + (size_t)(wc0 + 8 * i) * Ktot + k) : 0.f;
  };
  auto store_chunk = [&](int buf) {
    float* Ab = As + buf * CV_KC * CV_TM;
    float* Wb = Ws + buf * CV_KC * CV_WLD;
#pragma unroll
    for (int i = 0; i < 16; ++i) Ab[(lk0 + 2 * i) * CV_TM + lp] = ra[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) Wb[wk * CV_WLD + wc0 + 8 * i] = rw[i];
  };

  const int tx = tid & 15, ty = tid >> 4;     // 16 pixel-quads x 16 cout-quads
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load_chunk(0);
  store_chunk(0);
  __syncthreads();
  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunks) load_chunk(ch + 1);
    const float* Ab = As + buf * CV_KC * CV_TM;
    const float* Wb = Ws + buf * CV_KC * CV_WLD;
#pragma unroll
    for (int kk = 0; kk < CV_KC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(Ab + kk * CV_TM + tx * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(Ab + kk * CV_TM + 64 + tx * 4);
      const float4 w = *reinterpret_cast<const float4*>(Wb + kk * CV_WLD + ty * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(av[i], wv[j], acc[i][j]);
    }
    if (ch + 1 < nchunks) store_chunk(buf ^ 1);
    __syncthreads();
  }

  // epilogue: bias, ReLU, optional skip
  if (a.out_nhwc) {          // [B,Ho,Wo,64]: one float4 = 4 output channels of a pixel
    const float4 bs = __ldg(reinterpret_cast<const float4*>(a.bias + ty * 4));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long m = m_base + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
      if (m >= Mtot) continue;
      float4 v;
      v.x = fmaxf(acc[i][0] + bs.x, 0.f); v.y = fmaxf(acc[i][1] + bs.y, 0.f);
      v.z = fmaxf(acc[i][2] + bs.z, 0.f); v.w = fmaxf(acc[i][3] + bs.w, 0.f);
      const size_t o = (size_t)m * 64 + ty * 4;
      if (a.skip != nullptr) {
        const float4 sk = __ldg(reinterpret_cast<const float4*>(a.skip + o));
        v.x += sk.x; v.y += sk.y; v.z += sk.z; v.w += sk.w;
      }
      *reinterpret_cast<float4*>(a.out + o) = v;
    }
    return;
  }
  // NCHW float4 stores (4 consecutive pixels per store)
#pragma unroll
  for (int pg = 0; pg < 2; ++pg) {
    const long long m0 = m_base + pg * 64 + tx * 4;
    if (m0 >= Mtot) continue;
    const int b = (int)(m0 / HoWo);
    const int rem = (int)(m0 - (long long)b * HoWo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = ty * 4 + j;
      const float bs = __ldg(a.bias + co);
      float4 v;
      v.x = fmaxf(acc[pg * 4 + 0][j] + bs, 0.f);
      v.y = fmaxf(acc[pg * 4 + 1][j] + bs, 0.f);
      v.z = fmaxf(acc[pg * 4 + 2][j] + bs, 0.f);
      v.w = fmaxf(acc[pg * 4 + 3][j] + bs, 0.f);
      const size_t o = ((size_t)b * 64 + co) * HoWo + rem;
      if (a.skip != nullptr) {
        const float4 sk = __ldg(reinterpret_cast<const float4*>(a.skip + o));
        v.x += sk.x; v.y += sk.y; v.z += sk.z; v.w += sk.w;
      }
      *reinterpret_cast<float4*>(a.out + o) = v;
    }
  }
}

// =====================================================================================
// CBAM on the [64, P] control-point map (P = py*px <= 64), one CTA per image
// =====================================================================================
__global__ void __launch_bounds__(256) cbam_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                    const float* __restrict__ w0, const float* __restrict__ w2,
                                                    const float* __restrict__ spw, const float* __restrict__ spb,
                                                    int py, int px) {
  pdl_trigger();
  pdl_wait();
  __shared__ float xs[64][65];
  __shared__ float avg[64], mxv[64], hid[2][4], gate[64];
  __shared__ float spm[2][66];   // [mean|max][p]
  __shared__ float sg[64];
  const int P = py * px;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* xb = x + (size_t)b * 64 * P;
  for (int i = tid; i < 64 * P; i += 256) xs[i / P][i % P] = __ldg(xb + i);
  __syncthreads();
  if (tid < 64) {
    float s = 0.f, m = -INFINITY;
    for (int p = 0; p < P; ++p) { s += xs[tid][p]; m = fmaxf(m, xs[tid][p]); }
    avg[tid] = s / (float)P;
    mxv[tid] = m;
  }
  __syncthreads();
  if (tid < 8) {
    const int which = tid >> 2, j = tid & 3;
    const float* v = which ? mxv : avg;
    float s = 0.f;
    for (int c = 0; c < 64; ++c) s = __fmaf_rn(__ldg(w0 + j * 64 + c), v[c], s);
    hid[which][j] = fmaxf(s, 0.f);
  }
  __syncthreads();
  if (tid < 64) {
    float sa = 0.f, sm = 0.f;
    for (int j = 0; j < 4; ++j) {
      sa = __fmaf_rn(__ldg(w2 + tid * 4 + j), hid[0][j], sa);
      sm = __fmaf_rn(__ldg(w2 + tid * 4 + j), hid[1][j], sm);
    }
    gate[tid] = 1.f / (1.f + expf(-(sa + sm)));
  }
  __syncthreads();
  for (int i = tid; i < 64 * P; i += 256) xs[i / P][i % P] *= gate[i / P];
  __syncthreads();
  if (tid < P) {
    float s = 0.f, m = -INFINITY;
    for (int c = 0; c < 64; ++c) { s += xs[c][tid]; m = fmaxf(m, xs[c][tid]); }
    spm[0][tid] = s / 64.f;
    spm[1][tid] = m;
  }
  __syncthreads();
  if (tid < P) {
    const int y = tid / px, xx = tid - y * px;
    float s = __ldg(spb);
    for (int ch = 0; ch < 2; ++ch)
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx) {
          const int yy = y + dy - 1, xc = xx + dx - 1;
          if (yy >= 0 && yy < py && xc >= 0 && xc < px)
            s = __fmaf_rn(__ldg(spw + (ch * 3 + dy) * 3 + dx), spm[ch][yy * px + xc], s);
        }
    sg[tid] = 1.f / (1.f + expf(-s));
  }
  __syncthreads();
  float* ob = out + (size_t)b * 64 * P;
  for (int i = tid; i < 64 * P; i += 256) ob[i] = xs[i / P][i % P] * sg[i % P];
}

// =====================================================================================
// localisation head (C') and p_linear(en) -> p1, one CTA per image
// =====================================================================================
struct LocArgs {
  const float* e3;   // [B,64,F]
  const float *wa, *ba, *wb, *bb, *wc, *bc;      // loc1.0 [256,64], loc1.2 [2,256], loc2 [2F,2F]
  const float *wp0, *bp0, *wp1, *bp1;            // p_linear.0 [32,64], p_linear.1 [128,32]
  float* c_prime;    // [B,F,2]
  float* p1;         // [B,F,128]
  float* p1img;      // [B][4 chunks][hi|lo][8 k-groups][32 rows][4] tensor-core operand image of p1, or null
  int F;
};

__global__ void __launch_bounds__(256) loc_p1_kernel(LocArgs a) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float sm[];
  const int F = a.F, b = blockIdx.x, tid = threadIdx.x;
  float* en = sm;                 // [F][65]
  float* z1 = en + F * 65;        // [F][256]
  float* z2 = z1 + F * 256;       // [2F]
  float* t1 = z2 + 2 * F;         // [F][33]
  for (int i = tid; i < 64 * F; i += 256) {
    const int c = i / F, k = i - c * F;
    en[k * 65 + c] = __ldg(a.e3 + (size_t)b * 64 * F + i);
  }
  __syncthreads();
  // The CTA is one image and the whole launch is a single wave, so its time is the length of the dependent FMA chains
  // below: every loop carries four independent outputs (each still summed in index order, so results are unchanged).
  {  // z1[k][m] = relu(Wa[m,:] . en[k,:] + ba[m]), m = tid
    float w[64];                     // this thread's row of loc1.0: 16 x 16-byte loads (rows are 256 B apart across
#pragma unroll                      // lanes, so scalar loads were 64 fully uncoalesced requests per warp: LG-throttle bound)
    for (int c = 0; c < 64; c += 4) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(a.wa + tid * 64 + c));
      w[c] = t4.x; w[c + 1] = t4.y; w[c + 2] = t4.z; w[c + 3] = t4.w;
    }
    const float bs = __ldg(a.ba + tid);
    int k = 0;
    for (; k + 3 < F; k += 4) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        s0 = __fmaf_rn(w[c], en[k * 65 + c], s0); s1 = __fmaf_rn(w[c], en[(k + 1) * 65 + c], s1);
        s2 = __fmaf_rn(w[c], en[(k + 2) * 65 + c], s2); s3 = __fmaf_rn(w[c], en[(k + 3) * 65 + c], s3);
      }
      z1[k * 256 + tid] = fmaxf(s0 + bs, 0.f); z1[(k + 1) * 256 + tid] = fmaxf(s1 + bs, 0.f);
      z1[(k + 2) * 256 + tid] = fmaxf(s2 + bs, 0.f); z1[(k + 3) * 256 + tid] = fmaxf(s3 + bs, 0.f);
    }
    for (; k < F; ++k) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) s = __fmaf_rn(w[c], en[k * 65 + c], s);
      z1[k * 256 + tid] = fmaxf(s + bs, 0.f);
    }
  }
  // t1[k][q] = Wp0[q,:] . en[k,:] + bp0[q]: thread -> q = tid & 31, control points (tid >> 5) + 8 i, four at a time
  {
    const int q = tid & 31;
    const float bq = __ldg(a.bp0 + q);
    for (int k0 = tid >> 5; k0 < F; k0 += 32) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      const int k1 = min(k0 + 8, F - 1), k2 = min(k0 + 16, F - 1), k3 = min(k0 + 24, F - 1);
#pragma unroll 4
      for (int c4 = 0; c4 < 64; c4 += 4) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.wp0 + q * 64 + c4));
        const float wq[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c4 + u;
          s0 = __fmaf_rn(wq[u], en[k0 * 65 + c], s0); s1 = __fmaf_rn(wq[u], en[k1 * 65 + c], s1);
          s2 = __fmaf_rn(wq[u], en[k2 * 65 + c], s2); s3 = __fmaf_rn(wq[u], en[k3 * 65 + c], s3);
        }
      }
      t1[k0 * 33 + q] = s0 + bq;
      if (k0 + 8 < F) t1[(k0 + 8) * 33 + q] = s1 + bq;
      if (k0 + 16 < F) t1[(k0 + 16) * 33 + q] = s2 + bq;
      if (k0 + 24 < F) t1[(k0 + 24) * 33 + q] = s3 + bq;
    }
  }
  __syncthreads();
  // z2[k*2+j] = relu(Wb[j,:] . z1[k,:] + bb[j]) : 8 lanes per output
  for (int o0 = (tid >> 3); o0 < 2 * F; o0 += 32) {
    const int k = o0 >> 1, j = o0 & 1, l = tid & 7;
    float s = 0.f;
#pragma unroll 8
    for (int m = l; m < 256; m += 8) s = __fmaf_rn(__ldg(a.wb + j * 256 + m), z1[k * 256 + m], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (l == 0) z2[o0] = fmaxf(s + __ldg(a.bb + j), 0.f);
  }
  // p1[k][r] = Wp1[r,:] . t1[k,:] + bp1[r]: thread -> r = tid & 127, control points (tid >> 7) + 2 i, four at a time
  {
    const int r = tid & 127;
    float wr[32];
#pragma unroll
    for (int q = 0; q < 32; q += 4) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(a.wp1 + r * 32 + q));
      wr[q] = t4.x; wr[q + 1] = t4.y; wr[q + 2] = t4.z; wr[q + 3] = t4.w;
    }
    const float br = __ldg(a.bp1 + r);
    for (int k0 = tid >> 7; k0 < F; k0 += 8) {
      float sv[4] = {0.f, 0.f, 0.f, 0.f};
      int kk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) kk[u] = min(k0 + 2 * u, F - 1);
#pragma unroll
      for (int q = 0; q < 32; ++q) {
#pragma unroll
        for (int u = 0; u < 4; ++u) sv[u] = __fmaf_rn(wr[q], t1[kk[u] * 33 + q], sv[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + 2 * u;
        if (k >= F) break;
        const float pv = sv[u] + br;
        a.p1[((size_t)b * F + k) * 128 + r] = pv;
        if (a.p1img != nullptr) {      // B operand of the score GEMM: row n = control point k, K index = r
          const float hi = __uint_as_float(__float_as_uint(pv) & 0xFFFFE000u);
          float* o = a.p1img + (size_t)b * 8192 + (size_t)(r >> 5) * 2048 + ((r >> 2) & 7) * 128 + k * 4 + (r & 3);
          o[0] = hi;
          o[1024] = pv - hi;
        }
      }
    }
  }
  __syncthreads();
  if (tid < 2 * F) {
    float s = 0.f;
    int i = 0;
    if ((F & 1) == 0) {                // rows of loc2 are 8F bytes: 16-byte loads when that is a multiple of 16
#pragma unroll 4
      for (; i + 3 < 2 * F; i += 4) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.wc + tid * 2 * F + i));
        s = __fmaf_rn(w4.x, z2[i], s); s = __fmaf_rn(w4.y, z2[i + 1], s);
        s = __fmaf_rn(w4.z, z2[i + 2], s); s = __fmaf_rn(w4.w, z2[i + 3], s);
      }
    }
    for (; i < 2 * F; ++i) s = __fmaf_rn(__ldg(a.wc + tid * 2 * F + i), z2[i], s);
    a.c_prime[(size_t)b * 2 * F + tid] = s + __ldg(a.bc + tid);
  }
}

// =====================================================================================
// DGAB part 1, one CTA per (image, channel) plane [H, 64]:
//   u = LN1(x); gates from axial means + control-point features; a = u*(v_h*h_last + v_w*w_last);
//   x1 = x + proj_W(a); v = LN2(x1)
// =====================================================================================
struct DgabArgs {
  const float* x;     // de_feat [B,64,H,64]
  const float* e3;    // [B,64,F]  (y^T[b,c,:] is exactly this plane)
  const float *n1w, *n1b, *n2w, *n2b;   // [H,64]
  const float *wh, *ww;                 // mlp_h [H+1, H+F], mlp_w [65, 64+F]
  const float *wp, *bp;                 // proj [64,64]
  float* x1;          // [B,64,H,64]
  float* v;           // [B,64,H,64]
  int H, F;
};

constexpr int DG_MAXH = 32;

// DGAB gating block, one warp per plane: a [H,64] plane is 2H elements per lane, so LayerNorm statistics, the two
// gate soft-maxes and the row/column means are warp shuffles and the whole plane needs no block barrier (the round-1
// block-per-plane kernel spent most of its 5.7 us per plane waiting in ~13 __syncthreads; removed).  16 warps per CTA,
// one CTA per SM, weights shared in shared memory, per-warp scratch for the gated plane.
constexpr int DW_WARPS = 16;
template <int H>
__global__ void __launch_bounds__(DW_WARPS * 32, 1) dgab_warp_kernel(DgabArgs a, int nplanes) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float dsm[];
  constexpr int n = H * 64, E = 2 * H;             // E elements per lane: idx = lane + 32 i  ->  h = i >> 1, w = lane + 32 (i & 1)
  const int F = a.F, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int LW = 64 + F + 1, LH = H + F + 1;
  float* n1w = dsm;                      // [n] x4
  float* n1b = n1w + n;
  float* n2w = n1b + n;
  float* n2b = n2w + n;
  float* wpT = n2b + n;                  // [64][65]  wpT[w][j] = proj.weight[j][w]
  float* wws = wpT + 64 * 65;            // [65][LW]
  float* whs = wws + 65 * LW;            // [H+1][LH]
  float* scratch = dsm + ((4 * n + 64 * 65 + 65 * LW + (H + 1) * LH + 3) & ~3);   // 16-byte aligned (float4 reads of `as`)
  const int per_warp = (n + (64 + F) + (H + F) + 3) & ~3;
  float* as = scratch + (size_t)warp * per_warp;   // [H][64] gated plane (the normalised plane stays in registers)
  float* vecw = as + n;                            // [64+F] = colmean | y
  float* vech = vecw + 64 + F;                     // [H+F]  = rowmean | y

  for (int i = tid; i < 4096; i += DW_WARPS * 32) wpT[(i & 63) * 65 + (i >> 6)] = __ldg(a.wp + i);
  for (int i = tid; i < 65 * (64 + F); i += DW_WARPS * 32) { const int r = i / (64 + F); wws[r * LW + (i - r * (64 + F))] = __ldg(a.ww + i); }
  for (int i = tid; i < (H + 1) * (H + F); i += DW_WARPS * 32) { const int r = i / (H + F); whs[r * LH + (i - r * (H + F))] = __ldg(a.wh + i); }
  for (int i = tid; i < n; i += DW_WARPS * 32) {
    n1w[i] = __ldg(a.n1w + i); n1b[i] = __ldg(a.n1b + i); n2w[i] = __ldg(a.n2w + i); n2b[i] = __ldg(a.n2b + i);
  }
  const float bp0 = __ldg(a.bp + lane), bp1 = __ldg(a.bp + lane + 32);
  __syncthreads();

  const int stride = gridDim.x * DW_WARPS;
  for (int pl = blockIdx.x * DW_WARPS + warp; pl < nplanes; pl += stride) {
    const size_t plane = (size_t)pl * n;
    float xv[E];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) { xv[i] = __ldg(a.x + plane + lane + 32 * i); s += xv[i]; }
    for (int f = lane; f < F; f += 32) { const float y = __ldg(a.e3 + (size_t)pl * F + f); vecw[64 + f] = y; vech[H + f] = y; }
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    const float mean = s / (float)n;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) { const float d = xv[i] - mean; q = __fmaf_rn(d, d, q); }
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) q += __shfl_xor_sync(0xffffffffu, q, k);
    const float rstd = rsqrtf(q / (float)n + 1e-5f);
    float u[E];
    float c0 = 0.f, c1 = 0.f;                      // column sums of w = lane and w = lane + 32
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int idx = lane + 32 * i;
      u[i] = (xv[i] - mean) * rstd * n1w[idx] + n1b[idx];
      if (i & 1) c1 += u[i]; else c0 += u[i];
    }
    vecw[lane] = c0 / (float)H;
    vecw[lane + 32] = c1 / (float)H;
#pragma unroll
    for (int h = 0; h < H; ++h) {                  // row means
      float t = u[2 * h] + u[2 * h + 1];
#pragma unroll
      for (int k = 16; k >= 1; k >>= 1) t += __shfl_xor_sync(0xffffffffu, t, k);
      if (lane == (h & 31)) vech[h] = t / 64.f;
    }
    __syncwarp();
    // gate logits: width logits lane and lane + 32, the 65th by a warp reduction; height logits on lanes 0..H
    float l0 = 0.f, l1 = 0.f, l64 = 0.f, lhv = 0.f;
    {
      const float* r0 = wws + lane * LW;
      const float* r1 = wws + (lane + 32) * LW;
      float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
      for (int k = 0; k + 1 < 64 + F; k += 2) {
        const float v0 = vecw[k], v1 = vecw[k + 1];
        t0 = __fmaf_rn(r0[k], v0, t0); t1 = __fmaf_rn(r0[k + 1], v1, t1);
        t2 = __fmaf_rn(r1[k], v0, t2); t3 = __fmaf_rn(r1[k + 1], v1, t3);
      }
      if ((64 + F) & 1) { const float v0 = vecw[64 + F - 1]; t0 = __fmaf_rn(r0[64 + F - 1], v0, t0); t2 = __fmaf_rn(r1[64 + F - 1], v0, t2); }
      l0 = t0 + t1; l1 = t2 + t3;
      const float* r64 = wws + 64 * LW;
      float t = 0.f;
      for (int k = lane; k < 64 + F; k += 32) t = __fmaf_rn(r64[k], vecw[k], t);
#pragma unroll
      for (int k = 16; k >= 1; k >>= 1) t += __shfl_xor_sync(0xffffffffu, t, k);
      l64 = t;
      if (lane <= H) {
        const float* rh = whs + lane * LH;
        float h0 = 0.f, h1 = 0.f;
        for (int k = 0; k + 1 < H + F; k += 2) { h0 = __fmaf_rn(rh[k], vech[k], h0); h1 = __fmaf_rn(rh[k + 1], vech[k + 1], h1); }
        if ((H + F) & 1) h0 = __fmaf_rn(rh[H + F - 1], vech[H + F - 1], h0);
        lhv = h0 + h1;
      }
    }
    // soft-max over the 64 width logits (two per lane) and over the H height logits (lanes 0..H-1)
    float vw0, vw1, vhv;
    {
      float m = fmaxf(l0, l1);
#pragma unroll
      for (int k = 16; k >= 1; k >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, k));
      const float e0 = expf(l0 - m), e1 = expf(l1 - m);
      float t = e0 + e1;
#pragma unroll
      for (int k = 16; k >= 1; k >>= 1) t += __shfl_xor_sync(0xffffffffu, t, k);
      vw0 = e0 / t; vw1 = e1 / t;
      float mh = lane < H ? lhv : -INFINITY;
#pragma unroll
      for (int k = 16; k >= 1; k >>= 1) mh = fmaxf(mh, __shfl_xor_sync(0xffffffffu, mh, k));
      const float eh = lane < H ? expf(lhv - mh) : 0.f;
      float th = eh;
#pragma unroll
      for (int k = 16; k >= 1; k >>= 1) th += __shfl_xor_sync(0xffffffffu, th, k);
      vhv = eh / th;
    }
    const float hl = __shfl_sync(0xffffffffu, lhv, H), wl = l64;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const float vh = __shfl_sync(0xffffffffu, vhv, i >> 1);
      const float vw = (i & 1) ? vw1 : vw0;
      as[lane + 32 * i] = (vh * u[i]) * hl + (vw * u[i]) * wl;     // same association as DGAB.py:50
    }
    __syncwarp();
    // proj over the width axis + residual: lane -> output columns j = lane, lane + 32, all H rows
    float o[E];
#pragma unroll
    for (int i = 0; i < E; ++i) o[i] = 0.f;
#pragma unroll 1
    for (int w = 0; w < 64; w += 4) {
      float wa[4], wb[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) { wa[t] = wpT[(w + t) * 65 + lane]; wb[t] = wpT[(w + t) * 65 + lane + 32]; }
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float4 av = *reinterpret_cast<const float4*>(as + h * 64 + w);   // warp-wide broadcast
        o[2 * h] = __fmaf_rn(av.x, wa[0], o[2 * h]); o[2 * h] = __fmaf_rn(av.y, wa[1], o[2 * h]);
        o[2 * h] = __fmaf_rn(av.z, wa[2], o[2 * h]); o[2 * h] = __fmaf_rn(av.w, wa[3], o[2 * h]);
        o[2 * h + 1] = __fmaf_rn(av.x, wb[0], o[2 * h + 1]); o[2 * h + 1] = __fmaf_rn(av.y, wb[1], o[2 * h + 1]);
        o[2 * h + 1] = __fmaf_rn(av.z, wb[2], o[2 * h + 1]); o[2 * h + 1] = __fmaf_rn(av.w, wb[3], o[2 * h + 1]);
      }
    }
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      o[i] = xv[i] + (o[i] + ((i & 1) ? bp1 : bp0));
      a.x1[plane + lane + 32 * i] = o[i];
      s2 += o[i];
    }
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, k);
    const float mean2 = s2 / (float)n;
    float q2 = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) { const float d = o[i] - mean2; q2 = __fmaf_rn(d, d, q2); }
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, k);
    const float rstd2 = rsqrtf(q2 / (float)n + 1e-5f);
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int idx = lane + 32 * i;
      a.v[plane + idx] = (o[i] - mean2) * rstd2 * n2w[idx] + n2b[idx];
    }
    __syncwarp();          // this warp's scratch is rewritten by its next plane
  }
}

static size_t dgab_warp_smem(int H, int F) {
  const int n = H * 64;
  return sizeof(float) * (size_t)(4 * n + 64 * 65 + 65 * (64 + F + 1) + (H + 1) * (H + F + 1) +
                                  DW_WARPS * (n + (64 + F) + (H + F) + 4) + 8);
}


// =====================================================================================
// DGAB part 2: x2 = x1 + fc2(GELU(fc1(v))) over the width axis; rows = (b, c, h), K = 64
// =====================================================================================
struct MlpArgs {
  const float* v;     // [R,64]
  const float* x1;    // [R,64]
  const float *w1, *b1, *w2, *b2;   // fc1 [256,64], fc2 [64,256]
  float* out;         // [R,64]
  long long R;
};
constexpr int ML_SMEM = (64 * 128 + 64 * 128 + 64 * CV_WLD + 64 * CV_WLD) * 4;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

__global__ void __launch_bounds__(256, 2) dgab_mlp_kernel(MlpArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                    // [64 k][128 rows]
  float* Hs = As + 64 * 128;           // [64 m][128 rows]
  float* W1s = Hs + 64 * 128;          // [64 k][68]  (m within chunk)
  float* W2s = W1s + 64 * CV_WLD;      // [64 m][68]  (j)
  const int tid = threadIdx.x;
  const long long r_base = (long long)blockIdx.x * 128;
  {  // input tile, transposed: thread -> row lp, float4 along k
    const int lp = tid & 127, k40 = tid >> 7;
    const long long r = r_base + lp;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k4 = k40 + 2 * i;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < a.R) t = __ldg(reinterpret_cast<const float4*>(a.v + r * 64 + k4 * 4));
      As[(k4 * 4 + 0) * 128 + lp] = t.x;
      As[(k4 * 4 + 1) * 128 + lp] = t.y;
      As[(k4 * 4 + 2) * 128 + lp] = t.z;
      As[(k4 * 4 + 3) * 128 + lp] = t.w;
    }
  }
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int hc = 0; hc < 4; ++hc) {
    {  // weight chunks: W1s[k][mm] = fc1.w[hc*64+mm][k];  W2s[mm][j] = fc2.w[j][hc*64+mm]
      const int c = tid & 63, r0 = tid >> 6;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int rr = r0 + 4 * i;
        W1s[c * CV_WLD + rr] = __ldg(a.w1 + (size_t)(hc * 64 + rr) * 64 + c);
        W2s[c * CV_WLD + rr] = __ldg(a.w2 + (size_t)rr * 256 + hc * 64 + c);
      }
    }
    __syncthreads();
    float hacc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) hacc[i][j] = 0.f;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(As + k * 128 + tx * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(As + k * 128 + 64 + tx * 4);
      const float4 w = *reinterpret_cast<const float4*>(W1s + k * CV_WLD + ty * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) hacc[i][j] = __fmaf_rn(av[i], wv[j], hacc[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float bs = __ldg(a.b1 + hc * 64 + ty * 4 + j);
      float4 h0, h1;
      h0.x = gelu_erf(hacc[0][j] + bs); h0.y = gelu_erf(hacc[1][j] + bs);
      h0.z = gelu_erf(hacc[2][j] + bs); h0.w = gelu_erf(hacc[3][j] + bs);
      h1.x = gelu_erf(hacc[4][j] + bs); h1.y = gelu_erf(hacc[5][j] + bs);
      h1.z = gelu_erf(hacc[6][j] + bs); h1.w = gelu_erf(hacc[7][j] + bs);
      *reinterpret_cast<float4*>(Hs + (ty * 4 + j) * 128 + tx * 4) = h0;
      *reinterpret_cast<float4*>(Hs + (ty * 4 + j) * 128 + 64 + tx * 4) = h1;
    }
    __syncthreads();
#pragma unroll 8
    for (int m = 0; m < 64; ++m) {
      const float4 a0 = *reinterpret_cast<const float4*>(Hs + m * 128 + tx * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(Hs + m * 128 + 64 + tx * 4);
      const float4 w = *reinterpret_cast<const float4*>(W2s + m * CV_WLD + ty * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float4 b2 = __ldg(reinterpret_cast<const float4*>(a.b2 + ty * 4));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = r_base + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
    if (r >= a.R) continue;
    const float4 res = __ldg(reinterpret_cast<const float4*>(a.x1 + r * 64 + ty * 4));
    float4 o;
    o.x = res.x + (acc[i][0] + b2.x); o.y = res.y + (acc[i][1] + b2.y);
    o.z = res.z + (acc[i][2] + b2.z); o.w = res.w + (acc[i][3] + b2.w);
    *reinterpret_cast<float4*>(a.out + r * 64 + ty * 4) = o;
  }
}

// =====================================================================================
// score head: pc_score[b,p,k] = tanh(scale * f[p,:] . p1[b,k,:]),
//   f = feat_linear.1(feat_linear.0(de2[b,:,p]))           (tps_pp.py:303-308, 293-299)
// one CTA per (image, 128 consecutive pixels); F <= 32
// =====================================================================================
struct ScoreArgs {
  const float* de2;   // [B,64,n]
  const float* p1;    // [B,F,128]
  const float *wf0, *bf0, *wf1, *bf1;   // [32,64], [128,32]
  float* score;       // [B,n,F]
  int n, F;
  float scale;
};
constexpr int SC_SMEM = (64 * 128 + 32 * 128 + 128 * 128 + 64 * 36 + 32 * 132 + 128 * 36) * 4;

__global__ void __launch_bounds__(256, 1) score_kernel(ScoreArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                    // [64 c][128 px]
  float* T1 = As + 64 * 128;           // [32 q][128 px]
  float* Fs = T1 + 32 * 128;           // [128 r][128 px]
  float* W0 = Fs + 128 * 128;          // [64 c][36]  (q)
  float* W1 = W0 + 64 * 36;            // [32 q][132] (r)
  float* P1 = W1 + 32 * 132;           // [128 r][36] (k)
  const int tid = threadIdx.x, b = blockIdx.y, p0 = blockIdx.x * 128;
  const int n = a.n, F = a.F;
  for (int i = tid; i < 64 * 128; i += 256) {
    const int c = i >> 7, px = i & 127;
    As[i] = (p0 + px < n) ? __ldg(a.de2 + ((size_t)b * 64 + c) * n + p0 + px) : 0.f;
  }
  for (int i = tid; i < 32 * 64; i += 256) W0[(i & 63) * 36 + (i >> 6)] = __ldg(a.wf0 + i);       // wf0[q][c]
  for (int i = tid; i < 128 * 32; i += 256) W1[(i & 31) * 132 + (i >> 5)] = __ldg(a.wf1 + i);     // wf1[r][q]
  for (int i = tid; i < 128 * 32; i += 256) {
    const int k = i >> 7, r = i & 127;
    P1[r * 36 + k] = (k < F) ? __ldg(a.p1 + ((size_t)b * F + k) * 128 + r) : 0.f;
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  {  // t1[px][q]: 8 px x 2 q per thread
    float acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
#pragma unroll 8
    for (int c = 0; c < 64; ++c) {
      const float4 a0 = *reinterpret_cast<const float4*>(As + c * 128 + tx * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(As + c * 128 + 64 + tx * 4);
      const float2 w = *reinterpret_cast<const float2*>(W0 + c * 36 + ty * 2);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = __fmaf_rn(av[i], w.x, acc[i][0]);
        acc[i][1] = __fmaf_rn(av[i], w.y, acc[i][1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float bs = __ldg(a.bf0 + ty * 2 + j);
      *reinterpret_cast<float4*>(T1 + (ty * 2 + j) * 128 + tx * 4) =
          make_float4(acc[0][j] + bs, acc[1][j] + bs, acc[2][j] + bs, acc[3][j] + bs);
      *reinterpret_cast<float4*>(T1 + (ty * 2 + j) * 128 + 64 + tx * 4) =
          make_float4(acc[4][j] + bs, acc[5][j] + bs, acc[6][j] + bs, acc[7][j] + bs);
    }
  }
  __syncthreads();
  {  // f[px][r]: 8 px x 8 r per thread
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int q = 0; q < 32; ++q) {
      const float4 a0 = *reinterpret_cast<const float4*>(T1 + q * 128 + tx * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(T1 + q * 128 + 64 + tx * 4);
      const float4 w0 = *reinterpret_cast<const float4*>(W1 + q * 132 + ty * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(W1 + q * 132 + ty * 8 + 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(av[i], wv[j], acc[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float bs = __ldg(a.bf1 + ty * 8 + j);
      *reinterpret_cast<float4*>(Fs + (ty * 8 + j) * 128 + tx * 4) =
          make_float4(acc[0][j] + bs, acc[1][j] + bs, acc[2][j] + bs, acc[3][j] + bs);
      *reinterpret_cast<float4*>(Fs + (ty * 8 + j) * 128 + 64 + tx * 4) =
          make_float4(acc[4][j] + bs, acc[5][j] + bs, acc[6][j] + bs, acc[7][j] + bs);
    }
  }
  __syncthreads();
  {  // s[px][k]: 8 px x 2 k per thread, K = 128
    float acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
#pragma unroll 8
    for (int r = 0; r < 128; ++r) {
      const float4 a0 = *reinterpret_cast<const float4*>(Fs + r * 128 + tx * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(Fs + r * 128 + 64 + tx * 4);
      const float2 w = *reinterpret_cast<const float2*>(P1 + r * 36 + ty * 2);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = __fmaf_rn(av[i], w.x, acc[i][0]);
        acc[i][1] = __fmaf_rn(av[i], w.y, acc[i][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int px = p0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
      if (px >= n) continue;
      float* o = a.score + ((size_t)b * n + px) * F;
      const int k = ty * 2;
      if (k < F) o[k] = tanhf(acc[i][0] * a.scale);
      if (k + 1 < F) o[k + 1] = tanhf(acc[i][1] * a.scale);
    }
  }
}

// =====================================================================================
// host orchestration
// =====================================================================================
struct HeadDims {
  int B, h, w, H2, W2, F, py, px, ps;
  int h1, w1, h2, w2;   // after enc1 / enc2
};

static int head_dims(const tpspp_head_cfg* c, HeadDims* d) {
  TPSPP_REQUIRE(c != nullptr, "head cfg is NULL");
  TPSPP_REQUIRE(c->batch >= 0, "batch must be >= 0");
  TPSPP_REQUIRE(c->width == 64, "TPS_PP head needs feature-map width 64 (DGAB linear layers act on the width axis; "
                                "reference DGAB.py:36,52 -- SURVEY F4), got %d", c->width);
  TPSPP_REQUIRE(c->height >= 8 && c->height <= DG_MAXH && c->height % 8 == 0, "height must be a multiple of 8 in [8,%d]", DG_MAXH);
  TPSPP_REQUIRE(c->p_stride == 1 || c->p_stride == 2, "p_stride must be 1 or 2");
  d->B = c->batch; d->h = c->height; d->w = c->width; d->H2 = 2 * c->height; d->W2 = 2 * c->width;
  d->ps = c->p_stride;
  d->h1 = d->h / 2; d->w1 = d->w / 2;
  d->h2 = d->h1 / d->ps; d->w2 = d->w1 / d->ps;
  d->py = d->h2 / 2; d->px = d->w2;
  d->F = d->py * d->px;
  TPSPP_REQUIRE(d->py == c->point_h && d->px == c->point_w,
                "point_size (%d,%d) must equal the MSFA encoder's output lattice (%d,%d) (SURVEY F4)", c->point_h,
                c->point_w, d->py, d->px);
  TPSPP_REQUIRE(d->F <= 32 && d->F % 2 == 0, "num_fiducial must be even and <= 32 for the score kernel (got %d)", d->F);
  TPSPP_REQUIRE(c->precision == TPSPP_HEAD_FP32 || c->precision == TPSPP_HEAD_TC || c->precision == TPSPP_HEAD_BF16,
                "unknown precision %d", c->precision);
  return TPSPP_OK;
}

// the 14 convolutions in launch order: weight index, Cin total, kernel size
// plus the four linear layers the tensor-core engine runs as 1x1 convolutions over rows
struct ConvLayerDesc { int w_idx, Ctot, KS, N, NT; };
constexpr int kNumTcLayers = 18;
enum { TCL_FLIN0 = 14, TCL_FLIN1 = 15, TCL_FC1 = 16, TCL_FC2 = 17 };
static const ConvLayerDesc kConvLayers[kNumTcLayers] = {
    {TPSPP_P_DOWN0_W, 32, 1, 64, 64},   {TPSPP_P_DOWN1_W, 32, 1, 64, 64},    {TPSPP_P_DOWN2_W, 64, 1, 64, 64},
    {TPSPP_P_DOWN0_1_W, 64, 3, 64, 64}, {TPSPP_P_DOWN1_1_W, 64, 3, 64, 64},  {TPSPP_P_DOWNFEAT_W, 192, 1, 64, 64},
    {TPSPP_P_ENC0_W, 192, 3, 64, 64},   {TPSPP_P_ENC1_W, 64, 3, 64, 64},     {TPSPP_P_ENC2_W, 64, 3, 64, 64},
    {TPSPP_P_ENC3_W, 64, 3, 64, 64},    {TPSPP_P_DEC0_W, 64, 3, 64, 64},     {TPSPP_P_DEC1_W, 64, 3, 64, 64},
    {TPSPP_P_DEC2_W, 64, 3, 64, 64},    {TPSPP_P_DEC3_W, 64, 3, 64, 64},
    {TPSPP_P_FLIN0_W, 64, 1, 32, 32},   {TPSPP_P_FLIN1_W, 32, 1, 128, 64},   {TPSPP_P_FC1_W, 64, 1, 256, 64},
    {TPSPP_P_FC2_W, 256, 1, 64, 64}};
static size_t wprep_total_floats() {
  size_t t = 0;
  for (int i = 0; i < kNumTcLayers; ++i) t += conv_tc_wprep_floats(kConvLayers[i].Ctot, kConvLayers[i].KS, kConvLayers[i].N);
  return t;
}

static void head_offsets(const HeadDims& d, size_t* off, size_t* total) {
  const size_t B = d.B;
  size_t sz[TPSPP_WS_COUNT];
  const size_t big = B * 64 * d.H2 * d.W2, mid = B * 64 * d.h * d.w;
  sz[TPSPP_WS_F0] = big; sz[TPSPP_WS_F1] = big; sz[TPSPP_WS_F2] = mid; sz[TPSPP_WS_A0] = mid; sz[TPSPP_WS_A1] = mid;
  sz[TPSPP_WS_E0] = mid; sz[TPSPP_WS_E1] = B * 64 * d.h1 * d.w1; sz[TPSPP_WS_E2] = B * 64 * d.h2 * d.w2;
  sz[TPSPP_WS_E3] = B * 64 * d.F; sz[TPSPP_WS_CBAM] = B * 64 * d.F; sz[TPSPP_WS_D0] = sz[TPSPP_WS_E2];
  sz[TPSPP_WS_D1] = sz[TPSPP_WS_E1]; sz[TPSPP_WS_D2] = mid; sz[TPSPP_WS_DE] = mid; sz[TPSPP_WS_X1] = mid;
  sz[TPSPP_WS_V] = mid; sz[TPSPP_WS_DE2] = mid; sz[TPSPP_WS_P1] = B * d.F * 128;
  sz[TPSPP_WS_WPREP] = wprep_total_floats();
  sz[TPSPP_WS_T1] = B * d.h * d.w * 32;          // feat_linear.0 output, rows x 32
  sz[TPSPP_WS_FS] = B * d.h * d.w * 128;         // feat_linear.1 output, rows x 128
  sz[TPSPP_WS_HID] = mid * 4;                    // Mlp hidden, rows x 256
  sz[TPSPP_WS_P1IMG] = B * 8192;                 // per-image UMMA operand image of p1 (hi | lo)
  size_t cur = 0;
  for (int i = 0; i < TPSPP_WS_COUNT; ++i) {
    off[i] = cur;
    cur += (sz[i] * sizeof(float) + 255) / 256 * 256;
  }
  *total = cur + 256;
}

static ConvSrc mk_src(const float* p, int C, int H, int W, int nhwc, int uh = 1, int uw = 1, int bf16 = 0) {
  ConvSrc s; s.ptr = p; s.C = C; s.H = H; s.W = W; s.uh = uh; s.uw = uw; s.nhwc = nhwc; s.bf16 = bf16; return s;
}

static int run_conv(int KS, ConvSrc s0, ConvSrc s1, ConvSrc s2, const float* w, const float* bias, const float* skip,
                    float* out, int out_nhwc, int B, int Ho, int Wo, int sh, int sw, cudaStream_t st,
                    const float* wprep = nullptr, int mode = CM_TF32X3, int out_bf16 = 0, int skip_bf16 = 0) {
  ConvArgs a;
  a.out_nhwc = out_nhwc; a.out_bf16 = out_bf16; a.skip_bf16 = skip_bf16;
  a.act = CONV_ACT_RELU; a.act_scale = 1.f; a.Cout = 64; a.wimg_stride = 0; a.skip_pre = 0; a.out_cstride = 0; a.zi = 0;
  a.src[0] = s0; a.src[1] = s1; a.src[2] = s2;
  a.weight = w; a.bias = bias; a.skip = skip; a.out = out;
  a.B = B; a.Ho = Ho; a.Wo = Wo; a.Ctot = s0.C + s1.C + s2.C; a.sh = sh; a.sw = sw; a.pad = (KS == 3) ? 1 : 0;
  if (wprep != nullptr && conv_tc_eligible(a, KS)) return run_conv_tc(KS, a, wprep, 64, st, mode);
  const long long M = (long long)B * Ho * Wo;
  const unsigned grid = (unsigned)((M + CV_TM - 1) / CV_TM);
  if (KS == 1) conv_ffma_kernel<1><<<grid, 256, CV_SMEM, st>>>(a);
  else conv_ffma_kernel<3><<<grid, 256, CV_SMEM, st>>>(a);
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());
  return TPSPP_OK;
}

static int head_attrs_once() {
  static thread_local int done_dev = -1;
  int dev = 0;
  TPSPP_CHECK_CUDA(cudaGetDevice(&dev));
  if (done_dev == dev) return TPSPP_OK;
  TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_ffma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SMEM));
  TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_ffma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SMEM));
  TPSPP_CHECK_CUDA(cudaFuncSetAttribute(dgab_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ML_SMEM));
  TPSPP_CHECK_CUDA(cudaFuncSetAttribute(score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM));
  TPSPP_CHECK_CUDA(cudaFuncSetAttribute(loc_p1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  done_dev = dev;
  return TPSPP_OK;
}

}  // namespace tpspp

using namespace tpspp;

extern "C" size_t tpspp_head_workspace_bytes(const tpspp_head_cfg* cfg) {
  HeadDims d;
  if (head_dims(cfg, &d) != TPSPP_OK) return 0;
  size_t off[TPSPP_WS_COUNT], total;
  head_offsets(d, off, &total);
  return total;
}

extern "C" int tpspp_head_workspace_offsets(const tpspp_head_cfg* cfg, size_t* offsets) {
  HeadDims d;
  int rc = head_dims(cfg, &d);
  if (rc != TPSPP_OK) return rc;
  TPSPP_REQUIRE(offsets != nullptr, "offsets is NULL");
  size_t total;
  head_offsets(d, offsets, &total);
  return TPSPP_OK;
}

extern "C" int tpspp_head_fwd(const tpspp_head_cfg* cfg, const float* x, const float* o0, const float* o1,
                              const float* const* P, float* feat_grid, float* c_prime, float* pc_score,
                              void* workspace, tpspp_stream_t stream) {
  reset_launch_count();
  prof_begin((cudaStream_t)stream);
  HeadDims d;
  int rc = head_dims(cfg, &d);
  if (rc != TPSPP_OK) return rc;
  if (d.B == 0) return TPSPP_OK;
  TPSPP_REQUIRE(x && o0 && o1 && P && feat_grid && c_prime && pc_score && workspace, "tpspp_head_fwd: null pointer");
  TPSPP_REQUIRE(((uintptr_t)workspace & 255) == 0, "tpspp_head_fwd: workspace must be 256-byte aligned");
  for (int i = 0; i < TPSPP_P_COUNT; ++i) {
    TPSPP_REQUIRE(P[i] != nullptr, "tpspp_head_fwd: params[%d] is NULL", i);
    TPSPP_REQUIRE(((uintptr_t)P[i] & 15) == 0, "tpspp_head_fwd: params[%d] must be 16-byte aligned (vector loads)", i);
  }
  if (d.B == 0) return TPSPP_OK;
  rc = head_attrs_once();
  if (rc != TPSPP_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  struct PdlScope { PdlScope(bool on) { pdl_scope(on); } ~PdlScope() { pdl_scope(false); } } pdl_guard(cfg->precision != TPSPP_HEAD_FP32);
  size_t off[TPSPP_WS_COUNT], total;
  head_offsets(d, off, &total);
  auto W = [&](int i) { return reinterpret_cast<float*>((char*)workspace + off[i]); };
  const ConvSrc none = mk_src(nullptr, 0, 1, 1, 0);
  // Convolution activations stay NCHW: a producer thread of the tensor-core kernel owns one pixel, so its 16
  // per-chunk channel loads are coalesced across the warp and go registers -> TMEM without a shuffle.  Row-major
  // ("NHWC") sources are used by the linear layers only.
  constexpr int NCHW = 0, NHWC = 1;
  const int B = d.B, h = d.h, w = d.w, H2 = d.H2, W2 = d.W2;

  // tensor-core mode: one tiny launch re-lays every weight matrix as its UMMA operand image (hi/lo split)
  const float* wp[kNumTcLayers];
  for (int i = 0; i < kNumTcLayers; ++i) wp[i] = nullptr;
  const bool bf16 = cfg->precision == TPSPP_HEAD_BF16;   // convolutions with bf16 operands; linear layers stay 3xTF32
  const bool tc = cfg->precision == TPSPP_HEAD_TC || bf16;
  const bool weights_cached = (cfg->flags & TPSPP_HEAD_FLAG_WEIGHTS_CACHED) != 0;
  // operand mode per convolution: the four 1x1 layers of the fused down kernel keep the 3xTF32 images it reads (they
  // are HBM-bound); the ten 3x3 layers run the tf32 + bf16-correction form unless the A/B flag asks for 3xTF32
  int cmode[14];
  for (int i = 0; i < 14; ++i) {
    const bool one = kConvLayers[i].KS == 1;
    cmode[i] = (bf16 && !one) ? CM_BF16 : ((one || (cfg->flags & TPSPP_HEAD_FLAG_TF32X3_CONV)) ? CM_TF32X3 : CM_MIX);
  }
  // (bf16 mode: the four 1x1 layers stay 3xTF32 inside the fused down kernel -- it is HBM-bound, so rounding its operands
  //  buys nothing; run unfused they took 0.34 ms against the fused kernel's 0.27)
  if (tc) {
    WPrepLayer L[kNumTcLayers];
    float* cur = W(TPSPP_WS_WPREP);
    for (int i = 0; i < kNumTcLayers; ++i) {
      L[i].w = P[kConvLayers[i].w_idx]; L[i].out = cur; L[i].Ctot = kConvLayers[i].Ctot;
      L[i].taps = kConvLayers[i].KS * kConvLayers[i].KS; L[i].N = kConvLayers[i].N; L[i].NT = kConvLayers[i].NT; L[i].scale = nullptr; L[i].dg_cin = 0; L[i].dg_ci0 = 0;
      L[i].bf16 = i < 14 ? cmode[i] : CM_TF32X3;
      wp[i] = cur;
      cur += conv_tc_wprep_floats(kConvLayers[i].Ctot, kConvLayers[i].KS, kConvLayers[i].N);
    }
    if (!weights_cached) {     // else: the images of the previous call with these weights are still in the workspace
      rc = conv_tc_prepare_weights(L, kNumTcLayers, st);
      if (rc != TPSPP_OK) return rc;
    }
  }
#define RUN(...) do { rc = run_conv(__VA_ARGS__); if (rc != TPSPP_OK) return rc; } while (0)
  // down0/1/2 + grid()/down_feat (tps_pp.py:581-583,585,560-562): one fused tensor-core kernel in the 3xTF32 mode; f0/f1/f2
  // are still written (once) for the stride-2 convolutions and the MSFA encoder
  bool fused_down = false;
  if (tc && !(cfg->flags & TPSPP_HEAD_FLAG_UNFUSED_DOWN)) {
    rc = run_down_fused(x, o0, o1, wp[0], wp[1], wp[2], wp[5], P[TPSPP_P_DOWN0_B], P[TPSPP_P_DOWN1_B], P[TPSPP_P_DOWN2_B],
                        P[TPSPP_P_DOWNFEAT_B], W(TPSPP_WS_F0), W(TPSPP_WS_F1), W(TPSPP_WS_F2), feat_grid, B, h, w, st, bf16 ? 1 : 0,
                        (bf16 && (cfg->flags & TPSPP_HEAD_FLAG_FEATGRID_BF16)) ? 1 : 0);
    if (rc < 0) return rc;
    fused_down = rc == TPSPP_OK;
    TPSPP_REQUIRE(fused_down || !bf16, "head: the bf16 mode needs the fused down kernel's geometry (width 64, 16-byte aligned inputs)");
  }
  // bf16 mode: the LARGE head-internal activations are stored as bf16 (f0, f1, f2, a0, a1, e0, d2 -- 85 % of the intermediate
  // bytes); the small deep maps (e1, e2, e3, cbam, d0, d1), de (DGAB's input) and everything after it stay fp32
  const int bs = (bf16 && fused_down) ? 1 : 0;
  if (!fused_down) {
  RUN(1, mk_src(o0, 32, H2, W2, NCHW), none, none, P[TPSPP_P_DOWN0_W], P[TPSPP_P_DOWN0_B], nullptr, W(TPSPP_WS_F0), NCHW, B, H2, W2, 1, 1, st, wp[0], cmode[0]);
  RUN(1, mk_src(o1, 32, H2, W2, NCHW), none, none, P[TPSPP_P_DOWN1_W], P[TPSPP_P_DOWN1_B], nullptr, W(TPSPP_WS_F1), NCHW, B, H2, W2, 1, 1, st, wp[1], cmode[1]);
  RUN(1, mk_src(x, 64, h, w, NCHW), none, none, P[TPSPP_P_DOWN2_W], P[TPSPP_P_DOWN2_B], nullptr, W(TPSPP_WS_F2), NCHW, B, h, w, 1, 1, st, wp[2], cmode[2]);
  }
  // down0_1 / down1_1: 3x3 stride 2 (tps_pp.py:584)
  RUN(3, mk_src(W(TPSPP_WS_F0), 64, H2, W2, NCHW, 1, 1, bs), none, none, P[TPSPP_P_DOWN0_1_W], P[TPSPP_P_DOWN0_1_B], nullptr, W(TPSPP_WS_A0), NCHW, B, h, w, 2, 2, st, wp[3], cmode[3], bs);
  RUN(3, mk_src(W(TPSPP_WS_F1), 64, H2, W2, NCHW, 1, 1, bs), none, none, P[TPSPP_P_DOWN1_1_W], P[TPSPP_P_DOWN1_1_B], nullptr, W(TPSPP_WS_A1), NCHW, B, h, w, 2, 2, st, wp[4], cmode[4], bs);
  // grid(): down_feat(cat(f0, f1, up2(f2))) (tps_pp.py:560-562,585) -> feat_grid in the boundary layout (warp input)
  if (!fused_down)
  RUN(1, mk_src(W(TPSPP_WS_F0), 64, H2, W2, NCHW), mk_src(W(TPSPP_WS_F1), 64, H2, W2, NCHW), mk_src(W(TPSPP_WS_F2), 64, h, w, NCHW, 2, 2),
      P[TPSPP_P_DOWNFEAT_W], P[TPSPP_P_DOWNFEAT_B], nullptr, feat_grid, NCHW, B, H2, W2, 1, 1, st, wp[5], cmode[5]);
  // MSFA encoder (tps_pp.py:158-160): cat(a0, a1, f2) -> e0 -> e1 -> e2 -> e3
  RUN(3, mk_src(W(TPSPP_WS_A0), 64, h, w, NCHW, 1, 1, bs), mk_src(W(TPSPP_WS_A1), 64, h, w, NCHW, 1, 1, bs), mk_src(W(TPSPP_WS_F2), 64, h, w, NCHW, 1, 1, bs),
      P[TPSPP_P_ENC0_W], P[TPSPP_P_ENC0_B], nullptr, W(TPSPP_WS_E0), NCHW, B, h, w, 1, 1, st, wp[6], cmode[6], bs);
  RUN(3, mk_src(W(TPSPP_WS_E0), 64, h, w, NCHW, 1, 1, bs), none, none, P[TPSPP_P_ENC1_W], P[TPSPP_P_ENC1_B], nullptr, W(TPSPP_WS_E1), NCHW, B, d.h1, d.w1, 2, 2, st, wp[7], cmode[7]);
  RUN(3, mk_src(W(TPSPP_WS_E1), 64, d.h1, d.w1, NCHW), none, none, P[TPSPP_P_ENC2_W], P[TPSPP_P_ENC2_B], nullptr, W(TPSPP_WS_E2), NCHW, B, d.h2, d.w2, d.ps, d.ps, st, wp[8], cmode[8]);
  RUN(3, mk_src(W(TPSPP_WS_E2), 64, d.h2, d.w2, NCHW), none, none, P[TPSPP_P_ENC3_W], P[TPSPP_P_ENC3_B], nullptr, W(TPSPP_WS_E3), NCHW, B, d.py, d.px, 2, 1, st, wp[9], cmode[9]);
  // CBAM on the deepest map (tps_pp.py:163)
  launch_k(cbam_kernel, dim3(B), dim3(256), 0, st, (const float*)W(TPSPP_WS_E3), W(TPSPP_WS_CBAM), P[TPSPP_P_CBAM_MLP0_W], P[TPSPP_P_CBAM_MLP2_W],
           P[TPSPP_P_CBAM_SP_W], P[TPSPP_P_CBAM_SP_B], d.py, d.px);
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());
  // decoder (tps_pp.py:165-168): upsample + conv + skip
  RUN(3, mk_src(W(TPSPP_WS_CBAM), 64, d.py, d.px, NCHW, 2, 1), none, none, P[TPSPP_P_DEC0_W], P[TPSPP_P_DEC0_B], W(TPSPP_WS_E2), W(TPSPP_WS_D0), NCHW, B, d.h2, d.w2, 1, 1, st, wp[10], cmode[10]);
  RUN(3, mk_src(W(TPSPP_WS_D0), 64, d.h2, d.w2, NCHW, d.ps, d.ps), none, none, P[TPSPP_P_DEC1_W], P[TPSPP_P_DEC1_B], W(TPSPP_WS_E1), W(TPSPP_WS_D1), NCHW, B, d.h1, d.w1, 1, 1, st, wp[11], cmode[11]);
  RUN(3, mk_src(W(TPSPP_WS_D1), 64, d.h1, d.w1, NCHW, 2, 2), none, none, P[TPSPP_P_DEC2_W], P[TPSPP_P_DEC2_B], W(TPSPP_WS_E0), W(TPSPP_WS_D2), NCHW, B, h, w, 1, 1, st, wp[12], cmode[12], bs, bs);
  RUN(3, mk_src(W(TPSPP_WS_D2), 64, h, w, NCHW, 1, 1, bs), none, none, P[TPSPP_P_DEC3_W], P[TPSPP_P_DEC3_B], nullptr, W(TPSPP_WS_DE), NCHW, B, h, w, 1, 1, st, wp[13], cmode[13]);
#undef RUN
  // localisation + p_linear (tps_pp.py:321-323, 305)
  {
    LocArgs a;
    a.e3 = W(TPSPP_WS_E3);
    a.wa = P[TPSPP_P_LOC1A_W]; a.ba = P[TPSPP_P_LOC1A_B]; a.wb = P[TPSPP_P_LOC1B_W]; a.bb = P[TPSPP_P_LOC1B_B];
    a.wc = P[TPSPP_P_LOC2_W]; a.bc = P[TPSPP_P_LOC2_B];
    a.wp0 = P[TPSPP_P_PLIN0_W]; a.bp0 = P[TPSPP_P_PLIN0_B]; a.wp1 = P[TPSPP_P_PLIN1_W]; a.bp1 = P[TPSPP_P_PLIN1_B];
    a.c_prime = c_prime; a.p1 = W(TPSPP_WS_P1); a.F = d.F;
    a.p1img = (tc && d.F == 32) ? W(TPSPP_WS_P1IMG) : nullptr;
    const size_t smem = (size_t)(d.F * 65 + d.F * 256 + 2 * d.F + d.F * 33) * sizeof(float);
    launch_k(loc_p1_kernel, dim3(B), dim3(256), smem, st, a);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
  }
  // DGAB (DGAB.py:74-77)
  {
    DgabArgs a;
    a.x = W(TPSPP_WS_DE); a.e3 = W(TPSPP_WS_E3);
    a.n1w = P[TPSPP_P_NORM1_W]; a.n1b = P[TPSPP_P_NORM1_B]; a.n2w = P[TPSPP_P_NORM2_W]; a.n2b = P[TPSPP_P_NORM2_B];
    a.wh = P[TPSPP_P_MLP_H_W]; a.ww = P[TPSPP_P_MLP_W_W]; a.wp = P[TPSPP_P_PROJ_W]; a.bp = P[TPSPP_P_PROJ_B];
    a.x1 = W(TPSPP_WS_X1); a.v = W(TPSPP_WS_V); a.H = h; a.F = d.F;
    {      // warp-per-plane kernel (head_dims admits h = 8 or 16 and F <= 32 only)
      TPSPP_REQUIRE((h == 8 || h == 16) && d.F <= 32 && dgab_warp_smem(h, d.F) <= 220 * 1024, "DGAB: unsupported geometry h=%d F=%d", h, d.F);
      const size_t smem = dgab_warp_smem(h, d.F);
      auto kern = h == 8 ? dgab_warp_kernel<8> : dgab_warp_kernel<16>;
      TPSPP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int grid = sm_count();
      if (grid > (B * 64 + DW_WARPS - 1) / DW_WARPS) grid = (B * 64 + DW_WARPS - 1) / DW_WARPS;
      launch_k(kern, dim3(grid), dim3(DW_WARPS * 32), smem, st, a, B * 64);
    }
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
    const long long R = (long long)B * 64 * h;
    if (tc) {
      // Mlp over the width axis on the tensor cores: rows = (b, c, h), K = w.  Fused kernel first (hidden tensor stays
      // in tensor memory); the two-launch form is the fallback
      rc = run_mlp_fused(W(TPSPP_WS_V), W(TPSPP_WS_X1), wp[TCL_FC1], P[TPSPP_P_FC1_B], wp[TCL_FC2], P[TPSPP_P_FC2_B],
                         W(TPSPP_WS_DE2), R, st);
      if (rc < 0) return rc;
      if (rc == 1) {
      ConvArgs g;
      memset(&g, 0, sizeof(g));
      g.src[0] = mk_src(W(TPSPP_WS_V), 64, 1, (int)R, NHWC); g.src[1] = none; g.src[2] = none;
      g.B = 1; g.Ho = 1; g.Wo = (int)R; g.Ctot = 64; g.sh = 1; g.sw = 1; g.pad = 0; g.out_nhwc = 1;
      g.bias = P[TPSPP_P_FC1_B]; g.skip = nullptr; g.out = W(TPSPP_WS_HID); g.act = CONV_ACT_GELU; g.act_scale = 1.f;
      g.Cout = 256; g.wimg_stride = 0; g.weight = nullptr;
      rc = run_conv_tc(1, g, wp[TCL_FC1], 64, st);
      if (rc != TPSPP_OK) return rc;
      g.src[0] = mk_src(W(TPSPP_WS_HID), 256, 1, (int)R, NHWC); g.Ctot = 256;
      g.bias = P[TPSPP_P_FC2_B]; g.skip = W(TPSPP_WS_X1); g.out = W(TPSPP_WS_DE2); g.act = CONV_ACT_NONE; g.Cout = 64;
      rc = run_conv_tc(1, g, wp[TCL_FC2], 64, st);
      if (rc != TPSPP_OK) return rc;
      }
    } else {
      MlpArgs m;
      m.v = W(TPSPP_WS_V); m.x1 = W(TPSPP_WS_X1); m.w1 = P[TPSPP_P_FC1_W]; m.b1 = P[TPSPP_P_FC1_B];
      m.w2 = P[TPSPP_P_FC2_W]; m.b2 = P[TPSPP_P_FC2_B]; m.out = W(TPSPP_WS_DE2); m.R = R;
      dgab_mlp_kernel<<<(unsigned)((m.R + 127) / 128), 256, ML_SMEM, st>>>(m);
      count_launch();
      TPSPP_CHECK_CUDA(cudaGetLastError());
    }
  }
  // attention score (tps_pp.py:303-312)
  bool fused_score = false;
  if (tc && d.F == 32 && !(cfg->flags & TPSPP_HEAD_FLAG_UNFUSED_SCORE)) {
    // feat_linear.0 -> feat_linear.1 -> tanh(QK^T / 8) chained through tensor memory in one kernel
    rc = run_score_fused(W(TPSPP_WS_DE2), wp[TCL_FLIN0], wp[TCL_FLIN1], W(TPSPP_WS_P1IMG), P[TPSPP_P_FLIN0_B], P[TPSPP_P_FLIN1_B],
                         pc_score, B, h, w, d.F, 0.125f, st);     // 64^-0.5 (tps_pp.py:247)
    if (rc < 0) return rc;
    fused_score = rc == TPSPP_OK;
  }
  if (fused_score) {
  } else if (tc && d.F == 32) {
    // three chained 1x1 contractions on the tensor cores: feat_linear.0, feat_linear.1, then the
    // "QK^T" with per-image weights p1[b] and the tanh(64^-0.5 * .) epilogue
    const int n = h * w;
    ConvArgs g;
    memset(&g, 0, sizeof(g));
    g.src[1] = none; g.src[2] = none;
    g.sh = 1; g.sw = 1; g.pad = 0; g.out_nhwc = 1; g.skip = nullptr; g.act_scale = 1.f; g.wimg_stride = 0;
    g.src[0] = mk_src(W(TPSPP_WS_DE2), 64, h, w, NCHW); g.B = B; g.Ho = h; g.Wo = w; g.Ctot = 64;
    g.bias = P[TPSPP_P_FLIN0_B]; g.out = W(TPSPP_WS_T1); g.act = CONV_ACT_NONE; g.Cout = 32;
    rc = run_conv_tc(1, g, wp[TCL_FLIN0], 32, st);
    if (rc != TPSPP_OK) return rc;
    g.src[0] = mk_src(W(TPSPP_WS_T1), 32, h, w, NHWC); g.Ctot = 32;
    g.bias = P[TPSPP_P_FLIN1_B]; g.out = W(TPSPP_WS_FS); g.Cout = 128;
    rc = run_conv_tc(1, g, wp[TCL_FLIN1], 64, st);
    if (rc != TPSPP_OK) return rc;
    g.src[0] = mk_src(W(TPSPP_WS_FS), 128, h, w, NHWC); g.Ctot = 128;
    g.bias = nullptr; g.out = pc_score; g.Cout = 32; g.act = CONV_ACT_TANH; g.act_scale = 0.125f;   // 64^-0.5 (tps_pp.py:247)
    g.wimg_stride = 8192;
    (void)n;
    rc = run_conv_tc(1, g, W(TPSPP_WS_P1IMG), 32, st);
    if (rc != TPSPP_OK) return rc;
  } else {
    ScoreArgs a;
    a.de2 = W(TPSPP_WS_DE2); a.p1 = W(TPSPP_WS_P1);
    a.wf0 = P[TPSPP_P_FLIN0_W]; a.bf0 = P[TPSPP_P_FLIN0_B]; a.wf1 = P[TPSPP_P_FLIN1_W]; a.bf1 = P[TPSPP_P_FLIN1_B];
    a.score = pc_score; a.n = h * w; a.F = d.F; a.scale = 0.125f;   // 64^-0.5 (tps_pp.py:247)
    dim3 grid((a.n + 127) / 128, B);
    score_kernel<<<grid, 256, SC_SMEM, st>>>(a);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
  }
  return TPSPP_OK;
}
