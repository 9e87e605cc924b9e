// Training-path linear layers of the head (north_star (4), SURVEY K-p): forward and backward of the reference's
// `nn.Linear` / `torch.bmm` calls (DGAB.py:11-23,28-36,52; tps_pp.py:250-273,293-312) as native kernels, so that autograd of
// the rectifier's dense layers no longer goes through cuBLAS.
//
//   forward   y = x w^T + b                 row-major [R,K] x [N,K]^T: lin_tma_kernel (tcgen05, 3xTF32 split)
//   backward  gx = gy w                     the same kernel over the transposed weight image
//             gw[n][k] = sum_r gy[r][n] x[r][k],  gb[n] = sum_r gy[r][n]      wgrad_tc_kernel<ROWS> (conv_train.cu)
//   weight_batches > 1: torch.bmm(x, w^T) with one [N,K] weight per group of rows (the QK^T score of get_score).
// Shapes the tensor-core kernels do not take (R not a multiple of 128, K or N not a multiple of 32: DGAB's axial mlp_w /
// mlp_h, localization_fc1.2, CBAM's channel MLP) run on the fp32 CUDA-core kernels below -- they are < 1 % of the FLOPs.
#include "head.cuh"
#include "tc.cuh"

#include <string.h>

namespace tpspp {

constexpr int TC_TM = 128, TC_KC = 32;      // row tile and k chunk of the tcgen05 linear kernel (head_tc.cu)

// ---- operand image of a row-major weight for the tcgen05 linear kernel (same layout as wprep_kernel's fp32 hi|lo pair):
//      image row n' / column k' = w[n'][k'] (forward) or w[k'][n'] (transposed: the data-gradient GEMM), one image per batch ----
__global__ void __launch_bounds__(256) lin_wprep_kernel(const float* __restrict__ w, float* __restrict__ out, int Nn, int Kk, int NT,
                                                        int transposed, int batches) {
  const int Npad = (Nn + NT - 1) / NT * NT;
  const int per = Npad * Kk, nchunks = Kk >> 5;
  const long long total = (long long)per * batches;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int b = (int)(i / per), r = (int)(i - (long long)b * per), n = r / Kk, k = r - n * Kk;
    const float* wb = w + (size_t)b * Nn * Kk;
    float v = 0.f;
    if (n < Nn) v = transposed ? __ldg(wb + (size_t)k * Nn + n) : __ldg(wb + (size_t)n * Kk + k);
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    const int blk = n / NT, nn = n - blk * NT;
    float* o = out + (size_t)b * 2 * per + ((size_t)(blk * nchunks + (k >> 5)) * 2) * (NT * 32) + ((k >> 2) & 7) * (NT * 4) + nn * 4 + (k & 3);
    o[0] = hi;
    o[NT * 32] = v - hi;
  }
}

// ---- fp32 CUDA-core kernels for the small / odd shapes ----
// y[r][n] = sum_k x[r][k] w[b][n][k] (+ bias): 32 x 32 output tile per block, 32-wide k chunks through padded shared tiles
// (coalesced global reads for either weight orientation), thread = one column x four rows
__global__ void __launch_bounds__(256) lin_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ y, int K, int N,
                                                            long long rows_per_batch, int transposed,
                                                            const float* __restrict__ residual = nullptr, int gelu = 0) {
  // transposed = 0: y = x w^T (w [N][K]);  1: y = x w (w [K][N], the data gradient with K/N swapped by the caller)
  __shared__ float Xs[32][33], Ws[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int b = blockIdx.z, n0 = blockIdx.x * 32;
  const long long rb = (long long)b * rows_per_batch, r0 = rb + (long long)blockIdx.y * 32, rend = rb + rows_per_batch;
  const float* wb = w + (size_t)b * N * K;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rr = ty + 8 * j;
      Xs[rr][tx] = (r0 + rr < rend && k0 + tx < K) ? __ldg(x + (size_t)(r0 + rr) * K + k0 + tx) : 0.f;
      if (!transposed) Ws[rr][tx] = (n0 + rr < N && k0 + tx < K) ? __ldg(wb + (size_t)(n0 + rr) * K + k0 + tx) : 0.f;
      else Ws[tx][rr] = (n0 + tx < N && k0 + rr < K) ? __ldg(wb + (size_t)(k0 + rr) * N + n0 + tx) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float wv = Ws[tx][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(Xs[ty + 8 * j][k], wv, acc[j]);
    }
    __syncthreads();
  }
  if (n0 + tx >= N) return;
  const float bv = bias != nullptr ? __ldg(bias + n0 + tx) : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (r0 + ty + 8 * j < rend) {
      const size_t o = (size_t)(r0 + ty + 8 * j) * N + n0 + tx;
      float v = acc[j] + bv;
      if (gelu) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
      if (residual != nullptr) v += residual[o];
      y[o] = v;
    }
}
// gw[b][n][k] = sum over the batch's rows of gy[r][n] x[r][k]; gb[n] = sum_r gy[r][n] (column K of the same grid).
// block = 32 features (k) x 8 row lanes for one output feature n; fixed summation order (deterministic)
__global__ void __launch_bounds__(256) lin_small_wgrad_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                                              float* __restrict__ gw, float* __restrict__ gb, long long rows_per_batch,
                                                              int K, int N) {
  const int n = blockIdx.y, b = blockIdx.z;
  const int kx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + kx;                       // k == K: the bias column
  const long long r0 = (long long)b * rows_per_batch;
  float acc = 0.f;
  if (k <= K) {
    for (long long r = r0 + ry; r < r0 + rows_per_batch; r += 8) {
      const float g = __ldg(gy + (size_t)r * N + n);
      acc = fmaf(g, k < K ? __ldg(x + (size_t)r * K + k) : 1.f, acc);
    }
  }
  __shared__ float red[8][33];
  red[ry][kx] = acc;
  __syncthreads();
  if (ry == 0) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += red[j][kx];
    if (k < K) gw[((size_t)b * N + n) * K + k] = s;
    else if (k == K && gb != nullptr) gb[n] = s;
  }
}

struct LinDims { long long R; int K, N, batches; long long rpb; bool tc_fwd, tc_dx, tc_wg; int nt_fwd, nt_dx, sk_fwd; };

// y[r][n] = sum_s part[s][r][n] + bias[n]: the reduction of a split-K forward (fixed order: deterministic)
__device__ __forceinline__ float lin_gelu(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__global__ void __launch_bounds__(256) lin_splitk_reduce_kernel(const float4* __restrict__ part, const float* __restrict__ bias,
                                                                const float4* __restrict__ residual, float4* __restrict__ y,
                                                                long long total4, int n4, int splits, int gelu) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total4; i += (long long)gridDim.x * 256) {
    float4 acc = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias) + (i % n4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int s = 0; s < splits; ++s) {
      const float4 p = __ldg(part + (size_t)s * total4 + i);
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    if (gelu) { acc.x = lin_gelu(acc.x); acc.y = lin_gelu(acc.y); acc.z = lin_gelu(acc.z); acc.w = lin_gelu(acc.w); }
    if (residual != nullptr) {
      const float4 r = residual[i];            // may alias y (x = x + f(x)): read, then write the same element
      acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
    }
    y[i] = acc;
  }
}
// The same reduction with the NEXT sub-layer's LayerNorm attached (a pre-norm transformer layer reads LN(x) right after
// x += f(.)): one warp per row keeps the row in registers -- y = sum_s part[s] + bias + residual, y_ln = LN(y) * w + b
// (two-pass mean / variance in fp32, biased variance like torch).  splits == 0: y is already complete (unsplit linear), LN only.
constexpr int LRL_WARPS = 4;
// PER = float4 per lane the row is unrolled for (>= N / 128), SPLITS = partial sums (compile-time: every load of a row --
// partials, bias, residual, LayerNorm parameters -- is issued before the first add: one L2 round trip instead of one per split)
template <int PER, int SPLITS>
__global__ void __launch_bounds__(LRL_WARPS * 32) lin_reduce_ln_kernel(const float4* __restrict__ part, const float* __restrict__ bias,
                                                                       const float4* __restrict__ residual, float4* __restrict__ y,
                                                                       const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                                       float eps, float4* __restrict__ y_ln, long long R, int N) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int n4 = N >> 2, per = n4 >> 5;            // float4 per lane (N % 128 == 0)
  const long long total4 = R * n4;
  const float inv_n = 1.f / (float)N;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = (long long)blockIdx.x * LRL_WARPS + (threadIdx.x >> 5); r < R; r += (long long)gridDim.x * LRL_WARPS) {
    const long long i0 = r * n4 + lane;
    float4 v[PER], w[PER], b[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (j < per) {
        w[j] = __ldg(reinterpret_cast<const float4*>(ln_w) + j * 32 + lane);
        b[j] = ln_b != nullptr ? __ldg(reinterpret_cast<const float4*>(ln_b) + j * 32 + lane) : zero4;
      }
    }
    if (SPLITS > 0) {
      // SB splits of the whole row are in flight at a time (all of them unless that would take more than 16 float4 per lane)
      constexpr int SP = SPLITS > 0 ? SPLITS : 1, SB = PER * SP > 16 ? (16 / PER > 0 ? 16 / PER : 1) : SP;
      float4 res[PER];
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        if (j < per) {
          v[j] = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias) + j * 32 + lane) : zero4;
          res[j] = residual != nullptr ? residual[i0 + j * 32] : zero4;      // may alias y: read before the write below
        }
      }
#pragma unroll
      for (int s0 = 0; s0 < SP; s0 += SB) {
        float4 p[SB][PER];
#pragma unroll
        for (int j = 0; j < PER; ++j)
          if (j < per) {
#pragma unroll
            for (int s = 0; s < SB; ++s)
              if (s0 + s < SP) p[s][j] = __ldg(part + (size_t)(s0 + s) * total4 + i0 + j * 32);
          }
#pragma unroll
        for (int j = 0; j < PER; ++j)
          if (j < per) {
#pragma unroll
            for (int s = 0; s < SB; ++s)       // bias, split 0, 1, ...: the order of lin_splitk_reduce_kernel
              if (s0 + s < SP) { v[j].x += p[s][j].x; v[j].y += p[s][j].y; v[j].z += p[s][j].z; v[j].w += p[s][j].w; }
          }
      }
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        if (j < per) {
          if (residual != nullptr) { v[j].x += res[j].x; v[j].y += res[j].y; v[j].z += res[j].z; v[j].w += res[j].w; }
          y[i0 + j * 32] = v[j];
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < PER; ++j)
        if (j < per) v[j] = y[i0 + j * 32];
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j)
      if (j < per) sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_n;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (j < per) {
        const float a0 = v[j].x - mean, a1 = v[j].y - mean, a2 = v[j].z - mean, a3 = v[j].w - mean;
        sq += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * inv_n + eps);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (j < per) {
        float4 o4;
        o4.x = (v[j].x - mean) * rstd * w[j].x + b[j].x; o4.y = (v[j].y - mean) * rstd * w[j].y + b[j].y;
        o4.z = (v[j].z - mean) * rstd * w[j].z + b[j].z; o4.w = (v[j].w - mean) * rstd * w[j].w + b[j].w;
        y_ln[i0 + j * 32] = o4;
      }
    }
  }
}
// splits in {0, 1, 2, 4} (what lin_dims produces; 0 = y already complete)
static void launch_reduce_ln(const float4* part, const float* bias, const float4* residual, float4* y, const float* ln_w, const float* ln_b,
                             float eps, float4* y_ln, long long R, int N, int splits, cudaStream_t st) {
  const dim3 grid((unsigned)min((R + LRL_WARPS - 1) / LRL_WARPS, 8LL * sm_count())), block(LRL_WARPS * 32);
  const int per = N / 128;
#define TPSPP_LRL(PER, SP) launch_k(lin_reduce_ln_kernel<PER, SP>, grid, block, 0, st, part, bias, residual, y, ln_w, ln_b, eps, y_ln, R, N)
#define TPSPP_LRL_S(PER) do { if (splits == 0) TPSPP_LRL(PER, 0); else if (splits == 1) TPSPP_LRL(PER, 1); else if (splits == 2) TPSPP_LRL(PER, 2); \
                              else TPSPP_LRL(PER, 4); } while (0)
  if (per <= 2) TPSPP_LRL_S(2);
  else if (per <= 4) TPSPP_LRL_S(4);
  else TPSPP_LRL_S(8);
#undef TPSPP_LRL_S
#undef TPSPP_LRL
}
static bool lin_tc_shape(long long R, int K, int N, long long rpb, int* NT) {
  if (R % TC_TM || rpb % TC_TM || K % TC_KC || N % 32 || N > 4096 || K > 4096 || R > 0x7fffffffLL) return false;
  *NT = N % 64 == 0 ? 64 : 32;
  // K <= 64 (A stays in tensor memory across the column blocks) or a single column block: the TMA-fed lin_tma_kernel;
  // anything else (transformer-sized layers, e.g. 512 -> 1536): run_conv_tc drops to the shared-memory-operand kernel with
  // one CTA per (128-row tile, column block) -- still tcgen05 / 3xTF32, just without the TMA pipeline
  return true;
}
static int lin_dims(const tpspp_linear_cfg* c, LinDims* d) {
  TPSPP_REQUIRE(c != nullptr, "linear cfg is NULL");
  TPSPP_REQUIRE(c->rows >= 0 && c->in_features > 0 && c->out_features > 0, "linear: rows >= 0, in/out features > 0");
  TPSPP_REQUIRE(c->weight_batches >= 1 && c->weight_batches <= 65535, "linear: weight_batches must be in 1..65535");
  TPSPP_REQUIRE(c->rows % c->weight_batches == 0, "linear: rows must split evenly over the weight batches");
  TPSPP_REQUIRE(c->rows / c->weight_batches < 65535LL * 32, "linear: too many rows per weight batch");
  TPSPP_REQUIRE(c->rows * (long long)(c->in_features > c->out_features ? c->in_features : c->out_features) < (1LL << 40), "linear: too large");
  d->R = c->rows; d->K = c->in_features; d->N = c->out_features; d->batches = c->weight_batches;
  d->rpb = d->batches > 0 && d->R > 0 ? d->R / d->batches : 1;
  d->nt_fwd = d->nt_dx = 64;
  d->tc_fwd = d->R > 0 && lin_tc_shape(d->R, d->K, d->N, d->rpb, &d->nt_fwd);
  d->tc_dx = d->R > 0 && lin_tc_shape(d->R, d->N, d->K, d->rpb, &d->nt_dx);
  // split-K of the forward: few row tiles (inference steps of a transformer decoder) and a long K
  d->sk_fwd = 1;
  if (d->tc_fwd && d->batches == 1 && d->K >= 256) {
    const long long ctas = (d->R / TC_TM) * ((d->N + d->nt_fwd - 1) / d->nt_fwd);
    const int chunks = d->K / TC_KC;
    int sk = 1;
    while (sk < 4 && ctas * sk * 2 <= 2LL * sm_count() && chunks % (sk * 2) == 0 && chunks / (sk * 2) >= 2) sk *= 2;
    d->sk_fwd = sk;
  }
  d->tc_wg = d->R >= 2048 && d->rpb % 32 == 0 && d->K <= 1024 && d->N <= 1024 && (d->batches == 1 || d->rpb >= 256);
  return TPSPP_OK;
}
enum { LW_WFWD = 0, LW_WDX, LW_WG, LW_SPLITK, LW_COUNT };
static void lin_offsets(const LinDims& d, size_t* off, size_t* total) {
  size_t sz[LW_COUNT];
  const int npf = (d.N + d.nt_fwd - 1) / d.nt_fwd * d.nt_fwd, npx = (d.K + d.nt_dx - 1) / d.nt_dx * d.nt_dx;
  sz[LW_WFWD] = d.tc_fwd ? (size_t)2 * npf * d.K * d.batches : 0;
  sz[LW_WDX] = d.tc_dx ? (size_t)2 * npx * d.N * d.batches : 0;
  sz[LW_WG] = d.tc_wg ? wgrad_rows_ws_floats(d.R, d.K, d.N, d.batches) : 0;
  sz[LW_SPLITK] = d.sk_fwd > 1 ? (size_t)d.sk_fwd * d.R * d.N : 0;
  size_t cur = 0;
  for (int i = 0; i < LW_COUNT; ++i) {
    off[i] = cur;
    cur += (sz[i] * sizeof(float) + 255) / 256 * 256;
  }
  *total = cur + 256;
}

// out[R, Nn] = in[R, Kk] . img^T (+ bias) through the tcgen05 row-major linear kernel
static int lin_tc_run(const float* in, const float* wimg, const float* bias, float* out, const LinDims& d, int Kk, int Nn, int NT,
                      cudaStream_t st, int splitk = 1, float* part = nullptr, const float* residual = nullptr, int gelu = 0,
                      const float* ln_w = nullptr, const float* ln_b = nullptr, float ln_eps = 0.f, float* y_ln = nullptr) {
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.src[0].ptr = in; a.src[0].C = Kk; a.src[0].H = 1; a.src[0].W = (int)d.rpb; a.src[0].uh = a.src[0].uw = 1; a.src[0].nhwc = 1;
  a.src[1].H = a.src[1].W = a.src[1].uh = a.src[1].uw = 1; a.src[2] = a.src[1];
  a.bias = bias; a.out = out; a.B = d.batches; a.Ho = 1; a.Wo = (int)d.rpb; a.Ctot = Kk; a.sh = a.sw = 1; a.pad = 0;
  a.out_nhwc = 1; a.act = CONV_ACT_NONE; a.act_scale = 1.f; a.Cout = Nn;
  a.wimg_stride = d.batches > 1 ? (long long)2 * ((Nn + NT - 1) / NT * NT) * Kk : 0;
  if (splitk > 1 && part != nullptr && ((uintptr_t)out & 15) == 0 && (bias == nullptr || ((uintptr_t)bias & 15) == 0) &&
      ((uintptr_t)residual & 15) == 0 && Nn % 4 == 0) {
    a.bias = nullptr; a.out = part; a.splitk = splitk;
    const int rc = run_conv_tc(1, a, wimg, NT, st, CM_TF32X3);
    if (rc != TPSPP_OK) return rc;
    const long long total4 = d.R * Nn / 4;
    if (y_ln != nullptr)
      launch_reduce_ln(reinterpret_cast<const float4*>(part), bias, reinterpret_cast<const float4*>(residual), reinterpret_cast<float4*>(out),
                       ln_w, ln_b, ln_eps, reinterpret_cast<float4*>(y_ln), d.R, Nn, splitk, st);
    else
      launch_k(lin_splitk_reduce_kernel, dim3((unsigned)min((total4 + 255) / 256, 4LL * sm_count())), dim3(256), 0, st,
               reinterpret_cast<const float4*>(part), bias, reinterpret_cast<const float4*>(residual), reinterpret_cast<float4*>(out), total4,
               Nn / 4, splitk, gelu);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
    return TPSPP_OK;
  }
  // unsplit: the kernel's own epilogue applies the activation (branch-free erf GELU, |error| <= 4.7e-7) and adds the residual
  a.skip = residual; a.skip_pre = 0;
  if (gelu) a.act = CONV_ACT_GELU;
  const int rc = run_conv_tc(1, a, wimg, NT, st, CM_TF32X3);
  if (rc != TPSPP_OK || y_ln == nullptr) return rc;
  launch_reduce_ln(nullptr, nullptr, nullptr, reinterpret_cast<float4*>(out), ln_w, ln_b, ln_eps, reinterpret_cast<float4*>(y_ln), d.R, Nn, 0, st);
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());
  return TPSPP_OK;
}

}  // namespace tpspp

using namespace tpspp;

extern "C" size_t tpspp_linear_workspace_bytes(const tpspp_linear_cfg* cfg) {
  LinDims d;
  if (lin_dims(cfg, &d) != TPSPP_OK) return 0;
  size_t off[LW_COUNT], total;
  lin_offsets(d, off, &total);
  return total;
}

extern "C" int tpspp_linear_fwd(const tpspp_linear_cfg* cfg, const float* x, const float* w, const float* bias, float* y,
                                void* workspace, tpspp_stream_t stream) {
  return tpspp_linear_fwd_ex(cfg, x, w, bias, nullptr, TPSPP_ACT_NONE, y, workspace, stream);
}

extern "C" int tpspp_linear_fwd_ex(const tpspp_linear_cfg* cfg, const float* x, const float* w, const float* bias, const float* residual,
                                   int32_t act, float* y, void* workspace, tpspp_stream_t stream) {
  return tpspp_linear_ln_fwd(cfg, x, w, bias, residual, act, y, nullptr, nullptr, 0.f, nullptr, workspace, stream);
}

extern "C" int tpspp_linear_ln_fwd(const tpspp_linear_cfg* cfg, const float* x, const float* w, const float* bias, const float* residual,
                                   int32_t act, float* y, const float* ln_weight, const float* ln_bias, float ln_eps, float* y_ln,
                                   void* workspace, tpspp_stream_t stream) {
  reset_launch_count();
  if (y_ln != nullptr) {
    TPSPP_REQUIRE(cfg != nullptr && ln_weight != nullptr && act == TPSPP_ACT_NONE && cfg->weight_batches == 1,
                  "tpspp_linear_ln_fwd: the LayerNorm output needs ln_weight, no activation and a single weight");
    TPSPP_REQUIRE(cfg->out_features % 128 == 0 && cfg->out_features <= 1024, "tpspp_linear_ln_fwd: out_features must be a multiple of 128, at most 1024");
    TPSPP_REQUIRE(y_ln != y && y_ln != residual && y_ln != x, "tpspp_linear_ln_fwd: y_ln must not alias x, y or residual");
    TPSPP_REQUIRE((((uintptr_t)y_ln | (uintptr_t)ln_weight | (uintptr_t)ln_bias | (uintptr_t)y | (uintptr_t)bias | (uintptr_t)residual) & 15) == 0,
                  "tpspp_linear_ln_fwd: y, y_ln, bias, residual and the LayerNorm parameters must be 16-byte aligned");
  }
  TPSPP_REQUIRE(act == TPSPP_ACT_NONE || act == TPSPP_ACT_GELU, "tpspp_linear_fwd: unknown activation %d", act);
  TPSPP_REQUIRE(cfg == nullptr || cfg->weight_batches <= 1 || (residual == nullptr && act == TPSPP_ACT_NONE),
                "tpspp_linear_fwd_ex: batched weights (bmm) take no residual / activation");
  LinDims d;
  int rc = lin_dims(cfg, &d);
  if (rc != TPSPP_OK) return rc;
  if (d.R == 0) return TPSPP_OK;
  TPSPP_REQUIRE(x && w && y && workspace, "tpspp_linear_fwd: null pointer");
  TPSPP_REQUIRE(d.batches == 1 || bias == nullptr, "tpspp_linear_fwd: batched weights (bmm) take no bias");
  cudaStream_t st = (cudaStream_t)stream;
  size_t off[LW_COUNT], total;
  lin_offsets(d, off, &total);
  if (d.tc_fwd && !(((uintptr_t)x | (uintptr_t)y | (uintptr_t)workspace) & 15)) {
    float* img = reinterpret_cast<float*>((char*)workspace + off[LW_WFWD]);
    if (!(cfg->flags & TPSPP_LINEAR_FLAG_WEIGHTS_CACHED)) {
      const long long tot = (long long)((d.N + d.nt_fwd - 1) / d.nt_fwd * d.nt_fwd) * d.K * d.batches;
      lin_wprep_kernel<<<(unsigned)min((tot + 255) / 256, 2048LL), 256, 0, st>>>(w, img, d.N, d.K, d.nt_fwd, 0, d.batches);
      count_launch();
      TPSPP_CHECK_CUDA(cudaGetLastError());
    }
    return lin_tc_run(x, img, bias, y, d, d.K, d.N, d.nt_fwd, st, d.sk_fwd, reinterpret_cast<float*>((char*)workspace + off[LW_SPLITK]),
                      residual, act == TPSPP_ACT_GELU ? 1 : 0, ln_weight, ln_bias, ln_eps, y_ln);
  }
  lin_small_fwd_kernel<<<dim3((d.N + 31) / 32, (unsigned)((d.rpb + 31) / 32), d.batches), 256, 0, st>>>(x, w, bias, y, d.K, d.N, d.rpb, 0,
                                                                                                         residual, act == TPSPP_ACT_GELU ? 1 : 0);
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());
  if (y_ln != nullptr) {
    launch_reduce_ln(nullptr, nullptr, nullptr, reinterpret_cast<float4*>(y), ln_weight, ln_bias, ln_eps, reinterpret_cast<float4*>(y_ln), d.R, d.N, 0, st);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
  }
  return TPSPP_OK;
}

extern "C" int tpspp_linear_bwd(const tpspp_linear_cfg* cfg, const float* x, const float* w, const float* gy, float* gx, float* gw,
                                float* gb, void* workspace, tpspp_stream_t stream) {
  reset_launch_count();
  LinDims d;
  int rc = lin_dims(cfg, &d);
  if (rc != TPSPP_OK) return rc;
  if (d.R == 0) return TPSPP_OK;
  TPSPP_REQUIRE(x && w && gy && workspace, "tpspp_linear_bwd: null pointer");
  TPSPP_REQUIRE(d.batches == 1 || gb == nullptr, "tpspp_linear_bwd: batched weights (bmm) have no bias gradient");
  cudaStream_t st = (cudaStream_t)stream;
  size_t off[LW_COUNT], total;
  lin_offsets(d, off, &total);
  if (gx != nullptr) {
    if (d.tc_dx && !(((uintptr_t)gy | (uintptr_t)gx | (uintptr_t)workspace) & 15)) {
      float* img = reinterpret_cast<float*>((char*)workspace + off[LW_WDX]);
      const long long tot = (long long)((d.K + d.nt_dx - 1) / d.nt_dx * d.nt_dx) * d.N * d.batches;
      lin_wprep_kernel<<<(unsigned)min((tot + 255) / 256, 2048LL), 256, 0, st>>>(w, img, d.K, d.N, d.nt_dx, 1, d.batches);
      count_launch();
      TPSPP_CHECK_CUDA(cudaGetLastError());
      rc = lin_tc_run(gy, img, nullptr, gx, d, d.N, d.K, d.nt_dx, st);
      if (rc != TPSPP_OK) return rc;
    } else {
      // gx[r][k] = sum_n gy[r][n] w[n][k]: the forward kernel with the roles of K and N swapped and the weight read transposed
      lin_small_fwd_kernel<<<dim3((d.K + 31) / 32, (unsigned)((d.rpb + 31) / 32), d.batches), 256, 0, st>>>(gy, w, nullptr, gx, d.N, d.K, d.rpb, 1);
      count_launch();
      TPSPP_CHECK_CUDA(cudaGetLastError());
    }
  }
  if (gw != nullptr) {
    if (d.tc_wg && !(((uintptr_t)workspace) & 15)) {
      rc = run_wgrad_rows(gy, x, gw, gb, d.R, d.K, d.N, d.batches, reinterpret_cast<float*>((char*)workspace + off[LW_WG]), st);
      if (rc != TPSPP_OK) return rc;
    } else {
      lin_small_wgrad_kernel<<<dim3((d.K + 1 + 31) / 32, d.N, d.batches), 256, 0, st>>>(gy, x, gw, gb, d.rpb, d.K, d.N);
      count_launch();
      TPSPP_CHECK_CUDA(cudaGetLastError());
    }
  } else if (gb != nullptr) {
    TPSPP_REQUIRE(false, "tpspp_linear_bwd: the bias gradient comes with the weight gradient (pass gw)");
  }
  return TPSPP_OK;
}
