// Backward of the fused warp (autograd of reference tps_pp.py:481-496,606-615 and
// tps_preprocessor.py:72-83,270-282; sampler rules from ATen grid_sampler_2d_backward).
//
//   K-B1 warp_bwd_sample_kernel : recompute grid -> d src (scatter-add) and d grid
//   K-B2 grid_bwd_reduce_kernel : d pc_score (elementwise) and partial dT = Phi^T . dgrid
//   K-B3 cprime_bwd_kernel      : dC' = inv_delta_C[:, :F]^T . dT
//
// Maths: with taps v_nw..v_se (0 outside the plane) and upstream G[c]
//   d/d ix = sum_c G[c] * (uy*(v_ne - v_nw) + ty*(v_se - v_sw))
//   d/d iy = sum_c G[c] * (ux*(v_sw - v_nw) + tx*(v_se - v_ne))
//   d/d gx = d/d ix * (W-1)/2 * [0 < ix_unclipped < W-1]          (same for y)
//   dT = Phi^T dgrid,  dC' = inv_delta_C[:, :F]^T dT,  dpc_score = theta * P_hat o (dgrid . T[3:]^T)
#include "common.cuh"

namespace tpspp {

template <typename FT>
__device__ __forceinline__ void atomic_addf(FT* p, float v);
template <>
__device__ __forceinline__ void atomic_addf<float>(float* p, float v) { atomicAdd(p, v); }
template <>
__device__ __forceinline__ void atomic_addf<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  atomicAdd(p, __float2bfloat16_rn(v));
}

template <typename FT>
__device__ __forceinline__ void bwd_one_source(const FT* __restrict__ src, const FT* __restrict__ gout,
                                               FT* __restrict__ gsrc, const TapsGrad& g, int c_lo, int c_hi,
                                               size_t plane, int n, float& gix, float& giy) {
  const int dx = g.t.dx, dy = g.t.dy;
  for (int c = c_lo; c < c_hi; ++c) {
    const float G = ldf(gout + (size_t)c * n);
    const FT* s = src + (size_t)c * plane + g.t.off;
    const float vnw = ldf(s);
    const float vne = dx ? ldf(s + dx) : 0.f;
    const float vsw = dy ? ldf(s + dy) : 0.f;
    const float vse = (dx && dy) ? ldf(s + dy + dx) : 0.f;
    gix = __fmaf_rn(G, g.uy * (vne - vnw) + g.ty * (vse - vsw), gix);
    giy = __fmaf_rn(G, g.ux * (vsw - vnw) + g.tx * (vse - vne), giy);
    if (gsrc != nullptr) {
      FT* d = gsrc + (size_t)c * plane + g.t.off;
      atomic_addf<FT>(d, G * g.t.w[0]);
      if (dx) atomic_addf<FT>(d + dx, G * g.t.w[1]);
      if (dy) atomic_addf<FT>(d + dy, G * g.t.w[2]);
      if (dx && dy) atomic_addf<FT>(d + dy + dx, G * g.t.w[3]);
    }
  }
}

template <typename FT, int MODE>
__global__ void __launch_bounds__(256) warp_bwd_sample_kernel(WarpParams p, int cchunk, int use_atomic_grid) {
  extern __shared__ double Tsm[];
  const int b = blockIdx.y;
  compute_T(p, b, Tsm, threadIdx.x, blockDim.x);
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= p.n) return;
  double gx, gy;
  pixel_grid<MODE>(p, Tsm, b, pix, gx, gy);
  const int cb = blockIdx.z * cchunk;
  float ggx = 0.f, ggy = 0.f;
  {
    const TapsGrad g = make_taps_grad<double>(gx, gy, p.W0, p.H0);
    const size_t plane = (size_t)p.H0 * p.W0;
    float gix = 0.f, giy = 0.f;
    const int ce = min(p.C0, cb + cchunk);
    if (cb < ce)
      bwd_one_source<FT>((const FT*)p.src0 + (size_t)b * p.C0 * plane,
                         (const FT*)p.gout0 + (size_t)b * p.C0 * p.n + pix,
                         p.gsrc0 ? (FT*)p.gsrc0 + (size_t)b * p.C0 * plane : nullptr, g, cb, ce, plane, p.n,
                         gix, giy);
    ggx = gix * g.mx; ggy = giy * g.my;
  }
  if (p.C1 > 0 && p.gout1 != nullptr) {
    const TapsGrad g = make_taps_grad<double>(gx, gy, p.W1, p.H1);
    const size_t plane = (size_t)p.H1 * p.W1;
    float gix = 0.f, giy = 0.f;
    const int ce = min(p.C1, cb + cchunk);
    if (cb < ce)
      bwd_one_source<FT>((const FT*)p.src1 + (size_t)b * p.C1 * plane,
                         (const FT*)p.gout1 + (size_t)b * p.C1 * p.n + pix,
                         p.gsrc1 ? (FT*)p.gsrc1 + (size_t)b * p.C1 * plane : nullptr, g, cb, ce, plane, p.n,
                         gix, giy);
    ggx = __fmaf_rn(gix, g.mx, ggx); ggy = __fmaf_rn(giy, g.my, ggy);
  }
  float* gg = p.g_grid + ((size_t)b * p.n + pix) * 2;
  if (use_atomic_grid) {
    atomicAdd(gg, ggx);
    atomicAdd(gg + 1, ggy);
  } else {
    gg[0] = ggx;
    gg[1] = ggy;
  }
}

// One warp per pixel step: lane <-> rbf column (coalesced rows of P_hat / pc_score / g_pc_score),
// 8 pixel rows per CTA, fp64 accumulation of the column sums, deterministic partials.
//   partial layout: [B][nsplit][K][2] doubles
constexpr int GB_ROWS = 8;
constexpr int GB_MAXI = 4;   // ceil(128/32) column groups per lane

template <int MODE>
__global__ void __launch_bounds__(32 * GB_ROWS) grid_bwd_reduce_kernel(WarpParams p, int nsplit, double* partial) {
  extern __shared__ double sm[];          // T[2K] then reduction scratch [GB_ROWS][2K]
  double* Tsm = sm;
  double* red = sm + 2 * p.K;
  const int b = blockIdx.y, split = blockIdx.x;
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  compute_T(p, b, Tsm, threadIdx.x, blockDim.x);
  __syncthreads();
  const int per = (p.n + nsplit - 1) / nsplit;
  const int p_lo = split * per, p_hi = min(p.n, p_lo + per);
  const int ncol = (MODE == 0) ? p.F : p.K;          // streamed columns of P_hat
  double ax[GB_MAXI], ay[GB_MAXI];
#pragma unroll
  for (int i = 0; i < GB_MAXI; ++i) { ax[i] = 0.0; ay[i] = 0.0; }
  double a0x = 0, a0y = 0, a1x = 0, a1y = 0, a2x = 0, a2y = 0;   // affine columns (attention, lane 0)
  const float th = p.theta;
  for (int pix = p_lo + row; pix < p_hi; pix += GB_ROWS) {
    const float2 gg = __ldg(reinterpret_cast<const float2*>(p.g_grid + ((size_t)b * p.n + pix) * 2));
    const double dgx = (double)gg.x, dgy = (double)gg.y;
    const float* ph = p.P_hat + (size_t)pix * ncol;
    if (MODE == 0) {
      const float* s = p.score + ((size_t)b * p.n + pix) * p.F;
      float* gs = p.g_score ? p.g_score + ((size_t)b * p.n + pix) * p.F : nullptr;
#pragma unroll
      for (int i = 0; i < GB_MAXI; ++i) {
        const int k = lane + 32 * i;
        if (k < p.F) {
          const double h = (double)__ldg(ph + k);
          const double phi = h * (1.0 + (double)th * (double)__ldg(s + k));
          ax[i] = fma(phi, dgx, ax[i]);
          ay[i] = fma(phi, dgy, ay[i]);
          if (gs) {
            const double dphi = dgx * Tsm[2 * (3 + k)] + dgy * Tsm[2 * (3 + k) + 1];
            __stcs(gs + k, (float)((double)th * h * dphi));
          }
        }
      }
      if (lane == 0) {
        const double px = (double)__ldg(p.P + 2 * pix), py = (double)__ldg(p.P + 2 * pix + 1);
        a0x += dgx; a0y += dgy;
        a1x = fma(px, dgx, a1x); a1y = fma(px, dgy, a1y);
        a2x = fma(py, dgx, a2x); a2y = fma(py, dgy, a2y);
      }
    } else {
#pragma unroll
      for (int i = 0; i < GB_MAXI; ++i) {
        const int k = lane + 32 * i;
        if (k < p.K) {
          const double phi = (double)__ldg(ph + k);
          ax[i] = fma(phi, dgx, ax[i]);
          ay[i] = fma(phi, dgy, ay[i]);
        }
      }
    }
  }
  // per-row partial sums -> shared, then fixed-order sum over rows
  double* mine = red + (size_t)row * 2 * p.K;
  const int base = (MODE == 0) ? 3 : 0;
#pragma unroll
  for (int i = 0; i < GB_MAXI; ++i) {
    const int k = lane + 32 * i;
    if (k < ncol) { mine[2 * (base + k)] = ax[i]; mine[2 * (base + k) + 1] = ay[i]; }
  }
  if (MODE == 0 && lane == 0) {
    mine[0] = a0x; mine[1] = a0y; mine[2] = a1x; mine[3] = a1y; mine[4] = a2x; mine[5] = a2y;
  }
  __syncthreads();
  double* out = partial + ((size_t)b * nsplit + split) * 2 * p.K;
  for (int o = threadIdx.x; o < 2 * p.K; o += blockDim.x) {
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < GB_ROWS; ++r) acc += red[(size_t)r * 2 * p.K + o];
    out[o] = acc;
  }
}

__global__ void __launch_bounds__(256) cprime_bwd_kernel(WarpParams p, int nsplit, const double* partial) {
  extern __shared__ double dT[];   // [K][2]
  const int b = blockIdx.x;
  for (int o = threadIdx.x; o < 2 * p.K; o += blockDim.x) {
    double acc = 0.0;
    for (int s = 0; s < nsplit; ++s) acc += partial[((size_t)b * nsplit + s) * 2 * p.K + o];
    dT[o] = acc;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 2 * p.F; o += blockDim.x) {
    const int f = o >> 1, c = o & 1;
    double acc = 0.0;
    for (int k = 0; k < p.K; ++k) acc = fma((double)__ldg(p.hatC + (size_t)k * p.K + f), dT[2 * k + c], acc);
    p.g_c_prime[(size_t)b * p.F * 2 + o] = (float)acc;
  }
}

int bwd_nsplit(const tpspp_warp_cfg* cfg) {
  const int n = cfg->out_h * cfg->out_w;
  int ns = (2 * 148 + cfg->batch - 1) / (cfg->batch > 0 ? cfg->batch : 1);
  const int maxs = (n + 63) / 64;
  if (ns > maxs) ns = maxs;
  if (ns < 1) ns = 1;
  return ns;
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

template <typename FT>
static int launch_bwd_t(const tpspp_warp_cfg* cfg, WarpParams p, void* workspace, cudaStream_t st) {
  const int n = p.n;
  const int nsplit = bwd_nsplit(cfg);
  p.g_grid = reinterpret_cast<float*>(workspace);
  double* partial = reinterpret_cast<double*>((char*)workspace + align256((size_t)p.B * n * 2 * sizeof(float)));

  const size_t es = sizeof(FT);
  if (p.gsrc0) { TPSPP_CHECK_CUDA(cudaMemsetAsync(p.gsrc0, 0, (size_t)p.B * p.C0 * p.H0 * p.W0 * es, st)); count_launch(); }
  if (p.gsrc1) { TPSPP_CHECK_CUDA(cudaMemsetAsync(p.gsrc1, 0, (size_t)p.B * p.C1 * p.H1 * p.W1 * es, st)); count_launch(); }

  const int maxc = p.C0 > p.C1 ? p.C0 : p.C1;
  const int tiles = (n + 255) / 256;
  int zsplit = 1;
  const long long want = 4LL * sm_count();
  while ((long long)tiles * p.B * zsplit < want && zsplit < maxc && zsplit < 16) zsplit *= 2;
  const int cchunk = (maxc + zsplit - 1) / zsplit;
  zsplit = (maxc + cchunk - 1) / cchunk;
  const int use_atomic = zsplit > 1;
  if (use_atomic) { TPSPP_CHECK_CUDA(cudaMemsetAsync(p.g_grid, 0, (size_t)p.B * n * 2 * sizeof(float), st)); count_launch(); }
  dim3 grid(tiles, p.B, zsplit);
  const size_t smemT = (size_t)2 * p.K * sizeof(double);
  if (p.mode == TPSPP_MODE_ATTENTION)
    warp_bwd_sample_kernel<FT, 0><<<grid, 256, smemT, st>>>(p, cchunk, use_atomic);
  else
    warp_bwd_sample_kernel<FT, 1><<<grid, 256, smemT, st>>>(p, cchunk, use_atomic);
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());

  if (p.g_c_prime != nullptr || p.g_score != nullptr) {
    dim3 g2(nsplit, p.B);
    const size_t smem2 = (size_t)(1 + GB_ROWS) * 2 * p.K * sizeof(double);
    if (p.mode == TPSPP_MODE_ATTENTION)
      grid_bwd_reduce_kernel<0><<<g2, 32 * GB_ROWS, smem2, st>>>(p, nsplit, partial);
    else
      grid_bwd_reduce_kernel<1><<<g2, 32 * GB_ROWS, smem2, st>>>(p, nsplit, partial);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
    if (p.g_c_prime != nullptr) {
      cprime_bwd_kernel<<<p.B, 256, smemT, st>>>(p, nsplit, partial);
      count_launch();
      TPSPP_CHECK_CUDA(cudaGetLastError());
    }
  }
  return TPSPP_OK;
}

}  // namespace tpspp

using namespace tpspp;

extern "C" size_t tpspp_warp_workspace_bytes(const tpspp_warp_cfg* cfg) {
  if (validate_cfg(cfg) != TPSPP_OK) return 0;
  const size_t n = (size_t)cfg->out_h * cfg->out_w;
  const size_t K = (size_t)cfg->num_fiducial + 3;
  return align256((size_t)cfg->batch * n * 2 * sizeof(float)) +
         align256((size_t)cfg->batch * bwd_nsplit(cfg) * 2 * K * sizeof(double)) + 256;
}

extern "C" int tpspp_warp_bwd(const tpspp_warp_cfg* cfg, const void* src0, const void* src1,
                              const float* c_prime, const float* pc_score, const float* P_hat,
                              const float* P, const float* inv_delta_C, const void* gout0,
                              const void* gout1, void* gsrc0, void* gsrc1, float* g_c_prime,
                              float* g_pc_score, void* workspace, tpspp_stream_t stream) {
  reset_launch_count();
  int rc = validate_cfg(cfg);
  if (rc != TPSPP_OK) return rc;
  if (cfg->batch == 0) return TPSPP_OK;
  TPSPP_REQUIRE(src0 && gout0 && c_prime && P_hat && inv_delta_C && workspace,
                "tpspp_warp_bwd: null required pointer");
  TPSPP_REQUIRE(((uintptr_t)workspace & 255) == 0, "tpspp_warp_bwd: workspace must be 256-byte aligned");
  TPSPP_REQUIRE((cfg->channels1 == 0) == (src1 == nullptr), "tpspp_warp_bwd: src1 must be given exactly when channels1 > 0");
  TPSPP_REQUIRE(cfg->channels1 > 0 || (gout1 == nullptr && gsrc1 == nullptr), "tpspp_warp_bwd: gout1/gsrc1 without src1");
  TPSPP_REQUIRE(gsrc1 == nullptr || gout1 != nullptr, "tpspp_warp_bwd: gsrc1 requested without gout1");
  if (cfg->mode == TPSPP_MODE_ATTENTION)
    TPSPP_REQUIRE(pc_score && P, "tpspp_warp_bwd: attention mode needs pc_score and P");
  else
    TPSPP_REQUIRE(g_pc_score == nullptr, "tpspp_warp_bwd: classical mode has no pc_score gradient");
  if (cfg->batch == 0) return TPSPP_OK;
  WarpParams p;
  fill_params(cfg, &p);
  p.src0 = src0; p.src1 = src1; p.c_prime = c_prime; p.score = pc_score; p.P_hat = P_hat; p.P = P;
  p.hatC = inv_delta_C; p.gout0 = gout0; p.gout1 = gout1; p.gsrc0 = gsrc0; p.gsrc1 = gsrc1;
  p.g_c_prime = g_c_prime; p.g_score = g_pc_score;
  TPSPP_REQUIRE(p.B <= 65535, "batch %d exceeds the backward kernels' grid.y limit (65535)", p.B);
  if (cfg->feat_dtype == TPSPP_BF16) return launch_bwd_t<__nv_bfloat16>(cfg, p, workspace, (cudaStream_t)stream);
  return launch_bwd_t<float>(cfg, p, workspace, (cudaStream_t)stream);
}
