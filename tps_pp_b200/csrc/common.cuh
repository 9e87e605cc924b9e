// Shared device/host helpers for libtpspp (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/tpspp.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libtpspp is written for sm_100a (B200) only"
#endif

namespace tpspp {

// ---------------------------------------------------------------- host side
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void prof_begin(cudaStream_t st);      // no-op unless tpspp_launch_profile(1) was called on this thread
void reset_launch_count();
int sm_count();
// Programmatic dependent launch along one ABI call's kernel chain (tpspp_head_fwd, tpspp_stage_fwd): while pdl_scope is
// on, launch_k() sets cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel's CTAs are scheduled, run their
// prologue (barrier init, TMEM allocation) and park in pdl_wait() while the previous grid drains.  Only kernels whose every
// CTA executes pdl_wait() before it touches activations may be launched through launch_k (a grid that completes without
// having waited would break the chain's transitivity); TMEM kernels call pdl_trigger() only AFTER their own allocation, so a
// parked dependent can never hold the columns a still-unallocated CTA of the grid it waits for needs.  TPSPP_PDL=0 turns it off.
bool pdl_on();
void pdl_scope(bool on);
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_on() ? 1 : 0;
  if (cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...) != cudaSuccess && cfg.numAttrs != 0) {
    // the attribute was refused (a tool or driver without programmatic launch): the plain launch is always valid -- every chained
    // kernel's griddepcontrol instructions are no-ops in it.  A genuine launch error repeats here and is reported by the caller.
    cudaGetLastError();
    cfg.numAttrs = 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  }
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

#define TPSPP_CHECK_CUDA(expr)                                                        \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      tpspp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                       __FILE__, __LINE__);                                           \
      return TPSPP_E_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define TPSPP_REQUIRE(cond, ...)                                                      \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      tpspp::set_error(__VA_ARGS__);                                                  \
      return TPSPP_E_INVALID;                                                         \
    }                                                                                 \
  } while (0)

// Parameters shared by every warp kernel (passed by value).
struct WarpParams {
  int B, C0, H0, W0, C1, H1, W1, n, F, K;  // K = F + 3
  int mode;
  float theta;
  const void* src0;
  const void* src1;
  const float* c_prime;   // [B,F,2]
  const float* score;     // [B,n,F] or null
  const float* P_hat;     // [n,F] (attention) | [n,K] (classical)
  const float* P;         // [n,2] (attention)
  const float* hatC;      // [K,K]
  const float* grid_in;   // [B,n,2] explicit grid (sampler-only) or null
  void* out0;
  void* out1;
  float* grid_out;        // optional
  // backward
  const void* gout0;
  const void* gout1;
  void* gsrc0;
  void* gsrc1;
  double* T_ws;           // [B,K,2] fp64 T = inv_delta_C[:, :F] . C' (classical tiled forward workspace)
  float* g_grid;          // [B,n,2] workspace (accumulated with atomics)
  float* g_c_prime;
  float* g_score;
};

int validate_cfg(const tpspp_warp_cfg* cfg);
void fill_params(const tpspp_warp_cfg* cfg, WarpParams* p);
// number of pixel splits the grid-backward reduction uses for this cfg (deterministic)
int bwd_nsplit(const tpspp_warp_cfg* cfg);

// -------------------------------------------------------------- device side
#ifdef __CUDACC__

// Bilinear taps of one output pixel in one source plane.
// off = y0*W + x0, dx/dy = element offsets to the east / south neighbour (0 when that
// neighbour is outside the plane: its weight is then exactly 0 or ATen skips it),
// w = {nw, ne, sw, se}.  CT = float reproduces ATen's fp32 arithmetic bit for bit
// (ATen/native/cuda/GridSampler.cuh:23-31,52-56 + grid_sampler_2d_kernel);
// CT = double derives the same quantities without the fp32 rounding of ix/iy.
struct Taps {
  int off, dx, dy;
  float w[4];
};

template <typename CT>
__device__ __forceinline__ CT clip_coord(CT v, int size) {
  // fmax/fmin drop NaNs exactly like ::max/::min in the ATen CUDA kernel: NaN -> 0
  return fmin((CT)(size - 1), fmax(v, (CT)0));
}

template <typename CT>
__device__ __forceinline__ Taps make_taps(CT gx, CT gy, int W, int H) {
  CT ix = ((gx + (CT)1) / (CT)2) * (CT)(W - 1);
  CT iy = ((gy + (CT)1) / (CT)2) * (CT)(H - 1);
  ix = clip_coord<CT>(ix, W);
  iy = clip_coord<CT>(iy, H);
  CT x0f = floor(ix), y0f = floor(iy);
  int x0 = (int)x0f, y0 = (int)y0f;
  CT tx = ix - x0f, ty = iy - y0f;                    // (ix - ix_nw), (iy - iy_nw)
  CT ux = (x0f + (CT)1) - ix, uy = (y0f + (CT)1) - iy; // (ix_se - ix), (iy_se - iy)
  Taps t;
  t.off = y0 * W + x0;
  t.dx = (x0 + 1 < W) ? 1 : 0;
  t.dy = (y0 + 1 < H) ? W : 0;
  t.w[0] = (float)(ux * uy);
  t.w[1] = t.dx ? (float)(tx * uy) : 0.f;
  t.w[2] = t.dy ? (float)(ux * ty) : 0.f;
  t.w[3] = (t.dx && t.dy) ? (float)(tx * ty) : 0.f;
  return t;
}

// fp64-in taps with the FP64 pipe used as little as possible (make_taps<double> spends ~28 fp64
// instructions, as much as 6 control points of grid arithmetic).  Per coordinate: one
// DFMA for the pixel coordinate, the 2^52 "magic add" to round it to an integer whose low word is the
// int, two DADDs for the signed remainder, one conversion; floor/clip/weights then run in fp32 and
// integers.  Same taps as make_taps<double> up to one fp32 rounding of the fractional part (<= 6e-8).
__device__ __forceinline__ void lean_coord(double g, int size, int& i0, float& frac) {
  const double half = 0.5 * (double)(size - 1);
  const double x = fma(g, half, half);                        // ((g + 1) / 2) * (size - 1)
  const int hi = __double2hiint(x);
  const unsigned ex = ((unsigned)hi >> 20) & 0x7ffu;
  const double t = x + 4503599627370496.0;                    // 2^52: integer part lands in the low word
  const double d = x - (t - 4503599627370496.0);              // in [-0.5, 0.5]
  float df = (float)d;
  int r = __double2loint(t);
  if (df < 0.f) { r -= 1; df += 1.f; }
  const bool nan = ex == 0x7ffu && ((hi & 0xfffff) != 0 || __double2loint(x) != 0);
  if (hi < 0 || nan || r < 0) { r = 0; df = 0.f; }            // clip to 0 (ATen maps NaN there too)
  else if (ex >= 1023u + 30u || r >= size - 1) { r = size - 1; df = 0.f; }
  i0 = r; frac = df;
}

// Always-interior form: the NW tap is clamped to (size - 2) with fraction 1 at the far border, so the four
// taps sit at off, off+1, off+W, off+W+1 (no per-tap selects); needs W, H >= 2.  Same blend as make_taps for
// finite pixels: the far-border case only moves the unit weight from the NW to the NE/SW tap.
__device__ __forceinline__ int make_taps_lean(double gx, double gy, int W, int H, float* w) {
  int x0, y0; float tx, ty;
  lean_coord(gx, W, x0, tx);
  lean_coord(gy, H, y0, ty);
  if (x0 >= W - 1) { x0 = W - 2; tx = 1.f; }
  if (y0 >= H - 1) { y0 = H - 2; ty = 1.f; }
  const float ux = 1.f - tx, uy = 1.f - ty;
  w[0] = ux * uy; w[1] = tx * uy; w[2] = ux * ty; w[3] = tx * ty;
  return y0 * W + x0;
}

// Same, plus what the backward pass needs: the fractional parts and d(ix)/d(gx) including the
// border-clip rule (zero where the unclipped coordinate is <= 0 or >= size-1;
// GridSampler.cuh:58-81 clip_coordinates_set_grad).
struct TapsGrad {
  Taps t;
  float tx, ty, ux, uy;
  float mx, my;
};

template <typename CT>
__device__ __forceinline__ TapsGrad make_taps_grad(CT gx, CT gy, int W, int H) {
  CT ixu = ((gx + (CT)1) / (CT)2) * (CT)(W - 1);
  CT iyu = ((gy + (CT)1) / (CT)2) * (CT)(H - 1);
  CT ix = clip_coord<CT>(ixu, W), iy = clip_coord<CT>(iyu, H);
  CT x0f = floor(ix), y0f = floor(iy);
  int x0 = (int)x0f, y0 = (int)y0f;
  CT tx = ix - x0f, ty = iy - y0f;
  CT ux = (x0f + (CT)1) - ix, uy = (y0f + (CT)1) - iy;
  TapsGrad g;
  g.t.off = y0 * W + x0;
  g.t.dx = (x0 + 1 < W) ? 1 : 0;
  g.t.dy = (y0 + 1 < H) ? W : 0;
  g.t.w[0] = (float)(ux * uy);
  g.t.w[1] = g.t.dx ? (float)(tx * uy) : 0.f;
  g.t.w[2] = g.t.dy ? (float)(ux * ty) : 0.f;
  g.t.w[3] = (g.t.dx && g.t.dy) ? (float)(tx * ty) : 0.f;
  g.tx = (float)tx; g.ty = (float)ty; g.ux = (float)ux; g.uy = (float)uy;
  g.mx = (ixu > (CT)0 && ixu < (CT)(W - 1)) ? (float)(W - 1) * 0.5f : 0.f;
  g.my = (iyu > (CT)0 && iyu < (CT)(H - 1)) ? (float)(H - 1) * 0.5f : 0.f;
  return g;
}

// ATen accumulation order: acc=0; acc+=v_nw*nw; acc+=v_ne*ne; acc+=v_sw*sw; acc+=v_se*se,
// which nvcc contracts to this fma chain.
__device__ __forceinline__ float blend4(float vnw, float vne, float vsw, float vse, const float* w) {
  float acc = vnw * w[0];
  acc = __fmaf_rn(vne, w[1], acc);
  acc = __fmaf_rn(vsw, w[2], acc);
  acc = __fmaf_rn(vse, w[3], acc);
  return acc;
}

__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- mbarrier / bulk-copy (TMA engine, UBLKCP) wrappers ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time before it reports failure, which
// is wrong for a warp that polls several barriers round-robin)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy completing on an mbarrier (bytes % 16 == 0, both 16B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// shared -> global bulk store (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(0xffffffffu, lo, m);
  hi = __shfl_xor_sync(0xffffffffu, hi, m);
  return __hiloint2double(hi, lo);
}

// T = inv_delta_C[:, :F] . C'   (fp64, [K][2]); the three zero rows the reference appends to C'
// (tps_pp.py:489-494, tps_preprocessor.py:275-280) drop out.
__device__ __forceinline__ void compute_T(const WarpParams& p, int b, double* Tsm, int tid, int nthreads) {
  for (int o = tid; o < 2 * p.K; o += nthreads) {
    const int k = o >> 1, c = o & 1;
    const float* row = p.hatC + (size_t)k * p.K;
    const float* cp = p.c_prime + (size_t)b * p.F * 2 + c;
    double acc = 0.0;
    for (int f = 0; f < p.F; ++f) acc = fma((double)__ldg(row + f), (double)__ldg(cp + 2 * f), acc);
    Tsm[o] = acc;
  }
}

// fp64 sampling coordinate of one pixel (normalised), from T in shared memory.
template <int MODE>
__device__ __forceinline__ void pixel_grid(const WarpParams& p, const double* Tsm, int b, int pix,
                                           double& gx, double& gy) {
  if (MODE == 0) {
    const double px = (double)__ldg(p.P + 2 * pix), py = (double)__ldg(p.P + 2 * pix + 1);
    gx = Tsm[0] + px * Tsm[2] + py * Tsm[4];
    gy = Tsm[1] + px * Tsm[3] + py * Tsm[5];
    const float* ph = p.P_hat + (size_t)pix * p.F;
    const float* s = p.score + ((size_t)b * p.n + pix) * p.F;
    const double th = (double)p.theta;
    for (int k = 0; k < p.F; ++k) {
      const double phi = (double)__ldg(ph + k) * (1.0 + th * (double)__ldg(s + k));
      gx = fma(phi, Tsm[2 * (3 + k)], gx);
      gy = fma(phi, Tsm[2 * (3 + k) + 1], gy);
    }
  } else {
    gx = 0.0; gy = 0.0;
    const float* ph = p.P_hat + (size_t)pix * p.K;
    for (int k = 0; k < p.K; ++k) {
      const double phi = (double)__ldg(ph + k);
      gx = fma(phi, Tsm[2 * k], gx);
      gy = fma(phi, Tsm[2 * k + 1], gy);
    }
  }
}

#endif  // __CUDACC__
}  // namespace tpspp
