// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for the TPS_PP head, 64 output channels.
//
// Same contract as conv_ffma_kernel (head.cu): up to three concatenated NCHW inputs, nearest
// up-sampling, stride, zero padding, bias + ReLU (+ decoder skip) -- reference tps_pp.py:126-131,
// 149-169, 538-548, 560-562, 581-585.  fp32 accuracy comes from 3xTF32 error compensation: every
// fp32 operand is split into hi (the 19 bits the tf32 datapath reads) and lo = x - hi, and
//   D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi      (fp32 accumulate in TMEM)
// drops only the lo*lo term (2^-22 relative).
//
// Per CTA: a 128-pixel x 64-channel output tile; accumulator = 128 TMEM lanes x 64 fp32 columns.
// K is ordered (tap, cin) and cut into chunks of 32: a chunk lies inside one filter tap and one input
// tensor, so the im2col gather is 16 coalesced loads per thread.  All 256 threads stage chunk c+1
// (split + core-matrix layout, 16-byte k-groups: [k/4][row][4]) while the tensor core runs chunk c
// (2-deep ring, tcgen05.commit -> mbarrier releases a buffer); one thread issues the 12 MMAs of a chunk and
// brings the chunk's weight image in with one cp.async.bulk (TMA) so weights never touch registers or L1.
// Accumulators: 3 x (128 TMEM lanes x 64 fp32 columns), see the comment at the MMA issue.
#include "head.cuh"
#include "tc.cuh"

namespace tpspp {

constexpr int TC_TM = 128, TC_KC = 32, TC_N = 64;
constexpr int TC_A_LBO = TC_TM * 16 + 16;                   // bytes between 16-byte k-groups of A (padded: conflict-free stores)
constexpr int TC_A_BYTES = (TC_KC / 4) * TC_A_LBO;          // 16.1 KB per part (hi / lo)
constexpr int TC_B_BYTES = TC_N * TC_KC * 4;                // 8 KB per part
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // 48 KB
constexpr int TC_SMEM = 2 * TC_STAGE_BYTES + 64 + TC_TM * 16;   // + per-pixel geometry
constexpr int TC_TMEM_COLS = 256;   // 3 accumulators x 64 columns (power-of-two allocation)

struct ConvTcArgs {
  ConvArgs c;
  const float* wprep;   // [nchunks][hi|lo][8 k-groups][64 n][4]
};

template <int KS, bool NHWC_SRC>
__global__ void __launch_bounds__(256, 2) conv_tc_kernel(ConvTcArgs t) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * TC_STAGE_BYTES);
  uint64_t* wbars = bars + 2;                                    // weight chunk landed (TMA complete_tx)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  int4* geo = reinterpret_cast<int4*>(smem + 2 * TC_STAGE_BYTES + 64);   // per tile pixel: {image, iy0, ix0, valid}
  const ConvArgs& a = t.c;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&wbars[0], 1);
    mbar_init(&wbars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  const int HoWo = a.Ho * a.Wo;
  const long long Mtot = (long long)a.B * HoWo;
  const long long m_base = (long long)blockIdx.x * TC_TM;
  const int Ctot = a.Ctot;
  const int nchunks = (Ctot * KS * KS) / TC_KC;

  // per-pixel geometry of the tile, shared by both loader mappings
  if (tid < TC_TM) {
    const long long lm = m_base + tid;
    int4 gq = make_int4(0, 0, 0, 0);
    if (lm < Mtot) {
      const int lb = (int)(lm / HoWo);
      const int r = (int)(lm - (long long)lb * HoWo);
      const int oy = r / a.Wo, ox = r - oy * a.Wo;
      gq = make_int4(lb, oy * a.sh - a.pad, ox * a.sw - a.pad, 1);
    }
    geo[tid] = gq;
  }
  __syncthreads();
  // loader mappings.  NCHW source: thread = one pixel (lanes walk the contiguous pixel axis), k-groups g0+2i.
  // NHWC source: 8 lanes = the 8 k-groups of one pixel (128 contiguous bytes), pixels warp*16 + lane/8 + 4i.
  const int lp = NHWC_SRC ? (warp * 16 + (lane >> 3)) : (tid & (TC_TM - 1));
  const int g0 = NHWC_SRC ? (lane & 7) : (tid >> 7);
  int4 gq4[NHWC_SRC ? 4 : 1];
#pragma unroll
  for (int i = 0; i < (NHWC_SRC ? 4 : 1); ++i) gq4[i] = geo[lp + 4 * i];
  constexpr uint32_t IDESC = umma_instr_desc(TC_TM, TC_N, 2);

  int tap = 0, cin0 = 0;       // running (tap, first input channel) of the chunk: no divisions in the loop
  float4 v[4];
  // gather(ch): issue the global loads of chunk ch into v / wv (software pipeline: called right after the
  // previous chunk's registers were stored, so the loads fly during the fence / barrier / MMA issue)
  auto gather = [&](int ch) {
    // ---- chunk geometry: uniform over the CTA ----
    const int dy = tap / KS, dx = tap - dy * KS;
    int s = 0, c0 = cin0;
    if (c0 >= a.src[0].C) {
      c0 -= a.src[0].C; s = 1;
      if (c0 >= a.src[1].C) { c0 -= a.src[1].C; s = 2; }
    }
    const float* sp = s == 0 ? a.src[0].ptr : (s == 1 ? a.src[1].ptr : a.src[2].ptr);
    const int SC = s == 0 ? a.src[0].C : (s == 1 ? a.src[1].C : a.src[2].C);
    const int SH = s == 0 ? a.src[0].H : (s == 1 ? a.src[1].H : a.src[2].H);
    const int SW = s == 0 ? a.src[0].W : (s == 1 ? a.src[1].W : a.src[2].W);
    const int uh = s == 0 ? a.src[0].uh : (s == 1 ? a.src[1].uh : a.src[2].uh);
    const int uw = s == 0 ? a.src[0].uw : (s == 1 ? a.src[1].uw : a.src[2].uw);
    cin0 += TC_KC;
    if (cin0 >= Ctot) { cin0 = 0; ++tap; }
    // ---- gather: 4 x (4 channels of one pixel) per thread ----
    if (NHWC_SRC) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int4 gq = gq4[i];
        const int iy = gq.y + dy, ix = gq.z + dx;
        const bool ok = gq.w && iy >= 0 && ix >= 0 && iy < SH * uh && ix < SW * uw;
        const int sy = (uh == 2) ? (iy >> 1) : iy, sx = (uw == 2) ? (ix >> 1) : ix;
        const float* q = sp + (((size_t)gq.x * SH + (ok ? sy : 0)) * SW + (ok ? sx : 0)) * SC + c0 + g0 * 4;
        v[i] = ok ? __ldg(reinterpret_cast<const float4*>(q)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      const int4 gq = gq4[0];
      const int iy = gq.y + dy, ix = gq.z + dx;
      const bool ok = gq.w && iy >= 0 && ix >= 0 && iy < SH * uh && ix < SW * uw;
      const int sy = (uh == 2) ? (iy >> 1) : iy, sx = (uw == 2) ? (ix >> 1) : ix;
      const size_t plane = (size_t)SH * SW;
      const float* base = sp + ((size_t)gq.x * SC + c0) * plane + (size_t)(ok ? sy * SW + sx : 0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float* q = base + (size_t)((g0 + 2 * i) * 4) * plane;
        v[i] = ok ? make_float4(__ldg(q), __ldg(q + plane), __ldg(q + 2 * plane), __ldg(q + 3 * plane))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    (void)ch;
  };

  gather(0);
  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    // ---- the MMAs that read this buffer two chunks ago must have completed ----
    if (ch >= 2) mbar_wait_bounded(&bars[buf], (uint32_t)(((ch >> 1) - 1) & 1));
    unsigned char* st = smem + buf * TC_STAGE_BYTES;
    if (tid == 0) {   // weight image of this chunk (hi | lo, 16 KB): one TMA bulk copy straight into the stage
      mbar_arrive_expect_tx(&wbars[buf], 2 * TC_B_BYTES);
      bulk_g2s(st + 2 * TC_A_BYTES, t.wprep + (size_t)ch * (2 * TC_N * TC_KC), 2 * TC_B_BYTES, &wbars[buf],
               policy_evict_last());
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 hi, lo;
      split_tf32(v[i], hi, lo);
      const int off = NHWC_SRC ? (g0 * TC_A_LBO + (lp + 4 * i) * 16) : ((g0 + 2 * i) * TC_A_LBO + lp * 16);
      *reinterpret_cast<float4*>(st + off) = hi;
      *reinterpret_cast<float4*>(st + TC_A_BYTES + off) = lo;
    }
    if (ch + 1 < nchunks) gather(ch + 1);
    fence_proxy_async();     // generic-proxy stores -> visible to the tensor core's async proxy
    __syncthreads();
    if (tid == 0) {
      mbar_wait_bounded(&wbars[buf], (uint32_t)((ch >> 1) & 1));
      tc_fence_after();
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + TC_A_BYTES;
      const uint32_t b_hi = a_hi + 2 * TC_A_BYTES, b_lo = b_hi + TC_B_BYTES;
#pragma unroll
      for (int j = 0; j < TC_KC / 8; ++j) {       // one MMA = K 8 = two 16-byte k-groups
        const uint64_t dah = umma_smem_desc(a_hi + j * 2 * TC_A_LBO, TC_A_LBO, 128);
        const uint64_t dal = umma_smem_desc(a_lo + j * 2 * TC_A_LBO, TC_A_LBO, 128);
        const uint64_t dbh = umma_smem_desc(b_hi + j * 2 * (TC_N * 16), TC_N * 16, 128);
        const uint64_t dbl = umma_smem_desc(b_lo + j * 2 * (TC_N * 16), TC_N * 16, 128);
        // The tensor core truncates (does not round) when it adds into the fp32 accumulator, so the error grows
        // with the number of accumulation steps.  Three accumulators keep that at fp32 level: hi*hi of even /
        // odd chunks in D0 / D1 (half the steps at half the magnitude each) and the small lo*hi + hi*lo
        // corrections in D2, summed with round-to-nearest in the epilogue.
        umma<2>(tmem_d + 128, dal, dbh, IDESC, (ch | j) != 0 ? 1u : 0u);
        umma<2>(tmem_d + 128, dah, dbl, IDESC, 1u);
        umma<2>(tmem_d + (uint32_t)(buf * 64), dah, dbh, IDESC, (ch > 1 || j > 0) ? 1u : 0u);
      }
      umma_commit(&bars[buf]);
    }
  }
  {
    const int last = nchunks - 1;
    mbar_wait_bounded(&bars[last & 1], (uint32_t)((last >> 1) & 1));
    tc_fence_after();
  }
  // ---- epilogue: TMEM -> registers -> bias/ReLU/skip -> NCHW (lane = pixel: coalesced per channel) ----
  {
    const int wq = warp & 3, half = warp >> 2;
    float acc[32], part[32];
    const uint32_t taddr = tmem_d + ((uint32_t)(wq * 32) << 16) + (uint32_t)(half * 32);
    tmem_ld32(taddr, acc);
    tmem_ld32(taddr + 128, part);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] += part[j];
    if (nchunks > 1) {
      tmem_ld32(taddr + 64, part);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] += part[j];
    }
    const long long m = m_base + wq * 32 + lane;
    if (m < Mtot && a.out_nhwc) {      // [B,Ho,Wo,64]: this thread's 32 channels are contiguous
      const size_t o0 = (size_t)m * 64 + half * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 bs = __ldg(reinterpret_cast<const float4*>(a.bias + half * 32 + j));
        float4 r = make_float4(fmaxf(acc[j] + bs.x, 0.f), fmaxf(acc[j + 1] + bs.y, 0.f),
                               fmaxf(acc[j + 2] + bs.z, 0.f), fmaxf(acc[j + 3] + bs.w, 0.f));
        if (a.skip != nullptr) {
          const float4 sk = __ldg(reinterpret_cast<const float4*>(a.skip + o0 + j));
          r.x += sk.x; r.y += sk.y; r.z += sk.z; r.w += sk.w;
        }
        *reinterpret_cast<float4*>(a.out + o0 + j) = r;
      }
    } else if (m < Mtot) {
      const int b = (int)(m / HoWo);
      const int rem = (int)(m - (long long)b * HoWo);
      const size_t o0 = ((size_t)b * 64 + half * 32) * HoWo + rem;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float r = fmaxf(acc[j] + __ldg(a.bias + half * 32 + j), 0.f);
        const size_t o = o0 + (size_t)j * HoWo;
        if (a.skip != nullptr) r += __ldg(a.skip + o);
        a.out[o] = r;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, TC_TMEM_COLS);
}

// weight image: out[((ch*2 + part)*8 + kg)*256 + n*4 + j] with k = ch*32 + kg*4 + j = tap*Ctot + cin
struct WPrepArgs {
  WPrepLayer L[WPREP_MAX_LAYERS];
  int n;
};
__global__ void __launch_bounds__(256) wprep_kernel(WPrepArgs a) {
  const WPrepLayer L = a.L[blockIdx.y];
  const int Ktot = L.Ctot * L.taps;
  const int total = 64 * Ktot;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    const int n = i / Ktot, k = i - n * Ktot;          // coalesced writes are not needed here (tiny)
    const int tap = k / L.Ctot, cin = k - tap * L.Ctot;
    const float w = __ldg(L.w + (size_t)n * Ktot + cin * L.taps + tap);
    const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    const int ch = k >> 5, kg = (k >> 2) & 7, j = k & 3;
    float* o = L.out + (size_t)ch * 4096 + kg * 256 + n * 4 + j;
    o[0] = hi;
    o[2048] = w - hi;
  }
}

bool conv_tc_eligible(const ConvArgs& a, int KS) {
  if (a.Ctot % TC_KC) return false;
  for (int s = 0; s < 3; ++s) {
    if (a.src[s].C % TC_KC) return false;
    if (a.src[s].C > 0 && a.src[s].nhwc != a.src[0].nhwc) return false;   // one loader mapping per launch
  }
  (void)KS;
  return true;
}

size_t conv_tc_wprep_floats(int Ctot, int KS) { return (size_t)2 * 64 * Ctot * KS * KS; }

int conv_tc_prepare_weights(const WPrepLayer* layers, int nlayers, cudaStream_t st) {
  TPSPP_REQUIRE(nlayers <= WPREP_MAX_LAYERS, "too many layers for wprep");
  WPrepArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < nlayers; ++i) a.L[i] = layers[i];
  a.n = nlayers;
  dim3 grid(32, nlayers);
  wprep_kernel<<<grid, 256, 0, st>>>(a);
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());
  return TPSPP_OK;
}

int run_conv_tc(int KS, const ConvArgs& a, const float* wprep, cudaStream_t st) {
  static thread_local int attr_dev = -1;
  int dev = 0;
  TPSPP_CHECK_CUDA(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    attr_dev = dev;
  }
  ConvTcArgs t;
  t.c = a;
  t.wprep = wprep;
  const long long M = (long long)a.B * a.Ho * a.Wo;
  const unsigned grid = (unsigned)((M + TC_TM - 1) / TC_TM);
  const bool nhwc = a.src[0].nhwc != 0;
  if (KS == 1 && !nhwc) conv_tc_kernel<1, false><<<grid, 256, TC_SMEM, st>>>(t);
  else if (KS == 1) conv_tc_kernel<1, true><<<grid, 256, TC_SMEM, st>>>(t);
  else if (!nhwc) conv_tc_kernel<3, false><<<grid, 256, TC_SMEM, st>>>(t);
  else conv_tc_kernel<3, true><<<grid, 256, TC_SMEM, st>>>(t);
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());
  return TPSPP_OK;
}

}  // namespace tpspp
