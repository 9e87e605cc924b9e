// tcgen05 (5th-gen tensor core) implicit-GEMM engine of the TPS_PP head.
//
// Every dense contraction of the head is D[128 rows x N] += A[128 x K] . B[N x K]^T with M = 128 UMMA tiles and fp32
// accumulators in tensor memory:
//   * the 14 convolutions (reference tps_pp.py:126-131,149-169,538-548,560-562,581-585): up to three concatenated
//     inputs, nearest up-sampling, stride, zero padding, bias + ReLU (+ decoder skip);
//   * the linear layers that act on a contiguous feature axis, as row GEMMs: feat_linear.0/.1 (tps_pp.py:258-261,306),
//     the attention score QK^T with per-image weights p1[b] and a tanh(scale * .) epilogue (tps_pp.py:293-299),
//     DGAB's Mlp fc1 (+GELU) / fc2 (+residual) over the width axis (DGAB.py:17-23,76).
//
// fp32 accuracy comes from 3xTF32 error compensation: every fp32 operand is split into hi (the 19 bits the tf32
// datapath reads) and lo = x - hi, and  D_main += A_hi*B_hi,  D_corr += A_lo*B_hi + A_hi*B_lo  (separate TMEM
// accumulators, summed in the epilogue) drops only the lo*lo term (2^-22 relative).
//
// Kernels in this file (DESIGN.md section 4 has the measurements behind each):
//   conv_tma_kernel<KS,BF16>  convolutions whose 128-pixel tile is a rectangle of one image: activations arrive by TMA
//                             tensor-map boxes (halo tile shared by the nine taps), A operand in tensor memory (thread =
//                             pixel = TMEM lane), persistent CTAs, dedicated MMA and TMA warps.
//   conv_ts_kernel<KS,BF16>   same TS scheme with per-thread global gathers: the three smallest layers.
//   lin_tma_kernel<NT>        row-major linear layers (feat_linear.1, QK^T): 2-D swizzled TMA tiles, A kept in TMEM
//                             across the column blocks.
//   mlp_fused_kernel          DGAB's Mlp in one kernel: fc1 -> GELU -> fc2 -> + residual, hidden tensor stays in TMEM.
//   conv_tc_kernel<KS,..,NT>  the first-generation SS kernel (A operand staged in shared memory): feat_linear.0 only.
//   wprep_kernel              per-forward weight images (hi/lo split, UMMA core-matrix order).
// Warp roles are selected on a warp index obtained with __shfl_sync so the compiler can prove the branches uniform
// (a plain tid >> 5 costs two R2UR per global load inside a role); MMAs are issued by one elected lane.
#include "head.cuh"
#include "tc.cuh"

#include <cuda.h>
#include <string.h>

namespace tpspp {

constexpr int TC_TM = 128, TC_KC = 32;
constexpr int TC_A_LBO = TC_TM * 16 + 16;            // bytes between 16-byte k-groups of A (padded: conflict-free stores)
constexpr int TC_A_BYTES = (TC_KC / 4) * TC_A_LBO;   // 16.1 KB per part (hi / lo)
__host__ __device__ constexpr int tc_b_bytes(int NT) { return NT * TC_KC * 4; }             // per part
__host__ __device__ constexpr int tc_stage_bytes(int NT) { return 2 * TC_A_BYTES + 2 * tc_b_bytes(NT); }
__host__ __device__ constexpr int tc_smem_bytes(int NT) { return 2 * tc_stage_bytes(NT) + 64 + TC_TM * 16; }
constexpr int TC_PRODUCERS = 256, TC_THREADS = TC_PRODUCERS + 32;   // + the MMA-issuer warp
__host__ __device__ constexpr int tc_tmem_cols(int NT) { return NT == 64 ? 256 : 128; }     // 3*NT rounded to 2^k

struct ConvTcArgs {
  ConvArgs c;
  const float* wprep;   // [column block][nchunks][hi|lo][8 k-groups][NT n][4]
};

__device__ __forceinline__ float tc_gelu(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
// Branch-free GELU for the tensor-core Mlp epilogues: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, i.e. one
// fp32 ulp of the result and well below the 3xTF32 GEMM error), ~15 instructions.  libdevice's erff costs ~53 per
// element here because its two magnitude ranges diverge inside a warp; it was 60 % of the fused Mlp kernel.
__device__ __forceinline__ float tc_gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));   // MUFU.RCP; __frcp_rn was 22 instructions per element
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.f - p * t * __expf(-z * z);          // erf(|x| / sqrt 2)
  return 0.5f * x * (1.f + copysignf(e, x));
}
__device__ __forceinline__ float tc_act(float x, int act, float scale) {
  switch (act) {
    case CONV_ACT_RELU: return fmaxf(x, 0.f);
    case CONV_ACT_GELU: return tc_gelu(x);
    case CONV_ACT_TANH: return tanhf(x * scale);
    default: return x;
  }
}

template <int KS, bool NHWC_SRC, int NT>
__global__ void __launch_bounds__(TC_THREADS, 2) conv_tc_kernel(ConvTcArgs t) {
  constexpr int B_BYTES = tc_b_bytes(NT), STAGE = tc_stage_bytes(NT), HC = NT / 2;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE);   // "empty": MMAs that read the stage completed
  uint64_t* wbars = bars + 2;                                    // weight chunk landed (TMA complete_tx)
  uint64_t* fbars = bars + 4;                                    // "full": all 8 producer warps stored their part
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  int4* geo = reinterpret_cast<int4*>(smem + 2 * STAGE + 64);    // per tile row: {image, iy0, ix0, valid}
  const ConvArgs& a = t.c;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&wbars[0], 1);
    mbar_init(&wbars[1], 1);
    mbar_init(&fbars[0], TC_PRODUCERS / 32);
    mbar_init(&fbars[1], TC_PRODUCERS / 32);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, tc_tmem_cols(NT));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  pdl_trigger();          // this CTA holds its tensor memory: the next grid of the chain may be scheduled behind it
  pdl_wait();             // ... and nothing below runs before the previous grid has completed (no-op in a plain launch)

  const int HoWo = a.Ho * a.Wo;
  const long long Mtot = (long long)a.B * HoWo;
  const long long m_base = (long long)blockIdx.x * TC_TM;
  const int Ctot = a.Ctot;
  const int nchunks = (Ctot * KS * KS) / TC_KC;
  const int co_off = blockIdx.y * NT;
  // weight image of this column block (and of this tile's image when the weights are per image)
  const float* wimg = t.wprep + (size_t)blockIdx.y * nchunks * (2 * NT * TC_KC) +
                      (size_t)(m_base / HoWo) * (size_t)a.wimg_stride;

  // per-row geometry of the tile, shared by both loader mappings
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
  constexpr uint32_t IDESC = umma_instr_desc(TC_TM, NT, 2);
  if (warp == TC_PRODUCERS / 32) {
    // ===== MMA issuer warp: warp-uniform control flow, one elected lane issues =====
    for (int ch = 0; ch < nchunks; ++ch) {
      const int buf = ch & 1;
      const uint32_t ph = (uint32_t)((ch >> 1) & 1);
      mbar_wait_bounded(&fbars[buf], ph);      // A tile of this chunk stored (generic proxy, fenced)
      mbar_wait_bounded(&wbars[buf], ph);      // weight image landed (TMA)
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t a_hi = smem_u32(smem) + (uint32_t)(buf * STAGE), a_lo = a_hi + TC_A_BYTES;
        const uint32_t b_hi = a_hi + 2 * TC_A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
        for (int j = 0; j < TC_KC / 8; ++j) {       // one MMA = K 8 = two 16-byte k-groups
          const uint64_t dah = umma_smem_desc(a_hi + j * 2 * TC_A_LBO, TC_A_LBO, 128);
          const uint64_t dal = umma_smem_desc(a_lo + j * 2 * TC_A_LBO, TC_A_LBO, 128);
          const uint64_t dbh = umma_smem_desc(b_hi + j * 2 * (NT * 16), NT * 16, 128);
          const uint64_t dbl = umma_smem_desc(b_lo + j * 2 * (NT * 16), NT * 16, 128);
          // The tensor core truncates (does not round) when it adds into the fp32 accumulator, so the error
          // grows with the number of accumulation steps.  Three accumulators keep that at fp32 level: hi*hi
          // of even / odd chunks in D0 / D1 (half the steps at half the magnitude each) and the small
          // lo*hi + hi*lo corrections in D2, summed with round-to-nearest in the epilogue.
          umma<2>(tmem_d + 2 * NT, dal, dbh, IDESC, (ch | j) != 0 ? 1u : 0u);
          umma<2>(tmem_d + 2 * NT, dah, dbl, IDESC, 1u);
          umma<2>(tmem_d + (uint32_t)(buf * NT), dah, dbh, IDESC, (ch > 1 || j > 0) ? 1u : 0u);
        }
        umma_commit(&bars[buf]);                    // -> "empty" when these MMAs have read the stage
      }
      __syncwarp();
    }
  } else {
    // ===== producer warps =====
    // loader mappings.  NCHW source: thread = one pixel (lanes walk the contiguous pixel axis), k-groups g0+2i.
    // NHWC source: 8 lanes = the 8 k-groups of one pixel (128 contiguous bytes), pixels warp*16 + lane/8 + 4i.
    const int lp = NHWC_SRC ? (warp * 16 + (lane >> 3)) : (tid & (TC_TM - 1));
    const int g0 = NHWC_SRC ? (lane & 7) : (tid >> 7);
    int4 gq4[NHWC_SRC ? 4 : 1];
#pragma unroll
    for (int i = 0; i < (NHWC_SRC ? 4 : 1); ++i) gq4[i] = geo[lp + 4 * i];

    int tap = 0, cin0 = 0;       // running (tap, first input channel) of the chunk: no divisions in the loop
    float4 va[4], vb[4];
    // gather(v): issue the global loads of the next chunk into v.  Two register sets are kept in flight
    // (chunks c+1 and c+2), so a load has two full iterations to land before its data is split and stored.
    auto gather = [&](float4 (&v)[4]) {
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
    };

    auto produce = [&](int ch, float4 (&v)[4]) {
      const int buf = ch & 1;
      // the MMAs that read this stage two chunks ago must have completed
      if (ch >= 2) mbar_wait_bounded(&bars[buf], (uint32_t)(((ch >> 1) - 1) & 1));
      unsigned char* st = smem + buf * STAGE;
      if (tid == 0) {   // weight image of this chunk (hi | lo): one TMA bulk copy straight into the stage
        mbar_arrive_expect_tx(&wbars[buf], 2 * B_BYTES);
        bulk_g2s(st + 2 * TC_A_BYTES, wimg + (size_t)ch * (2 * NT * TC_KC), 2 * B_BYTES, &wbars[buf], policy_evict_last());
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 hi, lo;
        split_tf32(v[i], hi, lo);
        const int off = NHWC_SRC ? (g0 * TC_A_LBO + (lp + 4 * i) * 16) : ((g0 + 2 * i) * TC_A_LBO + lp * 16);
        *reinterpret_cast<float4*>(st + off) = hi;
        *reinterpret_cast<float4*>(st + TC_A_BYTES + off) = lo;
      }
      if (ch + 2 < nchunks) gather(v);   // refill this register set with chunk ch+2
      fence_proxy_async();     // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&fbars[buf]);
    };
    gather(va);
    if (nchunks > 1) gather(vb);
    for (int ch = 0; ch < nchunks; ch += 2) {
      produce(ch, va);
      if (ch + 1 < nchunks) produce(ch + 1, vb);
    }
    const int last = nchunks - 1;
    mbar_wait_bounded(&bars[last & 1], (uint32_t)((last >> 1) & 1));   // the last commit covers every MMA
    tc_fence_after();
  }
  // ---- epilogue (producer warps): TMEM -> registers -> bias / activation / skip -> global.  Warp w reads
  //      TMEM lanes 32*(w%4).. (its rows) and columns (w/4)*NT/2.., 16 columns per pass ----
  if (warp < TC_PRODUCERS / 32) {
    const int wq = warp & 3, half = warp >> 2;
    const long long m = m_base + wq * 32 + lane;
    const int b = (m < Mtot) ? (int)(m / HoWo) : 0;
    const int rem = (m < Mtot) ? (int)(m - (long long)b * HoWo) : 0;
#pragma unroll 1
    for (int pass = 0; pass < HC / 16; ++pass) {
      float acc[16], part[16];
      const uint32_t taddr = tmem_d + ((uint32_t)(wq * 32) << 16) + (uint32_t)(half * HC + pass * 16);
      tmem_ld_cols<16>(taddr, acc);
      tmem_ld_cols<16>(taddr + 2 * NT, part);
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] += part[j];
      if (nchunks > 1) {
        tmem_ld_cols<16>(taddr + NT, part);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += part[j];
      }
      const int cb = co_off + half * HC + pass * 16;     // first output channel of this pass
      if (m < Mtot && a.out_nhwc) {      // [rows, Cout]: 16 contiguous channels
        const size_t o0 = (size_t)m * a.Cout + cb;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.bias != nullptr) bs = __ldg(reinterpret_cast<const float4*>(a.bias + cb + j));
          float4 r = make_float4(tc_act(acc[j] + bs.x, a.act, a.act_scale), tc_act(acc[j + 1] + bs.y, a.act, a.act_scale),
                                 tc_act(acc[j + 2] + bs.z, a.act, a.act_scale), tc_act(acc[j + 3] + bs.w, a.act, a.act_scale));
          if (a.skip != nullptr) {
            const float4 sk = __ldg(reinterpret_cast<const float4*>(a.skip + o0 + j));
            r.x += sk.x; r.y += sk.y; r.z += sk.z; r.w += sk.w;
          }
          *reinterpret_cast<float4*>(a.out + o0 + j) = r;
        }
      } else if (m < Mtot) {             // [B, Cout, Ho, Wo]: lane = pixel, coalesced per channel
        const size_t o0 = ((size_t)b * a.Cout + cb) * HoWo + rem;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float r = tc_act(acc[j] + (a.bias != nullptr ? __ldg(a.bias + cb + j) : 0.f), a.act, a.act_scale);
          const size_t o = o0 + (size_t)j * HoWo;
          if (a.skip != nullptr) r += __ldg(a.skip + o);
          a.out[o] = r;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, tc_tmem_cols(NT));
}


// ------------------------------------------------------------------------------------------------
// Operand modes of the convolution kernels.
//   CM_TF32X3  3xTF32: D_main += A_hi B_hi, D_corr += A_lo B_hi + A_hi B_lo, all kind::tf32 (12 MMAs of K = 8 per chunk).
//   CM_BF16    single-pass bf16 operands (the opt-in reduced-precision mode).
//   CM_MIX     fp32-level accuracy at 2/3 of the tensor time: the main term stays kind::tf32 (4 MMAs of K = 8), the two
//              correction terms -- 2^-11 of the result, so 8 significant bits are enough -- run as kind::f16 with bf16
//              operands (2 + 2 MMAs of K = 16, each half the cycles of a tf32 MMA): A_lo(bf16) B(bf16) + A(bf16) B_lo(bf16).
//              bf16 keeps the fp32 exponent range, so unlike an fp16 split there is no overflow/underflow case.
//              Error per product: |a_lo| <= 2^-10 |a| and bf16 round-to-nearest is 2^-9 relative => <= 2^-18 |a w| per
//              correction term, unbiased (the all-tf32 form truncates a_lo/w_lo to 11 bits inside the tensor core: 2^-20,
//              biased); the dropped lo*lo term is 2^-21 as before.
//   A stage in tensor memory (64 columns): hi tf32 [0,32) | lo bf16x2 [32,48) | a bf16x2 [48,64)
//   weight image of a chunk (16 KB):       hi tf32 8 KB    | w bf16 4 KB      | lo bf16 4 KB   (k-group-major, see wprep_kernel)
// ------------------------------------------------------------------------------------------------
// byte offsets of the bf16 images inside a chunk's weight image (NT output rows: hi tf32 NT*128 B | w bf16 NT*64 B | lo bf16 NT*64 B)
__host__ __device__ constexpr uint32_t mix_b_bf(int NT) { return (uint32_t)NT * 128u; }
__host__ __device__ constexpr uint32_t mix_b_lo(int NT) { return (uint32_t)NT * 192u; }
constexpr uint32_t MIX_B_BF = mix_b_bf(64), MIX_B_LO = mix_b_lo(64);

// one elected lane: the 8 MMAs of a chunk (a_base = first TMEM column of the A stage, b_base = shared address of the image)
template <int NT>
__device__ __forceinline__ void mix_mma_chunk(uint32_t tmem_d, uint32_t a_base, uint32_t b_base, uint32_t idesc_tf, uint32_t idesc_bf,
                                              bool first) {
#pragma unroll
  for (int j = 0; j < TC_KC / 16; ++j) {        // corrections: K = 16 bf16 per MMA = two 16-byte k-groups
    const uint64_t dbw = umma_smem_desc(b_base + mix_b_bf(NT) + j * 2 * (NT * 16), NT * 16, 128);
    const uint64_t dbl = umma_smem_desc(b_base + mix_b_lo(NT) + j * 2 * (NT * 16), NT * 16, 128);
    umma_ts_f16(tmem_d + 64, a_base + 32 + j * 8, dbw, idesc_bf, (first && j == 0) ? 0u : 1u);
    umma_ts_f16(tmem_d + 64, a_base + 48 + j * 8, dbl, idesc_bf, 1u);
  }
#pragma unroll
  for (int j = 0; j < TC_KC / 8; ++j) {         // main term: K = 8 tf32 per MMA
    const uint64_t dbh = umma_smem_desc(b_base + j * 2 * (NT * 16), NT * 16, 128);
    umma_ts_tf32(tmem_d, a_base + j * 8, dbh, idesc_tf, (first && j == 0) ? 0u : 1u);
  }
}
// The MMAs of one chunk in the CM_MIX form, the commit that releases the stage, and -- hidden behind them -- the probes of
// the NEXT chunk's two barriers, as ONE asm block.  A barrier check costs the issuing thread ~200-250 cycles of latency even
// when the phase completed long ago (clock64 timeline, profiles/r02_conv_experiments.md); issued as separate statements each
// check is followed by the selp that consumes its predicate and the in-order thread stalls there, so two checks per chunk
// were ~500 of the MMA warp's ~900 cycles per chunk -- and that loop is the critical path of the 3x3 convolutions.  Here both
// test_wait go out back to back, the eight UTCHMMA and the UTCBAR follow without depending on them, and the predicates are
// read only after the issue (which blocks for the MMAs' ~270 cycles anyway).
template <int NT>
__device__ __forceinline__ void mix_mma_chunk_probed(uint32_t d, uint32_t a_base, uint64_t bdesc0, uint32_t idesc_tf, uint32_t idesc_bf,
                                                     uint32_t accumulate, uint32_t commit_bar, uint32_t next_a_bar, uint32_t next_w_bar,
                                                     uint32_t next_parity, uint32_t& ok_a, uint32_t& ok_w) {
  // descriptor address field is in 16-byte units: one MMA's two k-groups are 2 * NT * 16 bytes apart
  constexpr int KS2 = 2 * NT, BW0 = 8 * NT, BL0 = 12 * NT;
  asm volatile(
      "{\n"
      ".reg .pred pa, pw, pf, pt;\n"
      ".reg .b64 dsc;\n"
      ".reg .b32 ta, tc;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 pa, [%2], %4;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 pw, [%3], %4;\n"
      "setp.ne.b32 pf, %9, 0;\n"
      "setp.eq.u32 pt, %4, %4;\n"
      "add.u32 tc, %5, 64;\n"
      // corrections (kind::f16, bf16 operands, K = 16 each): lo(A) x bf16(W) and bf16(A) x lo(W), k-halves 0 and 1
      "add.u32 ta, %6, 32;\n add.u64 dsc, %7, %12;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [tc], [ta], dsc, %11, pf;\n"
      "add.u32 ta, %6, 48;\n add.u64 dsc, %7, %13;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [tc], [ta], dsc, %11, pt;\n"
      "add.u32 ta, %6, 40;\n add.u64 dsc, %7, %14;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [tc], [ta], dsc, %11, pt;\n"
      "add.u32 ta, %6, 56;\n add.u64 dsc, %7, %15;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [tc], [ta], dsc, %11, pt;\n"
      // main term (kind::tf32, K = 8 each)
      "tcgen05.mma.cta_group::1.kind::tf32 [%5], [%6], %7, %10, pf;\n"
      "add.u32 ta, %6, 8;\n add.u64 dsc, %7, %16;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%5], [ta], dsc, %10, pt;\n"
      "add.u32 ta, %6, 16;\n add.u64 dsc, %7, %17;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%5], [ta], dsc, %10, pt;\n"
      "add.u32 ta, %6, 24;\n add.u64 dsc, %7, %18;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%5], [ta], dsc, %10, pt;\n"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n"
      "selp.u32 %0, 1, 0, pa;\n"
      "selp.u32 %1, 1, 0, pw;\n"
      "}\n"
      : "=r"(ok_a), "=r"(ok_w)
      : "r"(next_a_bar), "r"(next_w_bar), "r"(next_parity), "r"(d), "r"(a_base), "l"(bdesc0), "r"(commit_bar), "r"(accumulate),
        "r"(idesc_tf), "r"(idesc_bf), "n"(BW0), "n"(BL0), "n"(BW0 + KS2), "n"(BL0 + KS2), "n"(KS2), "n"(2 * KS2), "n"(3 * KS2)
      : "memory");
}
// producer thread: its 16 channels (half kh of the chunk) -> the three parts of the A stage
__device__ __forceinline__ void mix_split(const float (&v)[16], float (&hi)[16], uint32_t (&lo2)[8], uint32_t (&a2)[8]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) hi[i] = __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * i] - hi[2 * i], v[2 * i + 1] - hi[2 * i + 1]);   // .x (low half) = even channel
    const __nv_bfloat162 x2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    lo2[i] = *reinterpret_cast<const uint32_t*>(&l2);
    a2[i] = *reinterpret_cast<const uint32_t*>(&x2);
  }
}
__device__ __forceinline__ void mix_store_a(uint32_t stage_addr, int kh, const float (&v)[16]) {
  float hi[16];
  uint32_t lo2[8], a2[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) hi[i] = __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * i] - hi[2 * i], v[2 * i + 1] - hi[2 * i + 1]);   // .x (low half) = even channel
    const __nv_bfloat162 x2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    lo2[i] = *reinterpret_cast<const uint32_t*>(&l2);
    a2[i] = *reinterpret_cast<const uint32_t*>(&x2);
  }
  tmem_st16(stage_addr + (uint32_t)(kh * 16), hi);
  tmem_st8(stage_addr + (uint32_t)(32 + kh * 8), lo2);
  tmem_st8(stage_addr + (uint32_t)(48 + kh * 8), a2);
}

// ------------------------------------------------------------------------------------------------
// TS variant for the convolutions (NCHW sources, 64 output channels): the A operand never touches
// shared memory.  With N = 64 an SS-mode UTCHMMA re-reads its 128x8 A slice from shared memory on
// every issue (6 KB per 32-cycle MMA = 192 B/clk > the 128 B/clk of the shared-memory pipe, three
// times per k-step for 3xTF32), which capped the SS kernel at ~30 % tensor-pipe activity.  Here each
// producer thread owns one output pixel = one TMEM lane: 16 coalesced channel loads -> hi/lo split in
// registers -> two tcgen05.st (16 columns each) into a 2-stage A ring in tensor memory; the MMA warp
// issues tcgen05.mma with A from TMEM and B (the TMA-loaded weight image) from shared memory.
// TMEM columns (256 per CTA, 2 CTAs/SM): D_main [0,64) | D_corr [64,128) | A stage s: hi 128+64s, lo +32.
// ------------------------------------------------------------------------------------------------
template <int PLANE>
__device__ __forceinline__ void ts_load16(const float* q, float (&v)[16], bool ok) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = ok ? __ldg(q + i * PLANE) : 0.f;
}

constexpr int TS_B_BYTES = tc_b_bytes(64);                 // per part
constexpr int TS_STAGE = 2 * TS_B_BYTES;                   // weight image hi | lo
constexpr int TS_SMEM = 2 * TS_STAGE + 64 + TC_TM * 16 + 256;

// BF16 = true: single-pass bf16 operands (kind::f16, fp32 accumulate) instead of 3xTF32 -- the opt-in reduced
// precision mode (TPSPP_HEAD_BF16): A packs two channels per TMEM column, the weight image is bf16.
template <int KS, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 2) conv_ts_kernel(ConvTcArgs t) {
  constexpr bool BF16 = MODE == CM_BF16, MIX = MODE == CM_MIX;
  constexpr int NT = 64;
  constexpr int W_BYTES = BF16 ? NT * TC_KC * 2 : 2 * TS_B_BYTES;      // weight image bytes per chunk
  constexpr int A_COL0 = BF16 ? 64 : 128;                              // first TMEM column of the A ring
  constexpr int A_STAGE = BF16 ? 16 : 64;                              // TMEM columns per A stage
  constexpr int TMEM_COLS = BF16 ? 128 : 256;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * TS_STAGE);   // "empty"
  uint64_t* wbars = bars + 2;
  uint64_t* fbars = bars + 4;                                          // "full"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  int4* geo = reinterpret_cast<int4*>(smem + 2 * TS_STAGE + 64);
  float* bias_s = reinterpret_cast<float*>(smem + 2 * TS_STAGE + 64 + TC_TM * 16);   // [64]
  const ConvArgs& a = t.c;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: role branches stay on the uniform datapath

  if (tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    mbar_init(&wbars[0], 1); mbar_init(&wbars[1], 1);
    mbar_init(&fbars[0], TC_PRODUCERS / 32); mbar_init(&fbars[1], TC_PRODUCERS / 32);
    fence_barrier_init();
  }
  if (tid < NT) bias_s[tid] = (a.bias != nullptr && tid < a.Cout) ? __ldg(a.bias + tid) : 0.f;
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  pdl_trigger();          // this CTA holds its tensor memory: the next grid of the chain may be scheduled behind it
  pdl_wait();             // ... and nothing below runs before the previous grid has completed (no-op in a plain launch)

  const int HoWo = a.Ho * a.Wo;
  const long long Mtot = (long long)a.B * HoWo;
  const long long m_base = (long long)blockIdx.x * TC_TM;
  const int Ctot = a.Ctot;
  const int nchunks = (Ctot * KS * KS) / TC_KC;
  const unsigned char* wimg = reinterpret_cast<const unsigned char*>(t.wprep);

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
  constexpr uint32_t IDESC = umma_instr_desc(TC_TM, NT, BF16 ? 1 : 2);
  constexpr uint32_t IDESC_BF = umma_instr_desc(TC_TM, NT, 1);
  if (warp == TC_PRODUCERS / 32) {
    // ===== MMA issuer warp =====
    for (int ch = 0; ch < nchunks; ++ch) {
      const int buf = ch & 1;
      const uint32_t ph = (uint32_t)((ch >> 1) & 1);
      mbar_wait_bounded(&fbars[buf], ph);      // A of this chunk is in tensor memory
      mbar_wait_bounded(&wbars[buf], ph);      // weight image landed (TMA)
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t b_hi = smem_u32(smem) + (uint32_t)(buf * TS_STAGE), b_lo = b_hi + TS_B_BYTES;
        const uint32_t a_hi = tmem_d + A_COL0 + (uint32_t)(buf * A_STAGE), a_lo = a_hi + 32;
        if (BF16) {
#pragma unroll
          for (int j = 0; j < TC_KC / 16; ++j) {     // one MMA = K 16 = two 16-byte k-groups of 8 bf16
            const uint64_t db = umma_smem_desc(b_hi + j * 2 * (NT * 16), NT * 16, 128);
            umma_ts_f16(tmem_d, a_hi + j * 8, db, IDESC, (ch | j) != 0 ? 1u : 0u);
          }
        } else if (MIX) {
          mix_mma_chunk<NT>(tmem_d, a_hi, b_hi, IDESC, IDESC_BF, ch == 0);
        } else {
#pragma unroll
          for (int j = 0; j < TC_KC / 8; ++j) {
            const uint64_t dbh = umma_smem_desc(b_hi + j * 2 * (NT * 16), NT * 16, 128);
            const uint64_t dbl = umma_smem_desc(b_lo + j * 2 * (NT * 16), NT * 16, 128);
            umma_ts_tf32(tmem_d + 64, a_lo + j * 8, dbh, IDESC, (ch | j) != 0 ? 1u : 0u);   // corrections
            umma_ts_tf32(tmem_d + 64, a_hi + j * 8, dbl, IDESC, 1u);
            umma_ts_tf32(tmem_d, a_hi + j * 8, dbh, IDESC, (ch | j) != 0 ? 1u : 0u);        // main
          }
        }
        umma_commit(&bars[buf]);
      }
      __syncwarp();
    }
  } else {
    // ===== producer warps: thread = output pixel = TMEM lane (warp w owns lanes 32*(w%4)..), the two
    //       warps that share a lane quarter take the chunk's channels [0,16) and [16,32) =====
    const int row = (warp & 3) * 32 + lane;
    const int kh = warp >> 2;
    const int4 gq = geo[row];
    const uint32_t lane_addr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    // running (tap, source, channel) state of the next chunk to gather -- uniform over the CTA, advanced
    // incrementally so the loop has no divisions and touches the source descriptors only when the source changes
    int tap = 0, cin0 = 0, c_in = 0, s_idx = 0;
    const float* t_base = nullptr;   // this thread's first channel of the current source, pixel (0,0) of its image
    int s_C = 0, s_W = 0, s_uh = 1, s_uw = 1, lim_y = 0, lim_x = 0;
    size_t s_plane = 0;
    auto load_src = [&](int si) {
      const float* sp = si == 0 ? a.src[0].ptr : (si == 1 ? a.src[1].ptr : a.src[2].ptr);
      s_C = si == 0 ? a.src[0].C : (si == 1 ? a.src[1].C : a.src[2].C);
      const int SH = si == 0 ? a.src[0].H : (si == 1 ? a.src[1].H : a.src[2].H);
      s_W = si == 0 ? a.src[0].W : (si == 1 ? a.src[1].W : a.src[2].W);
      s_uh = si == 0 ? a.src[0].uh : (si == 1 ? a.src[1].uh : a.src[2].uh);
      s_uw = si == 0 ? a.src[0].uw : (si == 1 ? a.src[1].uw : a.src[2].uw);
      s_plane = (size_t)SH * s_W;
      lim_y = SH * s_uh; lim_x = s_W * s_uw;
      t_base = sp + ((size_t)gq.x * s_C + kh * 16) * s_plane;
    };
    load_src(0);
    float va[16], vb[16];
    auto gather = [&](float (&v)[16]) {
      const int dy = tap / KS, dx = tap - dy * KS;
      const int iy = gq.y + dy, ix = gq.z + dx;
      bool ok = gq.w && iy >= 0 && ix >= 0 && iy < lim_y && ix < lim_x;
      // zero insertion (the data gradient of a strided convolution is a stride-1 convolution over the gradient with zeros
      // between its pixels): only even positions of an up-sampled axis carry data
      if (a.zi && (((s_uh == 2) && (iy & 1)) || ((s_uw == 2) && (ix & 1)))) ok = false;
      const int sy = (s_uh == 2) ? (iy >> 1) : iy, sx = (s_uw == 2) ? (ix >> 1) : ix;
      const float* q = t_base + (size_t)c_in * s_plane + (ok ? sy * s_W + sx : 0);
      // the 16 channel planes are a compile-time stride apart for the plane sizes of this network: one LDG with an
      // immediate offset per channel instead of 64-bit pointer arithmetic per load
      switch (s_plane) {
        case 4096: ts_load16<4096>(q, v, ok); break;
        case 1024: ts_load16<1024>(q, v, ok); break;
        case 256: ts_load16<256>(q, v, ok); break;
        case 64: ts_load16<64>(q, v, ok); break;
        case 32: ts_load16<32>(q, v, ok); break;
        default:
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v[i] = ok ? __ldg(q) : 0.f;
            q += s_plane;
          }
      }
      c_in += TC_KC; cin0 += TC_KC;
      if (c_in >= s_C) {                 // next chunk starts in another source tensor (or the next tap)
        c_in = 0;
        if (cin0 >= Ctot) { cin0 = 0; ++tap; s_idx = 0; } else { ++s_idx; }
        load_src(s_idx);
      }
    };
    auto produce = [&](int ch, float (&v)[16]) {
      const int buf = ch & 1;
      if (ch >= 2) {
        mbar_wait_bounded(&bars[buf], (uint32_t)(((ch >> 1) - 1) & 1));   // MMAs of chunk ch-2 have read the stage
        tc_fence_after();
      }
      if (tid == 0) {
        unsigned char* st = smem + buf * TS_STAGE;
        mbar_arrive_expect_tx(&wbars[buf], W_BYTES);
        bulk_g2s(st, wimg + (size_t)ch * W_BYTES, W_BYTES, &wbars[buf], policy_evict_last());
      }
      if (BF16) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);   // .x (low half) = even channel
          pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        tmem_st8(lane_addr + (uint32_t)(A_COL0 + buf * A_STAGE + kh * 8), pk);
      } else if (MIX) {
        mix_store_a(lane_addr + (uint32_t)(A_COL0 + buf * A_STAGE), kh, v);
      } else {
        float hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          hi[i] = __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u);
          lo[i] = v[i] - hi[i];
        }
        const uint32_t col = (uint32_t)(A_COL0 + buf * A_STAGE + kh * 16);
        tmem_st16(lane_addr + col, hi);
        tmem_st16(lane_addr + col + 32, lo);
      }
      if (ch + 2 < nchunks) gather(v);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&fbars[buf]);
    };
    gather(va);
    if (nchunks > 1) gather(vb);
    for (int ch = 0; ch < nchunks; ch += 2) {
      produce(ch, va);
      if (ch + 1 < nchunks) produce(ch + 1, vb);
    }
    const int last = nchunks - 1;
    mbar_wait_bounded(&bars[last & 1], (uint32_t)((last >> 1) & 1));
    tc_fence_after();

    // ---- epilogue ----
    const int wq = warp & 3, half = warp >> 2;
    const long long m = m_base + wq * 32 + lane;
    const int b = (m < Mtot) ? (int)(m / HoWo) : 0;
    const int rem = (m < Mtot) ? (int)(m - (long long)b * HoWo) : 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      float acc[16], part[16];
      const uint32_t taddr = lane_addr + (uint32_t)(half * 32 + pass * 16);
      tmem_ld_cols<16>(taddr, acc);
      if (!BF16) {
        tmem_ld_cols<16>(taddr + 64, part);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += part[j];
      }
      const int cb = half * 32 + pass * 16;
      if (m < Mtot && a.out_nhwc) {
        const size_t o0 = (size_t)m * a.Cout + cb;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.bias != nullptr) bs = __ldg(reinterpret_cast<const float4*>(a.bias + cb + j));
          float4 r = make_float4(tc_act(acc[j] + bs.x, a.act, a.act_scale), tc_act(acc[j + 1] + bs.y, a.act, a.act_scale),
                                 tc_act(acc[j + 2] + bs.z, a.act, a.act_scale), tc_act(acc[j + 3] + bs.w, a.act, a.act_scale));
          if (a.skip != nullptr) {
            const float4 sk = __ldg(reinterpret_cast<const float4*>(a.skip + o0 + j));
            r.x += sk.x; r.y += sk.y; r.z += sk.z; r.w += sk.w;
          }
          *reinterpret_cast<float4*>(a.out + o0 + j) = r;
        }
      } else if (m < Mtot) {             // [B, 64, Ho, Wo]: lane = pixel, coalesced per channel
        const float4* b4 = reinterpret_cast<const float4*>(bias_s + cb);       // warp-uniform: broadcast loads
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bs = b4[j >> 2];
          acc[j] += bs.x; acc[j + 1] += bs.y; acc[j + 2] += bs.z; acc[j + 3] += bs.w;
        }
        const size_t o0 = ((size_t)b * (a.out_cstride ? a.out_cstride : a.Cout) + cb) * HoWo + rem;
        if (cb >= a.Cout) continue;
        if (a.skip != nullptr && a.skip_pre) {
          const float* sk = a.skip + o0;
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += __ldg(sk + (size_t)j * HoWo);
        }
        if (a.act == CONV_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = fmaxf(acc[j], 0.f);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = tc_act(acc[j], a.act, a.act_scale);
        }
        if (a.skip != nullptr && !a.skip_pre) {
          const float* sk = a.skip + o0;
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += __ldg(sk + (size_t)j * HoWo);
        }
        float* po = a.out + o0;
#pragma unroll
        for (int j = 0; j < 16; ++j) po[(size_t)j * HoWo] = acc[j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// TMA-staged form of the TS convolution (stride-1 convolutions whose 128-pixel tile is a TH x TW rectangle of one
// image).  The per-thread global gathers of conv_ts_kernel cost ~2000 warp instructions per 32-channel chunk (address
// arithmetic, padding predicates, the chunk state machine) against ~400 that do useful work, so the tensor pipe idled.
// Here the activations arrive by one cp.async.bulk.tensor (4-D tensor map over [B,C,H,W], box = 32 channels x the
// tile's halo rectangle, out-of-bounds = the convolution's zero padding) per 32-channel group, and the K loop runs
// (channel group, tap): the nine taps of a 3x3 filter read the same shared-memory halo tile at shifted offsets, so
// global/L2 traffic per tile drops 9x and a producer thread's chunk is 16 LDS + the hi/lo split + two tcgen05.st.
// Nearest-upsampled sources use a box in the low-resolution tensor and index it with (y >> 1, x >> 1).
// ------------------------------------------------------------------------------------------------
struct ConvTmaArgs {
  ConvTcArgs t;
  CUtensorMap tmap[3];
  int TW, TH, BW, BH;      // tile and box (halo) extent in pixels; CHS = BW * BH floats per channel
  int TX, TPI;             // tiles per image row, tiles per image
  int S;                   // convolution stride (1, or 2 for 3x3: one box per filter row, rows strided by the tensor map)
  int XH;                  // x halo of the box in elements (3x3: 4 floats or 8 bf16 = 16 bytes; 1x1: 0)
};
constexpr int TM_XH = 4;   // x halo of a 3x3 box: the innermost TMA coordinate must be 16-byte aligned (probed: x = -1 traps)
constexpr int TM_TILE_MAX = 32 * (4 * 72) * 4;                            // 3x3 at TW = 64: 4 rows x 72 columns
constexpr int TM_TILE_BF = 17408;      // bf16 NHWC source: 66 x 4 pixels x 32 ch x 2 B = 16.5 KB (1 KB aligned); four buffers in the room of two fp32 ones
__host__ __device__ constexpr int tm_tile_bytes(int KS) { return KS == 3 ? ((TM_TILE_MAX + 1023) / 1024) * 1024 : 16384; }
__host__ __device__ constexpr int tm_tile_bufs(int KS) { return KS == 3 ? 2 : 4; }
// (3x3: + 4 KB so that the bf16 kernel's three 12 KB weight stages fit where the other modes keep two 16 KB stages)
__host__ __device__ constexpr int tm_smem_bytes(int KS) { return tm_tile_bufs(KS) * tm_tile_bytes(KS) + 2 * TS_STAGE + (KS == 3 ? 4096 : 0) + 512 + 1024; }
constexpr int TM_THREADS = TC_THREADS + 32;                                // 8 producer warps, MMA warp, TMA warp

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
      : "memory");
}

// BF16 = true: single-pass bf16 operands (kind::f16, fp32 accumulate) -- the opt-in TPSPP_HEAD_BF16 mode: A packs two
// channels per TMEM column, the weight image is bf16, no correction accumulator.
template <int KS, int MODE, int NT = 64>
__global__ void __launch_bounds__(TM_THREADS, 2) conv_tma_kernel(const __grid_constant__ ConvTmaArgs g) {
  constexpr bool BF16 = MODE == CM_BF16, MIX = MODE == CM_MIX;
  constexpr int T = KS * KS;
  static_assert(NT == 64 || NT == 32, "64 or 32 output channels per tile");
  const int XH = g.XH;
  // bf16-STORED activations (bf16 mode of the head, this kernel's <3, CM_BF16> form only): 2-byte tiles -- three tile buffers
  // in the room of two, half the L2 / HBM bytes per tile (with fp32 tiles the bf16 kernel was bound by its tile loads)
  const bool SBF = (BF16 && KS == 3) && g.t.c.src[0].bf16 != 0;
  const int ntb_rt = SBF ? 4 : tm_tile_bufs(KS);
  const int tile_stride = SBF ? TM_TILE_BF : tm_tile_bytes(KS);
  // bf16 3x3: a chunk is one FILTER ROW (3 taps x 32 channels, K = 96) -- a third of the stage hand-offs, which is what bounds
  // this kernel once the operands need a single MMA pass (profiles/r02_conv_experiments.md: 780 cycles per hand-off);
  // the A stage holds 3 x 16 packed columns, the weight stage the three taps' 4 KB images
  constexpr int CT = (BF16 && KS == 3) ? 3 : 1;                  // taps per chunk
  constexpr int W_TAP = BF16 ? NT * TC_KC * 2 : 2 * tc_b_bytes(NT);        // weight image bytes per (tap, channel group)
  constexpr int W_BYTES = CT * W_TAP;                                      // ... per chunk
  // ... and, with a single 64-column accumulator, tensor memory has room for THREE 48-column A stages (columns 64..207) and
  // shared memory for three 12 KB weight stages: one more chunk in flight per CTA
  constexpr int NS = (BF16 && KS == 3) ? 3 : 2;                  // A / weight stages
  constexpr int A_BASE = NS == 3 ? 64 : 128, A_STRIDE = NS == 3 ? 48 : 64;       // TMEM columns of the A ring
  constexpr int WS_STRIDE = NS == 3 ? 12288 : TS_STAGE;          // bytes between weight stages
  constexpr int TILE_BYTES = tm_tile_bytes(KS);
  constexpr int NTB = tm_tile_bufs(KS);                          // activation tile buffers in flight
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* tiles = smem;                                   // NTB activation tiles [32 ch][BH][BW]
  unsigned char* wst = smem + NTB * TILE_BYTES;                  // 2 weight stages (hi | lo image of a chunk)
  uint64_t* bars = reinterpret_cast<uint64_t*>(wst + (NS == 3 ? 3 * 12288 : 2 * TS_STAGE));
  uint64_t* a_empty = bars;          // [NS] MMAs that read A stage / weight stage completed
  uint64_t* a_full = bars + 3;       // [NS] all 8 producer warps stored their part of the chunk
  uint64_t* w_full = bars + 6;       // [NS] weight image landed
  uint64_t* t_full = bars + 9;       // [4] activation tile landed
  uint64_t* t_empty = bars + 13;     // [4] all 8 producer warps are done reading the tile
  uint64_t* d_empty = bars + 17;     // [1] the epilogue has drained the accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  float* bias_s = reinterpret_cast<float*>(bars + 20);          // [64]
  const ConvArgs& a = g.t.c;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&a_empty[i], 1); mbar_init(&a_full[i], TC_PRODUCERS / 32); mbar_init(&w_full[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], TC_PRODUCERS / 32); }
    mbar_init(d_empty, TC_PRODUCERS / 32);
    fence_barrier_init();
  }
  if (tid < NT) bias_s[tid] = (a.bias != nullptr && tid < a.Cout) ? __ldg(a.bias + tid) : 0.f;
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  pdl_trigger();          // this CTA holds its tensor memory: the next grid of the chain may be scheduled behind it
  pdl_wait();             // ... and nothing below runs before the previous grid has completed (no-op in a plain launch)

  const int HoWo = a.Ho * a.Wo;
  const int ntiles = (int)(((long long)a.B * HoWo) / TC_TM);
  const int ncc = a.Ctot / TC_KC;
  const int nchunks = ncc * T / CT;
  const int CHS = g.BW * g.BH;
  const uint32_t tile_tx = (uint32_t)(CHS * TC_KC * (SBF ? 2 : 4));
  const unsigned char* wimg = reinterpret_cast<const unsigned char*>(g.t.wprep);
  constexpr uint32_t IDESC = umma_instr_desc(TC_TM, NT, BF16 ? 1 : 2);
  constexpr uint32_t IDESC_BF = umma_instr_desc(TC_TM, NT, 1);

  // persistent over tiles blockIdx.x, blockIdx.x + gridDim.x, ...: TMEM, barriers and the TMA pipeline are set up once,
  // and the first activation tile of the next output tile is already in flight while this one runs its epilogue
  // (channel-group counter gcc and chunk counter gch run across tiles and drive the mbarrier phases)
  const int S = g.S, NG = (KS == 3 && S == 2) ? 3 : 1, TG = T / NG;   // tile loads per channel group, taps per load
  auto issue_tile = [&](int tile, int cc, int grp, int gcc) {
    const int img = tile / g.TPI;
    const int trem = tile - img * g.TPI;
    const int tyy = trem / g.TX;
    const int oy0 = tyy * g.TH, ox0 = (trem - tyy * g.TX) * g.TW;
    int s = 0, c0 = cc * TC_KC;
    if (c0 >= a.src[0].C) { c0 -= a.src[0].C; s = 1; if (c0 >= a.src[1].C) { c0 -= a.src[1].C; s = 2; } }
    const int up = (s == 0 ? a.src[0].uh : (s == 1 ? a.src[1].uh : a.src[2].uh)) == 2;
    const int y0 = S * oy0 - a.pad + grp;
    const int tb = gcc % ntb_rt;
    mbar_arrive_expect_tx(&t_full[tb], tile_tx);
    if (SBF)      // [B,H,W,64] bf16: coordinates (channel, x, y, image)
      tma_load_4d(tiles + tb * tile_stride, &g.tmap[s], c0, S * ox0 - XH, y0, img, &t_full[tb], policy_evict_first());
    else
      tma_load_4d(tiles + tb * tile_stride, &g.tmap[s], up ? ((ox0 - 2 * XH) >> 1) : (S * ox0 - XH), up ? (y0 >> 1) : y0, c0, img,
                  &t_full[tb], policy_evict_first());
  };

  if (warp == TC_PRODUCERS / 32) {
    // ===== MMA issuer warp (chunk order: channel group major, tap minor) =====
    // A barrier wait costs the issuing warp 130-260 cycles even when the barrier completed long ago (clock64 timeline,
    // profiles/r02_conv_experiments.md), so the barriers of chunk g+1 are probed (non-blocking test_wait) BEFORE the MMAs of
    // chunk g are issued -- the probe's latency hides behind the issue -- and only a failed probe is waited for afterwards.
    int gch = 0, it = 0;
    int ok_a = 0, ok_w = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      if (it >= 1) {
        mbar_wait_bounded(d_empty, (uint32_t)((it - 1) & 1));
        tc_fence_after();
      }
      for (int ch = 0; ch < nchunks; ++ch, ++gch) {
        const int buf = gch % NS, nbuf = (gch + 1) % NS;
        const uint32_t ph = (uint32_t)((gch / NS) & 1);
        if (!ok_a) mbar_wait_bounded(&a_full[buf], ph);
        if (!ok_w) mbar_wait_bounded(&w_full[buf], ph);
        const uint32_t ph1 = (uint32_t)(((gch + 1) / NS) & 1);
        if (!MIX) {
          ok_a = __shfl_sync(0xffffffffu, mbar_test_wait(&a_full[nbuf], ph1) ? 1 : 0, 0);
          ok_w = __shfl_sync(0xffffffffu, mbar_test_wait(&w_full[nbuf], ph1) ? 1 : 0, 0);
        }
        tc_fence_after();
        if (MIX) {
          uint32_t pa = 0, pw = 0;
          if (elect_one_sync()) {
            const uint32_t b_hi = smem_u32(wst) + (uint32_t)(buf * WS_STRIDE);
            mix_mma_chunk_probed<NT>(tmem_d, tmem_d + A_BASE + (uint32_t)(buf * A_STRIDE), umma_smem_desc(b_hi, NT * 16, 128), IDESC, IDESC_BF,
                                 ch != 0 ? 1u : 0u, smem_u32(&a_empty[buf]), smem_u32(&a_full[nbuf]), smem_u32(&w_full[nbuf]),
                                 ph1, pa, pw);
          }
          // the probing lane is the elected one: OR-reduce so that every lane of the warp holds its answer
          ok_a = __any_sync(0xffffffffu, pa != 0) ? 1 : 0;
          ok_w = __any_sync(0xffffffffu, pw != 0) ? 1 : 0;
          continue;
        }
        if (elect_one_sync()) {
          const uint32_t b_hi = smem_u32(wst) + (uint32_t)(buf * WS_STRIDE), b_lo = b_hi + tc_b_bytes(NT);
          const uint32_t a_hi = tmem_d + A_BASE + (uint32_t)(buf * A_STRIDE), a_lo = a_hi + 32;
          if (BF16) {
#pragma unroll
            for (int t = 0; t < CT; ++t)
#pragma unroll
              for (int j = 0; j < TC_KC / 16; ++j) {   // one MMA = K 16 = two 16-byte k-groups of 8 bf16
                const uint64_t db = umma_smem_desc(b_hi + t * W_TAP + j * 2 * (NT * 16), NT * 16, 128);
                umma_ts_f16(tmem_d, a_hi + t * 16 + j * 8, db, IDESC, (ch | t | j) != 0 ? 1u : 0u);
              }
          } else if (MIX) {
            mix_mma_chunk<NT>(tmem_d, a_hi, b_hi, IDESC, IDESC_BF, ch == 0);
          } else {
#pragma unroll
            for (int j = 0; j < TC_KC / 8; ++j) {
              const uint64_t dbh = umma_smem_desc(b_hi + j * 2 * (NT * 16), NT * 16, 128);
              const uint64_t dbl = umma_smem_desc(b_lo + j * 2 * (NT * 16), NT * 16, 128);
              umma_ts_tf32(tmem_d + 64, a_lo + j * 8, dbh, IDESC, (ch | j) != 0 ? 1u : 0u);
              umma_ts_tf32(tmem_d + 64, a_hi + j * 8, dbl, IDESC, 1u);
              umma_ts_tf32(tmem_d, a_hi + j * 8, dbh, IDESC, (ch | j) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&a_empty[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp == TC_PRODUCERS / 32 + 1) {
    // ===== TMA warp: activation tiles NTB - 1 loads ahead of the producers, the weight image of chunk g as soon as the
    //       MMAs of chunk g - 2 have released its stage.  (Issued from a producer thread, each tile load stalled all
    //       eight producer warps for ~1300 cycles -- 40 % of a stride-2 convolution's time.) =====
    const int ngroups = ncc * NG;                       // tile loads per output tile
    long long nloads = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) nloads += ngroups;
    int l_tile = blockIdx.x, l_cc = 0, l_grp = 0, l_idx = 0;      // cursor of the next tile load to issue
    auto issue_next_tile = [&]() {
      if (l_idx >= nloads) return;
      if (l_idx >= ntb_rt) mbar_wait_bounded(&t_empty[l_idx % ntb_rt], (uint32_t)(((l_idx / ntb_rt) - 1) & 1));
      if (elect_one_sync()) issue_tile(l_tile, l_cc, l_grp, l_idx);
      __syncwarp();
      ++l_idx;
      if (++l_grp == NG) { l_grp = 0; if (++l_cc == ncc) { l_cc = 0; l_tile += gridDim.x; } }
    };
    for (int i = 0; i < ntb_rt - 1; ++i) issue_next_tile();
    int gch = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      for (int cc = 0; cc < ncc; ++cc)
        for (int grp = 0; grp < NG; ++grp) {
          issue_next_tile();
          for (int tg = 0; tg < TG; tg += CT, ++gch) {
            const int buf = gch % NS;
            if (gch >= NS) mbar_wait_bounded(&a_empty[buf], (uint32_t)(((gch / NS) - 1) & 1));
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(&w_full[buf], W_BYTES);
#pragma unroll
              for (int t = 0; t < CT; ++t)
                bulk_g2s(wst + buf * WS_STRIDE + t * W_TAP, wimg + (size_t)((grp * TG + tg + t) * ncc + cc) * W_TAP, W_TAP, &w_full[buf],
                         policy_evict_last());
            }
            __syncwarp();
          }
        }
  } else {
    // ===== producer warps: thread = output pixel = TMEM lane; warps w and w+4 split a chunk's 32 channels =====
    const int row = (warp & 3) * 32 + lane;
    const int kh = warp >> 2;
    const int pr = row / g.TW, pc = row - pr * g.TW;            // position inside the tile rectangle
    const uint32_t lane_addr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    int gch = 0, gcc = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int img = tile / g.TPI;
      const int trem = tile - img * g.TPI;
      const int tyy = trem / g.TX;
      const int oy0 = tyy * g.TH, ox0 = (trem - tyy * g.TX) * g.TW;      // tile = TH x TW rectangle (tiles row-major in the image)
      for (int cc = 0; cc < ncc; ++cc)
      for (int grp = 0; grp < NG; ++grp, ++gcc) {
        const int tb = gcc % ntb_rt;
        int s = 0, c0 = cc * TC_KC;
        if (c0 >= a.src[0].C) { c0 -= a.src[0].C; s = 1; if (c0 >= a.src[1].C) { s = 2; } }
        const bool up = (s == 0 ? a.src[0].uh : (s == 1 ? a.src[1].uh : a.src[2].uh)) == 2;
        const int by = up ? ((oy0 - a.pad) >> 1) : (oy0 - a.pad), bx = up ? ((ox0 - 2 * XH) >> 1) : (S * ox0 - XH);
        const uint32_t esz = SBF ? 2u : 4u;
        const uint32_t cstride = esz * (uint32_t)CHS;
        const uint32_t tile_b = smem_u32(tiles + tb * tile_stride);
        const uint32_t tile_a = tile_b + (uint32_t)(kh * 16) * cstride;
        mbar_wait_bounded(&t_full[tb], (uint32_t)((gcc / ntb_rt) & 1));
#pragma unroll 1
        for (int tg = 0; tg < TG; tg += CT, ++gch) {
          const int buf = gch % NS;
          if constexpr (CT == 3) {
            // bf16 filter-row chunk: the three taps of row dy read the same halo rows at column offsets 0, 1, 2
            uint32_t pk[3][8];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              const int tap = grp * TG + tg + t;
              const int dy = tap / KS, dx = tap - dy * KS;
              int iy = oy0 + pr + dy - a.pad, ix = S * (ox0 + pc) + dx - a.pad;
              const bool zero = a.zi && up && ((iy | ix) & 1);
              if (up) { iy >>= 1; ix >>= 1; }
              const uint32_t qa = tile_a + esz * (uint32_t)((S == 2 ? pr : iy - by) * g.BW + (ix - bx));
              if (SBF) {
                // pixel-major bf16 tile (64-byte rows, SWIZZLE_64B): this thread's 16 channels are two 16-byte pieces that are
                // already the packed TMEM columns -- 2 LDS.128 per tap instead of 16 scalar loads + 8 conversions
                const uint32_t r = (uint32_t)((S == 2 ? pr : iy - by) * g.BW + (ix - bx));
                const uint32_t sw = (r >> 1) & 3u;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                  const uint32_t ad = tile_b + r * 64u + ((((uint32_t)(kh * 2 + q)) ^ sw) << 4);
                  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                               : "=r"(pk[t][4 * q]), "=r"(pk[t][4 * q + 1]), "=r"(pk[t][4 * q + 2]), "=r"(pk[t][4 * q + 3]) : "r"(ad));
                }
              } else {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[i]) : "r"(qa + (uint32_t)i * cstride));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const __nv_bfloat162 h2 = __floats2bfloat162_rn(zero ? 0.f : v[2 * i], zero ? 0.f : v[2 * i + 1]);
                  pk[t][i] = *reinterpret_cast<const uint32_t*>(&h2);
                }
              }
            }
            if (gch >= NS) {
              mbar_wait_bounded(&a_empty[buf], (uint32_t)(((gch / NS) - 1) & 1));
              tc_fence_after();
            }
#pragma unroll
            for (int t = 0; t < 3; ++t) tmem_st8(lane_addr + (uint32_t)(A_BASE + buf * A_STRIDE + t * 16 + kh * 8), pk[t]);
            if (tg + CT >= TG) {
              __syncwarp();
              if (lane == 0) mbar_arrive(&t_empty[tb]);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[buf]);
            continue;
          }
          const int tap = grp * TG + tg;
          const int dy = tap / KS, dx = tap - dy * KS;
          int iy = oy0 + pr + dy - a.pad, ix = S * (ox0 + pc) + dx - a.pad;
          const bool zero = a.zi && up && ((iy | ix) & 1);       // zero-inserted source (data gradient of a stride-2 convolution)
          if (up) { iy >>= 1; ix >>= 1; }
          // stride 2: the box of this filter row already holds input rows 2 oy + dy - pad, one per tile row
          // 32-bit shared-memory addresses: with generic pointers the 16 channel loads cost ~6 instructions each
          // (64-bit index arithmetic on the run-time channel stride) -- a fifth of the 3x3 kernels' instruction stream
          const uint32_t qa = tile_a + 4u * (uint32_t)((S == 2 ? pr : iy - by) * g.BW + (ix - bx));
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[i]) : "r"(qa + (uint32_t)i * cstride));
          if (zero) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
          }
          // split first, wait second: only the tcgen05.st sit between "stage free" and "stage full"
          float mhi[16];
          uint32_t mlo2[8], ma2[8];
          if (MIX) mix_split(v, mhi, mlo2, ma2);
          if (gch >= NS) {
            mbar_wait_bounded(&a_empty[buf], (uint32_t)(((gch / NS) - 1) & 1));   // MMAs of chunk gch-NS have read the stage
            tc_fence_after();
          }
          if (BF16) {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);   // .x (low half) = even channel
              pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            tmem_st8(lane_addr + (uint32_t)(A_BASE + buf * A_STRIDE + kh * 8), pk);
          } else if (MIX) {
            const uint32_t sa = lane_addr + (uint32_t)(A_BASE + buf * A_STRIDE);
            tmem_st16(sa + (uint32_t)(kh * 16), mhi);
            tmem_st8(sa + (uint32_t)(32 + kh * 8), mlo2);
            tmem_st8(sa + (uint32_t)(48 + kh * 8), ma2);
          } else {
            float hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              hi[i] = __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u);
              lo[i] = v[i] - hi[i];
            }
            const uint32_t col = (uint32_t)(A_BASE + buf * A_STRIDE + kh * 16);
            tmem_st16(lane_addr + col, hi);
            tmem_st16(lane_addr + col + 32, lo);
          }
          if (tg == TG - 1) {                      // the tile's last reads are in registers: hand the buffer back
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[tb]);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[buf]);
        }
      }
      const int last = gch - 1;
      mbar_wait_bounded(&a_empty[last % NS], (uint32_t)((last / NS) & 1));
      tc_fence_after();

      // ---- epilogue (NCHW): lane = pixel, coalesced per channel ----
      const int wq = warp & 3, half = warp >> 2;
      const int oy = oy0 + pr, ox = ox0 + pc;
      const bool relu = a.act == CONV_ACT_RELU;
#pragma unroll 1
      for (int pass = 0; pass < NT / 32; ++pass) {
        const int cb = half * (NT / 2) + pass * 16;
        if (cb >= a.Cout) break;                 // a 32-channel layer on a 64-row weight image padded with zeros
        float acc[16], part[16];
        const uint32_t taddr = tmem_d + ((uint32_t)(wq * 32) << 16) + (uint32_t)cb;
        tmem_ld_cols<16>(taddr, acc);
        if (BF16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) part[j] = 0.f;
        } else {
          tmem_ld_cols<16>(taddr + 64, part);
        }
        const float4* b4 = reinterpret_cast<const float4*>(bias_s + cb);       // warp-uniform: broadcast loads
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bs = b4[j >> 2];
          acc[j] += part[j] + bs.x; acc[j + 1] += part[j + 1] + bs.y; acc[j + 2] += part[j + 2] + bs.z; acc[j + 3] += part[j + 3] + bs.w;
        }
        if (a.skip != nullptr && a.skip_pre) {   // residual block: the identity is added BEFORE the activation
          const float* sk = a.skip + ((size_t)img * (a.out_cstride ? a.out_cstride : a.Cout) + cb) * HoWo + (size_t)oy * a.Wo + ox;
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += __ldg(sk + (size_t)j * HoWo);
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = fmaxf(acc[j], 0.f);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = tc_act(acc[j], a.act, a.act_scale);
        }
        const size_t o0 = ((size_t)img * (a.out_cstride ? a.out_cstride : a.Cout) + cb) * HoWo + (size_t)oy * a.Wo + ox;
        if (a.skip != nullptr && !a.skip_pre) {
          if (BF16 && a.skip_bf16) {          // [B,Ho,Wo,64] bf16: this pixel's 16 channels are 32 contiguous bytes
            const uint4* sk = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.skip) +
                                                             (((size_t)img * a.Ho + oy) * a.Wo + ox) * 64 + cb);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const uint4 u = __ldg(sk + q);
              const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                acc[8 * q + 2 * i] += __uint_as_float(w4[i] << 16);
                acc[8 * q + 2 * i + 1] += __uint_as_float(w4[i] & 0xFFFF0000u);
              }
            }
          } else {
            const float* sk = a.skip + o0;
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] += __ldg(sk + (size_t)j * HoWo);
          }
        }
        if (BF16 && a.out_bf16) {
          uint4* po = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + (((size_t)img * a.Ho + oy) * a.Wo + ox) * 64 + cb);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t w4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[8 * q + 2 * i], acc[8 * q + 2 * i + 1]);
              w4[i] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            po[q] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
          continue;
        }
        float* po = a.out + o0;
#pragma unroll
        for (int j = 0; j < 16; ++j) po[(size_t)j * HoWo] = acc[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty);       // the MMA warp may overwrite the accumulator with the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 256);
}


// ---- host side of the TMA path ----
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder() {
  // function-local static: initialised exactly once, thread-safe (C++11)
  static const tmap_encode_fn fn = []() -> tmap_encode_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<tmap_encode_fn>(p);
    return nullptr;
  }();
  return fn;
}

// cuTensorMapEncodeTiled costs a driver call per map and a forward needs ~25 of them; the head's buffers keep their
// addresses from call to call (caller-owned workspace), so encoded maps are kept in a small per-thread cache keyed on
// every argument of the encode.
struct TmapKey {
  const void* ptr;
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], es[4];
  int rank, swizzle, bf16;
};
static bool tmap_cached(CUtensorMap* out, int rank, const void* ptr, const cuuint64_t* dims, const cuuint64_t* strides,
                        const cuuint32_t* box, const cuuint32_t* es, CUtensorMapSwizzle swz, bool bf16 = false) {
  constexpr int CAP = 96;
  struct Entry { TmapKey k; CUtensorMap m; };
  static thread_local Entry cache[CAP];
  static thread_local int used = 0, next = 0;
  TmapKey k;
  memset(&k, 0, sizeof(k));
  k.ptr = ptr; k.rank = rank; k.swizzle = (int)swz; k.bf16 = bf16 ? 1 : 0;
  for (int i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; k.es[i] = es[i]; }
  for (int i = 0; i + 1 < rank; ++i) k.strides[i] = strides[i];
  for (int i = 0; i < used; ++i)
    if (memcmp(&cache[i].k, &k, sizeof(k)) == 0) { *out = cache[i].m; return true; }
  tmap_encode_fn enc = tmap_encoder();
  if (enc == nullptr) return false;
  CUtensorMap m;
  if (enc(&m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, es,
          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  const int slot = used < CAP ? used++ : (next++ % CAP);
  cache[slot].k = k; cache[slot].m = m;
  *out = m;
  return true;
}

// NCHW convolution whose tile is a rectangle of one image, every source either full or half resolution
static bool conv_tma_plan(int KS, const ConvArgs& a, ConvTmaArgs* g, int mode = CM_TF32X3) {
  // bf16-stored activations (bf16 mode of the head): 3x3 bf16-operand kernel only, every source of the launch bf16, no up-sampling
  const bool sbf = a.src[0].bf16 != 0;
  if ((sbf || a.out_bf16 || a.skip_bf16) && !(KS == 3 && mode == CM_BF16 && a.Cout == 64)) return false;
  const int XH = KS == 3 ? (sbf ? 1 : TM_XH) : 0;       // bf16 sources are NHWC: x is not the innermost TMA coordinate, no alignment halo
  if (a.sh != a.sw || (a.sh != 1 && !(a.sh == 2 && KS == 3)) || a.out_nhwc || a.wimg_stride != 0 || (a.Cout != 64 && a.Cout != 32)) return false;
  if (a.pad != (KS == 3 ? 1 : 0)) return false;
  const int S = a.sh;
  // 1x1: whole rows (up to 128 pixels); 3x3: at most 64 columns so that the halo box of 32 channels fits the tile buffer
  const int TWmax = KS == 3 ? 64 : 128;
  const int TW = a.Wo < TWmax ? a.Wo : TWmax;
  if (TW < 16 || 128 % TW || a.Wo % TW) return false;
  const int TH = 128 / TW;
  if (a.Ho % TH) return false;
  // stride 1: one halo box serves all taps.  stride 2: one box per filter row -- full-width input rows 2 oy + dy - 1
  // (the tensor map walks rows with element stride 2; TMA cannot stride the innermost dimension, so the threads read
  // columns 2 ox + dx - 1 themselves)
  const int BW = KS == 3 ? S * TW + 2 * XH : TW, BH = (KS == 3 && S == 1) ? TH + 2 : TH;
  if (BW > 256 || BH > 256 || (size_t)BW * BH * TC_KC * (sbf ? 2 : 4) > (size_t)(sbf ? TM_TILE_BF : tm_tile_bytes(KS))) return false;
  if ((a.out_bf16 || a.skip_bf16) && a.Cout != 64) return false;
  for (int s = 0; s < 3; ++s) {
    const ConvSrc& sc = a.src[s];
    if (sc.C == 0) continue;
    if (sc.nhwc || sc.C % TC_KC || sc.uh != sc.uw || (sc.uh != 1 && sc.uh != 2) || (S == 2 && sc.uh != 1)) return false;
    if ((sc.bf16 != 0) != sbf || (sbf && (sc.uh != 1 || sc.C != 64))) return false;
    if (sc.H * sc.uh != a.Ho * S || sc.W * sc.uw != a.Wo * S) return false;
    if (sbf) {
      // [B, H, W, 64] bf16: box = 32 channels (64 bytes, SWIZZLE_64B: a thread's 32-byte reads are conflict-free) x BW x BH rows
      if ((uintptr_t)sc.ptr & 15) return false;
      const cuuint64_t dims[4] = {(cuuint64_t)sc.C, (cuuint64_t)sc.W, (cuuint64_t)sc.H, (cuuint64_t)a.B};
      const cuuint64_t strides[3] = {(cuuint64_t)sc.C * 2, (cuuint64_t)sc.W * sc.C * 2, (cuuint64_t)sc.W * sc.H * sc.C * 2};
      const cuuint32_t box[4] = {(cuuint32_t)TC_KC, (cuuint32_t)BW, (cuuint32_t)(BH * S), 1};
      const cuuint32_t es[4] = {1, 1, (cuuint32_t)S, 1};
      if (!tmap_cached(&g->tmap[s], 4, sc.ptr, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_64B, true)) return false;
      continue;
    }
    if ((sc.W * 4) % 16 || ((uintptr_t)sc.ptr & 15)) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)sc.W, (cuuint64_t)sc.H, (cuuint64_t)sc.C, (cuuint64_t)a.B};
    const cuuint64_t strides[3] = {(cuuint64_t)sc.W * 4, (cuuint64_t)sc.W * sc.H * 4, (cuuint64_t)sc.W * sc.H * sc.C * 4};
    const cuuint32_t box[4] = {(cuuint32_t)BW, (cuuint32_t)(BH * S), (cuuint32_t)TC_KC, 1};
    const cuuint32_t es[4] = {1, (cuuint32_t)S, 1, 1};
    if (!tmap_cached(&g->tmap[s], 4, sc.ptr, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_NONE)) return false;
  }
  g->TW = TW; g->TH = TH; g->BW = BW; g->BH = BH; g->S = S; g->XH = XH;
  g->TX = a.Wo / TW; g->TPI = (a.Wo / TW) * (a.Ho / TH);
  return true;
}

// ------------------------------------------------------------------------------------------------
// Row-major linear layers (feat_linear.1, the QK^T score with per-image weights, DGAB's Mlp fc1/fc2):
// out[R, Cout] = act(A[R, K] . W^T + b) (+ skip), R a multiple of 128.  Same TS scheme as the convolutions --
// thread = row = TMEM lane -- with the [128 rows x 32 k] operand tile brought in by one 2-D TMA box per chunk
// (128-byte rows, SWIZZLE_128B so that the per-row 16-byte reads of a warp hit distinct banks), persistent
// over row tiles.  When K <= 64 the split A operand of a tile stays in tensor memory while the kernel walks the
// Cout / NT column blocks, so fc1 (K = 64, Cout = 256) reads and splits its input once instead of four times;
// the weight images stream through a 4-stage shared-memory ring filled by the MMA warp three items ahead.
// ------------------------------------------------------------------------------------------------
struct LinTmaArgs {
  ConvTcArgs t;
  CUtensorMap tmap;
  int rows_per_img;        // rows that share one weight image (per-image weights), else 0
  int col_split;           // 1: K > 64 with several column blocks -- blockIdx.y owns ONE 64/32-column block of the output (the
                           //    streamed-K path of this kernel handles a single block per CTA); 0: a CTA walks all blocks
};
constexpr int LN_WSTAGES = 4;
__host__ __device__ constexpr int ln_smem_bytes() { return 2 * 16384 + LN_WSTAGES * TS_STAGE + 1536 + 1024; }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}

template <int NT>
__global__ void __launch_bounds__(TC_THREADS, 2) lin_tma_kernel(const __grid_constant__ LinTmaArgs g) {
  constexpr int W_BYTES = 2 * NT * TC_KC * 4;                      // hi | lo image of one (block, chunk)
  constexpr int HC = NT / 2;                                       // output columns per producer-warp half
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* tiles = smem;                                     // 2 x [128 rows][32 floats], 128B-swizzled
  unsigned char* wst = smem + 2 * 16384;                           // LN_WSTAGES weight stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(wst + LN_WSTAGES * TS_STAGE);
  uint64_t* a_empty = bars;                  // [2]
  uint64_t* a_full = bars + 2;               // [2]
  uint64_t* t_full = bars + 4;               // [2]
  uint64_t* t_empty = bars + 6;              // [2]
  uint64_t* w_full = bars + 8;               // [4]
  uint64_t* w_empty = bars + 12;             // [4]
  uint64_t* d_full = bars + 16;
  uint64_t* d_empty = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  float* bias_s = reinterpret_cast<float*>(bars + 20);             // [Cout <= 256]
  const ConvArgs& a = g.t.c;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_empty[i], 1); mbar_init(&a_full[i], TC_PRODUCERS / 32);
      mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], TC_PRODUCERS / 32);
    }
    for (int i = 0; i < LN_WSTAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(d_full, 1); mbar_init(d_empty, TC_PRODUCERS / 32);
    fence_barrier_init();
  }
  const int nb0 = g.col_split ? (int)blockIdx.y : 0;              // first column block of this CTA
  for (int i = tid; i < (g.col_split ? NT : a.Cout) && i < 256; i += TC_THREADS)
    bias_s[i] = a.bias != nullptr ? __ldg(a.bias + nb0 * NT + i) : 0.f;
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  pdl_trigger();          // this CTA holds its tensor memory: the next grid of the chain may be scheduled behind it
  pdl_wait();             // ... and nothing below runs before the previous grid has completed (no-op in a plain launch)

  const long long R = (long long)a.B * a.Ho * a.Wo;
  const int ntiles = (int)(R / TC_TM);
  // split-K (a.splitk > 1, small-row GEMMs such as the NRTR decode steps: a launch is a few dozen CTAs that would each walk all
  // of K serially at ~1 us per chunk): blockIdx.z owns nchunks consecutive chunks and writes its own partial output
  const int ksp = a.splitk > 1 ? a.splitk : 1;
  const int nchunks_total = a.Ctot / TC_KC;
  const int nchunks = nchunks_total / ksp;
  const int ch0 = (int)blockIdx.z * nchunks;
  const int nblocks = g.col_split ? 1 : a.Cout / NT;
  const float* wbase = g.t.wprep;
  constexpr uint32_t IDESC = umma_instr_desc(TC_TM, NT, 2);

  if (warp == TC_PRODUCERS / 32) {
    // ===== MMA issuer warp: items (tile, block, chunk); also streams the weight images =====
    // prefetch cursor: three items ahead of the issue cursor
    int p_tile = blockIdx.x, p_nb = 0, p_ch = 0;
    int wi_load = 0;
    auto load_next_weights = [&]() {                       // whole warp: the cursor stays warp-uniform
      if (p_tile >= ntiles) return;
      const int st = wi_load & (LN_WSTAGES - 1);
      if (wi_load >= LN_WSTAGES) mbar_wait_bounded(&w_empty[st], (uint32_t)(((wi_load / LN_WSTAGES) - 1) & 1));
      const float* wsrc = wbase + (size_t)(g.rows_per_img > 0 ? ((long long)p_tile * TC_TM) / g.rows_per_img : 0) * (size_t)a.wimg_stride +
                          (size_t)((p_nb + nb0) * nchunks_total + ch0 + p_ch) * (2 * NT * TC_KC);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&w_full[st], W_BYTES);
        bulk_g2s(wst + st * TS_STAGE, wsrc, W_BYTES, &w_full[st], policy_evict_last());
      }
      __syncwarp();
      ++wi_load;
      if (++p_ch == nchunks) { p_ch = 0; if (++p_nb == nblocks) { p_nb = 0; p_tile += gridDim.x; } }
    };
    load_next_weights(); load_next_weights(); load_next_weights();
    // operand tiles are issued from here as well, two chunks ahead: when this warp has seen a_full of chunk c, the
    // producers have already released that chunk's tile buffer (they arrive on t_empty before a_full), so the load of
    // chunk c + 2 never waits -- and the eight producer warps no longer stall behind one thread's TMA bookkeeping
    int t_tile = blockIdx.x, t_ch = 0, t_idx = 0;
    auto load_next_tile = [&]() {
      if (t_tile >= ntiles) return;
      const int tb = t_idx & 1;
      if (t_idx >= 2) mbar_wait_bounded(&t_empty[tb], (uint32_t)(((t_idx >> 1) - 1) & 1));
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&t_full[tb], 16384);
        tma_load_2d(tiles + tb * 16384, &g.tmap, (ch0 + t_ch) * TC_KC, t_tile * TC_TM, &t_full[tb], policy_evict_first());
      }
      __syncwarp();
      ++t_idx;
      if (++t_ch == nchunks) { t_ch = 0; t_tile += gridDim.x; }
    };
    load_next_tile(); load_next_tile();
    int wi = 0, di = 0, it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      for (int nb = 0; nb < nblocks; ++nb, ++di) {
        if (di >= 1) {
          mbar_wait_bounded(d_empty, (uint32_t)((di - 1) & 1));
          tc_fence_after();
        }
        for (int ch = 0; ch < nchunks; ++ch, ++wi) {
          const int st = wi & (LN_WSTAGES - 1);
          const int ga = it * nchunks + ch;                 // A-stage use counter of this chunk
          mbar_wait_bounded(&w_full[st], (uint32_t)((wi / LN_WSTAGES) & 1));
          if (nb == 0) { mbar_wait_bounded(&a_full[ga & 1], (uint32_t)((ga >> 1) & 1)); load_next_tile(); }
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t b_hi = smem_u32(wst) + (uint32_t)(st * TS_STAGE), b_lo = b_hi + NT * TC_KC * 4;
            const uint32_t a_hi = tmem_d + 128 + (uint32_t)((ga & 1) * 64), a_lo = a_hi + 32;
#pragma unroll
            for (int j = 0; j < TC_KC / 8; ++j) {
              const uint64_t dbh = umma_smem_desc(b_hi + j * 2 * (NT * 16), NT * 16, 128);
              const uint64_t dbl = umma_smem_desc(b_lo + j * 2 * (NT * 16), NT * 16, 128);
              umma_ts_tf32(tmem_d + 64, a_lo + j * 8, dbh, IDESC, (ch | j) != 0 ? 1u : 0u);
              umma_ts_tf32(tmem_d + 64, a_hi + j * 8, dbl, IDESC, 1u);
              umma_ts_tf32(tmem_d, a_hi + j * 8, dbh, IDESC, (ch | j) != 0 ? 1u : 0u);
            }
            umma_commit(&w_empty[st]);
            if (nb == nblocks - 1) umma_commit(&a_empty[ga & 1]);
            if (ch == nchunks - 1) umma_commit(d_full);
          }
          __syncwarp();
          load_next_weights();
        }
      }
    }
  } else {
    // ===== producer / epilogue warps =====
    const int row = (warp & 3) * 32 + lane;
    const int kh = warp >> 2;
    const uint32_t lane_addr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    int gcc = 0, ga = 0, di = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int ch = 0; ch < nchunks; ++ch, ++gcc, ++ga) {
        const int tb = gcc & 1;
        mbar_wait_bounded(&t_full[tb], (uint32_t)((gcc >> 1) & 1));
        // this row's 16 floats of the chunk half: 16-byte pieces 4kh .. 4kh+3, stored at piece ^ (row & 7)
        const unsigned char* rp = tiles + tb * 16384 + row * 128;
        float4 v4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v4[i] = *reinterpret_cast<const float4*>(rp + (((kh * 4 + i) ^ (row & 7)) << 4));
        const float* v = reinterpret_cast<const float*>(v4);
        float hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          hi[i] = __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u);
          lo[i] = v[i] - hi[i];
        }
        const int buf = ga & 1;
        if (ga >= 2) {
          mbar_wait_bounded(&a_empty[buf], (uint32_t)(((ga >> 1) - 1) & 1));
          tc_fence_after();
        }
        const uint32_t col = (uint32_t)(128 + buf * 64 + kh * 16);
        tmem_st16(lane_addr + col, hi);
        tmem_st16(lane_addr + col + 32, lo);
        // only now is the tile buffer released: the tcgen05.st above consume every value read from it, so no shared-memory
        // load of this warp can still be in flight when the TMA producer is allowed to overwrite the buffer
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[tb]);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[buf]);
      }
      const long long m = (long long)tile * TC_TM + row;
      for (int nb = 0; nb < nblocks; ++nb, ++di) {
        mbar_wait_bounded(d_full, (uint32_t)(di & 1));
        tc_fence_after();
        const int cbl = nb * NT + kh * HC;                        // ... within the columns this CTA owns (bias_s index)
        const int cb0 = nb0 * NT + cbl;                           // first output column of this thread's part
        const size_t o0 = (size_t)m * a.Cout + cb0 + (size_t)blockIdx.z * (size_t)R * a.Cout;
#pragma unroll 1
        for (int pass = 0; pass < HC / 16; ++pass) {
          float acc[16], part[16];
          const uint32_t taddr = lane_addr + (uint32_t)(kh * HC + pass * 16);
          tmem_ld_cols<16>(taddr, acc);
          tmem_ld_cols<16>(taddr + 64, part);
          const int cb = cbl + pass * 16;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 bs = *reinterpret_cast<const float4*>(bias_s + cb + j);      // shared memory, warp-uniform
            float4 r = make_float4(acc[j] + (part[j] + bs.x), acc[j + 1] + (part[j + 1] + bs.y), acc[j + 2] + (part[j + 2] + bs.z),
                                   acc[j + 3] + (part[j + 3] + bs.w));
            if (a.act == CONV_ACT_GELU) { r.x = tc_gelu_fast(r.x); r.y = tc_gelu_fast(r.y); r.z = tc_gelu_fast(r.z); r.w = tc_gelu_fast(r.w); }
            else if (a.act == CONV_ACT_TANH) { r.x = tanhf(r.x * a.act_scale); r.y = tanhf(r.y * a.act_scale); r.z = tanhf(r.z * a.act_scale); r.w = tanhf(r.w * a.act_scale); }
            else if (a.act == CONV_ACT_RELU) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
            if (a.skip != nullptr) {
              const float4 sk = __ldg(reinterpret_cast<const float4*>(a.skip + o0 + pass * 16 + j));
              r.x += sk.x; r.y += sk.y; r.z += sk.z; r.w += sk.w;
            }
            *reinterpret_cast<float4*>(a.out + o0 + pass * 16 + j) = r;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d_empty);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 256);
}

// ------------------------------------------------------------------------------------------------
// DGAB's Mlp fused: out = x1 + fc2(GELU(fc1(v) + b1)) + b2 over rows of 64 (DGAB.py:17-23,76) without the
// [R,256] hidden tensor ever leaving the SM (it was 268 MB written + 268 MB read per step).  One 512-column CTA
// per SM, persistent over 128-row tiles:
//   TMEM   A1 [0,128)  : the split input tile (K = 64: two chunks of hi 32 | lo 32), kept for all four hidden blocks
//          D1 [128,256): two fc1 accumulators (64-wide hidden blocks alternate; correction terms merged in)
//          A2 [256,384): GELU(D1 + b1) split again -- the A operand of fc2 for that block (K = 64)
//          D2 [384,512): fc2 accumulator, summed over the four hidden blocks
//   smem   fc1 weight images resident (128 KB), fc2 weight chunks through a 4-stage ring, the input tile's two chunks
//   warps  0-15 produce A1, run the per-block GELU epilogue (D1 -> A2) and the final epilogue; warp 16 issues MMAs
//          (fc1 of block j+1 is issued before fc2 of block j, so it overlaps block j's GELU); warp 17 feeds TMA.
// ------------------------------------------------------------------------------------------------
struct MlpFusedArgs {
  CUtensorMap tmap;                  // v [R, 64] row-major, box 32 x 128, SWIZZLE_128B
  const float *w1img, *w2img;        // wprep images of fc1 (4 blocks x 2 chunks) and fc2 (8 chunks), NT = 64
  const float *b1, *b2, *skip;
  float* out;
  long long R;
};
constexpr int MF_WARPS_E = 16;
constexpr int MF_THREADS = (MF_WARPS_E + 2) * 32;
constexpr int MF_W1_BYTES = 8 * TS_STAGE, MF_W2_STAGES = 4;
constexpr int MF_SMEM = MF_W1_BYTES + MF_W2_STAGES * TS_STAGE + 2 * 16384 + 1536 + 1024;

__global__ void __launch_bounds__(MF_THREADS, 1) mlp_fused_kernel(const __grid_constant__ MlpFusedArgs g) {
  constexpr int NT = 64;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* w1s = smem;                                        // [4 blocks][2 chunks] x 16 KB
  unsigned char* w2s = smem + MF_W1_BYTES;                          // ring of fc2 chunks
  unsigned char* vts = w2s + MF_W2_STAGES * TS_STAGE;               // input tile: chunk 0 | chunk 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(vts + 2 * 16384);
  uint64_t* w1_full = bars;            // 1
  uint64_t* v_full = bars + 1;         // [2]
  uint64_t* v_empty = bars + 3;        // [2]
  uint64_t* a1_full = bars + 5;
  uint64_t* a1_empty = bars + 6;
  uint64_t* d1_full = bars + 7;        // [2]
  uint64_t* d1_empty = bars + 9;       // [2]
  uint64_t* a2_full = bars + 11;
  uint64_t* a2_empty = bars + 12;
  uint64_t* d2_full = bars + 13;
  uint64_t* d2_empty = bars + 14;
  uint64_t* w2_full = bars + 15;       // [4]
  uint64_t* w2_empty = bars + 19;      // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);
  float* b1s = reinterpret_cast<float*>(bars + 24);               // [256] fc1 bias (warp-uniform float4 reads)
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (tid == 0) {
    mbar_init(w1_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], MF_WARPS_E); }
    mbar_init(a1_full, MF_WARPS_E); mbar_init(a1_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], MF_WARPS_E); }
    mbar_init(a2_full, MF_WARPS_E); mbar_init(a2_empty, 1);
    mbar_init(d2_full, 1); mbar_init(d2_empty, MF_WARPS_E);
    for (int i = 0; i < MF_W2_STAGES; ++i) { mbar_init(&w2_full[i], 1); mbar_init(&w2_empty[i], 1); }
    fence_barrier_init();
  }
  if (tid < 256) b1s[tid] = __ldg(g.b1 + tid);
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  pdl_trigger();          // this CTA holds its tensor memory: the next grid of the chain may be scheduled behind it
  pdl_wait();             // ... and nothing below runs before the previous grid has completed (no-op in a plain launch)
  const int ntiles = (int)(g.R / TC_TM);
  const int n_my = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  constexpr uint32_t IDESC = umma_instr_desc(TC_TM, NT, 2);
  constexpr uint32_t PART = NT * TC_KC * 4;                          // bytes of the hi (or lo) image of a chunk

  if (warp == MF_WARPS_E) {
    // ===== MMA issuer warp =====
    if (n_my > 0) mbar_wait_bounded(w1_full, 0);
    // d_corr == d_main merges the 3xTF32 correction terms into the one accumulator (fc1: K = 64, 24 accumulation steps)
    auto mma_chunk = [&](uint32_t d_main, uint32_t d_corr, uint32_t a_base, uint32_t b_base, bool first) {   // one elected lane
#pragma unroll
      for (int kk = 0; kk < TC_KC / 8; ++kk) {
        const uint64_t dbh = umma_smem_desc(b_base + kk * 2 * (NT * 16), NT * 16, 128);
        const uint64_t dbl = umma_smem_desc(b_base + PART + kk * 2 * (NT * 16), NT * 16, 128);
        const uint32_t a_hi = a_base + kk * 8, a_lo = a_hi + 32;
        umma_ts_tf32(d_main, a_hi, dbh, IDESC, (first && kk == 0) ? 0u : 1u);
        umma_ts_tf32(d_corr, a_lo, dbh, IDESC, (first && kk == 0 && d_corr != d_main) ? 0u : 1u);
        umma_ts_tf32(d_corr, a_hi, dbl, IDESC, 1u);
      }
    };
    auto issue_fc1 = [&](int gj) {           // hidden block gj & 3 of tile gj >> 2: D1[gj & 1] = A1 . W1_j^T
      const int db = gj & 1;
      if (gj >= 2) { mbar_wait_bounded(&d1_empty[db], (uint32_t)(((gj >> 1) - 1) & 1)); }
      tc_fence_after();
      if (elect_one_sync()) {
        const int j = gj & 3;
        const uint32_t d1 = tmem_d + 128 + (uint32_t)(db * 64);
        for (int c = 0; c < 2; ++c)
          mma_chunk(d1, d1, tmem_d + (uint32_t)(c * 64), smem_u32(w1s) + (uint32_t)((j * 2 + c) * TS_STAGE), c == 0);
        umma_commit(&d1_full[db]);
        if (j == 3) umma_commit(a1_empty);
      }
      __syncwarp();
    };
    for (int it = 0; it < n_my; ++it) {
      mbar_wait_bounded(a1_full, (uint32_t)(it & 1));
      issue_fc1(it * 4);
      for (int j = 0; j < 4; ++j) {
        const int gj = it * 4 + j;
        if (j < 3) issue_fc1(gj + 1);        // overlaps the GELU epilogue of block j
        mbar_wait_bounded(a2_full, (uint32_t)(gj & 1));
        if (j == 0 && it >= 1) mbar_wait_bounded(d2_empty, (uint32_t)((it - 1) & 1));
        for (int c = 0; c < 2; ++c) {
          const int wi = gj * 2 + c, st = wi & (MF_W2_STAGES - 1);
          mbar_wait_bounded(&w2_full[st], (uint32_t)((wi / MF_W2_STAGES) & 1));
          tc_fence_after();
          if (elect_one_sync()) {
            mma_chunk(tmem_d + 384, tmem_d + 448, tmem_d + 256 + (uint32_t)(c * 64), smem_u32(w2s) + (uint32_t)(st * TS_STAGE), j == 0 && c == 0);
            umma_commit(&w2_empty[st]);
            if (c == 1) { umma_commit(a2_empty); if (j == 3) umma_commit(d2_full); }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == MF_WARPS_E + 1) {
    // ===== TMA warp: resident fc1 images once, then input-tile chunks and fc2 chunks, polled round-robin =====
    if (n_my > 0 && elect_one_sync()) {
      mbar_arrive_expect_tx(w1_full, MF_W1_BYTES);
      for (int i = 0; i < 8; ++i)
        bulk_g2s(w1s + i * TS_STAGE, reinterpret_cast<const unsigned char*>(g.w1img) + (size_t)i * TS_STAGE, TS_STAGE, w1_full,
                 policy_evict_last());
    }
    __syncwarp();
    const int total_v = n_my * 2, total_w = n_my * 8;
    int vi = 0, wi = 0;
    while (vi < total_v || wi < total_w) {
      if (vi < total_v) {
        const int it = vi >> 1, c = vi & 1;
        int ok = 1;
        if (it >= 1) ok = mbar_test_wait(&v_empty[c], (uint32_t)((it - 1) & 1)) ? 1 : 0;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok) {
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&v_full[c], 16384);
            tma_load_2d(vts + c * 16384, &g.tmap, c * TC_KC, (int)((blockIdx.x + (long long)it * gridDim.x) * TC_TM), &v_full[c],
                        policy_evict_first());
          }
          __syncwarp();
          ++vi;
        }
      }
      if (wi < total_w) {
        const int st = wi & (MF_W2_STAGES - 1);
        int ok = 1;
        if (wi >= MF_W2_STAGES) ok = mbar_test_wait(&w2_empty[st], (uint32_t)(((wi / MF_W2_STAGES) - 1) & 1)) ? 1 : 0;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok) {
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&w2_full[st], TS_STAGE);
            bulk_g2s(w2s + st * TS_STAGE, reinterpret_cast<const unsigned char*>(g.w2img) + (size_t)(wi & 7) * TS_STAGE, TS_STAGE,
                     &w2_full[st], policy_evict_last());
          }
          __syncwarp();
          ++wi;
        }
      }
    }
  } else {
    // ===== producer / epilogue warps: lane quarter q = warp & 3 (TMEM lanes), column part p = warp >> 2 =====
    const int q = warp & 3, p = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16);
    // A1 of tile `it`: this thread's 8 floats of each chunk (pieces 2p, 2p+1 of its 128-byte row)
    auto produce_a1 = [&](int it) {
      if (it >= 1) { mbar_wait_bounded(a1_empty, (uint32_t)((it - 1) & 1)); tc_fence_after(); }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        mbar_wait_bounded(&v_full[c], (uint32_t)(it & 1));
        const unsigned char* rp = vts + c * 16384 + row * 128;
        const float4 x0 = *reinterpret_cast<const float4*>(rp + (((2 * p) ^ (row & 7)) << 4));
        const float4 x1 = *reinterpret_cast<const float4*>(rp + (((2 * p + 1) ^ (row & 7)) << 4));
        const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          hi[i] = __float_as_uint(v[i]) & 0xFFFFE000u;
          lo[i] = __float_as_uint(v[i] - __uint_as_float(hi[i]));
        }
        tmem_st8(lane_addr + (uint32_t)(c * 64 + p * 8), hi);
        tmem_st8(lane_addr + (uint32_t)(c * 64 + 32 + p * 8), lo);
        __syncwarp();
        if (lane == 0) mbar_arrive(&v_empty[c]);       // after the tcgen05.st that consumed the loads
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a1_full);
    };
    if (n_my > 0) produce_a1(0);
    for (int it = 0; it < n_my; ++it) {
      const long long m = ((long long)blockIdx.x + (long long)it * gridDim.x) * TC_TM + row;
      // ---- per hidden block: D1 -> + b1 -> GELU -> split -> A2 ----
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const int gj = it * 4 + j;
        const int db = gj & 1;
        mbar_wait_bounded(&d1_full[db], (uint32_t)((gj >> 1) & 1));
        tc_fence_after();
        float acc[16];
        tmem_ld_cols<16>(lane_addr + (uint32_t)(128 + db * 64 + p * 16), acc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d1_empty[db]);       // D1 is in registers: fc1 of block gj + 2 may overwrite it
        float hi[16], lo[16];
        const float4* bq = reinterpret_cast<const float4*>(b1s + j * 64 + p * 16);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 b4 = bq[i >> 2];
          const float bj = (i & 3) == 0 ? b4.x : (i & 3) == 1 ? b4.y : (i & 3) == 2 ? b4.z : b4.w;
          const float h = tc_gelu_fast(acc[i] + bj);
          hi[i] = __uint_as_float(__float_as_uint(h) & 0xFFFFE000u);
          lo[i] = h - hi[i];
        }
        if (gj >= 1) { mbar_wait_bounded(a2_empty, (uint32_t)((gj - 1) & 1)); tc_fence_after(); }
        const uint32_t col = (uint32_t)(256 + (p >> 1) * 64 + (p & 1) * 16);
        tmem_st16(lane_addr + col, hi);
        tmem_st16(lane_addr + col + 32, lo);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a2_full);
      }
      // the next tile's A1 goes in before this tile's final epilogue (A1 was released by fc1 of block 3), so the MMA
      // warp starts the next tile while fc2 of block 3 drains; the residual is fetched before waiting for D2
      if (it + 1 < n_my) produce_a1(it + 1);
      const size_t o0 = (size_t)m * 64 + p * 16;
      float4 skv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) skv[i] = __ldg(reinterpret_cast<const float4*>(g.skip + o0 + 4 * i));
      // ---- final epilogue: D2 + b2 + skip -> out[m, p*16 .. p*16+15] ----
      mbar_wait_bounded(d2_full, (uint32_t)(it & 1));
      tc_fence_after();
      {
        float acc[16], part[16];
        tmem_ld_cols<16>(lane_addr + (uint32_t)(384 + p * 16), acc);
        tmem_ld_cols<16>(lane_addr + (uint32_t)(448 + p * 16), part);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d2_empty);
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 bs = __ldg(reinterpret_cast<const float4*>(g.b2 + p * 16 + i));
          const float4 sk = skv[i >> 2];
          float4 r;
          r.x = (acc[i] + (part[i] + bs.x)) + sk.x; r.y = (acc[i + 1] + (part[i + 1] + bs.y)) + sk.y;
          r.z = (acc[i + 2] + (part[i + 2] + bs.z)) + sk.z; r.w = (acc[i + 3] + (part[i + 3] + bs.w)) + sk.w;
          *reinterpret_cast<float4*>(g.out + o0 + i) = r;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 512);
}

// ------------------------------------------------------------------------------------------------
// down0 + down1 + down2 + grid()/down_feat in ONE kernel (reference tps_pp.py:538-540,548,560-562,581-585):
//     f0 = relu(W0 o0 + b0)   f1 = relu(W1 o1 + b1)   f2 = relu(W2 x + b2)
//     feat_grid = relu(Wf cat(f0, f1, up2(f2)) + bf)
// As four launches these wrote f0/f1 (268 MB each at B = 256) and read them back for down_feat, re-read f2 four
// times for the nearest up-sampling: 1.5 GB of DRAM traffic for 0.67 GB of compulsory I/O.  Here a 128-pixel tile
// (2 rows x 64 columns of the 2h x 128 map = 1 row x 32 columns of the h x 64 map) goes through the MLP-style
// TMEM chain: the three 1x1 GEMMs land in tensor memory, their bias+ReLU epilogue stores f_j once (the stride-2
// 3x3 convolutions need f0/f1, the MSFA encoder needs f2) and re-feeds the split values as the A operand of down_feat;
// up2(f2) is the same TMEM lane mapping evaluated at (y >> 1, x >> 1), i.e. down2 is computed on the replicated
// rows (4x redundant, 8 MFLOP/img) and only the even/even lanes store f2.
//   TMEM   A1 ring [0,128)   : two slots of a split 32-channel input chunk (hi 32 | lo 32); 4 chunks per tile (o0, o1, x, x)
//          D1      [128,256) : two accumulators of the f_j GEMMs (blocks alternate; 3xTF32 corrections merged, K <= 64)
//          A2      [256,384) : split relu(D1 + b_j): two 32-channel chunks, the A operand of down_feat for block j
//          D2      [384,512) : down_feat accumulator main | corrections, summed over the three blocks (K = 192)
//   smem   W0, W1, W2 images resident (64 KB), Wf chunks through a 4-stage ring, two sets of input boxes (o0 | o1 | x)
//   warps  0-15 workers (lane quarter q = TMEM lanes = pixels, part p = 8 / 16 of the channels), 16 MMA issuer, 17 TMA
// ------------------------------------------------------------------------------------------------
// 16 consecutive channels of one pixel -> 32 contiguous bytes of a [.., 64] bf16 row
__device__ __forceinline__ void store_bf16x16(__nv_bfloat16* dst, const float (&acc)[16]) {
  uint4* po = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint32_t w4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[8 * q + 2 * i], acc[8 * q + 2 * i + 1]);
      w4[i] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    po[q] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
  }
}

struct DownFusedArgs {
  CUtensorMap tm_o0, tm_o1, tm_x;      // o0/o1 [B,32,2h,128] box 64 x 2 x 32; x [B,64,h,64] box 32 x 1 x 64
  const float *w0img, *w1img, *w2img, *wfimg;
  const float *b0, *b1, *b2, *bf;
  float *f0, *f1, *f2, *fg;
  int B, h;                            // h = rows of x (the outs have 2h rows, 128 columns)
  int out_bf16;                        // 1: f0 / f1 / f2 are stored as [B,H,W,64] bf16
  int fg_bf16;                         // 1: feat_grid is stored as bf16 planes [B,64,2h,128] (TPSPP_HEAD_FLAG_FEATGRID_BF16)
};
constexpr int DF_WARPS = 16;
constexpr int DF_THREADS = (DF_WARPS + 2) * 32;
constexpr int DF_WF_STAGES = 4;
constexpr int DF_IN_SET = 16384 + 16384 + 8192;                   // o0 box | o1 box | x box
constexpr int DF_SMEM = 4 * TS_STAGE + DF_WF_STAGES * TS_STAGE + 2 * DF_IN_SET + 2048 + 1024;

__global__ void __launch_bounds__(DF_THREADS, 1) down_fused_kernel(const __grid_constant__ DownFusedArgs g) {
  constexpr int NT = 64;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* wres = smem;                                       // W0 | W1 | W2 chunk 0 | W2 chunk 1, 16 KB each
  unsigned char* wfs = smem + 4 * TS_STAGE;                         // ring of down_feat weight chunks
  unsigned char* ins = wfs + DF_WF_STAGES * TS_STAGE;               // 2 sets x (o0 16 KB | o1 16 KB | x 8 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ins + 2 * DF_IN_SET);
  uint64_t* wres_full = bars;          // 1
  uint64_t* in_full = bars + 1;        // [2 sets][3 inputs]
  uint64_t* in_empty = bars + 7;       // [2][3]
  uint64_t* a1_full = bars + 13;       // [2]
  uint64_t* a1_empty = bars + 15;      // [2]
  uint64_t* d1_full = bars + 17;       // [2]
  uint64_t* d1_empty = bars + 19;      // [2]
  uint64_t* a2_full = bars + 21;
  uint64_t* a2_empty = bars + 22;
  uint64_t* d2_full = bars + 23;
  uint64_t* d2_empty = bars + 24;
  uint64_t* wf_full = bars + 25;       // [4]
  uint64_t* wf_empty = bars + 29;      // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 33);
  float* bias_s = reinterpret_cast<float*>(bars + 34);             // [4][64]: b0 | b1 | b2 | bf
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (tid == 0) {
    mbar_init(wres_full, 1);
    for (int i = 0; i < 6; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], DF_WARPS); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a1_full[i], DF_WARPS); mbar_init(&a1_empty[i], 1);
      mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], DF_WARPS);
    }
    mbar_init(a2_full, DF_WARPS); mbar_init(a2_empty, 1);
    mbar_init(d2_full, 1); mbar_init(d2_empty, DF_WARPS);
    for (int i = 0; i < DF_WF_STAGES; ++i) { mbar_init(&wf_full[i], 1); mbar_init(&wf_empty[i], 1); }
    fence_barrier_init();
  }
  if (tid < 256) {
    const float* bsrc = (tid >> 6) == 0 ? g.b0 : (tid >> 6) == 1 ? g.b1 : (tid >> 6) == 2 ? g.b2 : g.bf;
    bias_s[tid] = __ldg(bsrc + (tid & 63));
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  pdl_trigger();          // this CTA holds its tensor memory: the next grid of the chain may be scheduled behind it
  pdl_wait();             // ... and nothing below runs before the previous grid has completed (no-op in a plain launch)
  const int tiles_per_img = g.h * 2;
  const int ntiles = g.B * tiles_per_img;
  const int n_my = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  constexpr uint32_t IDESC = umma_instr_desc(TC_TM, NT, 2);
  constexpr uint32_t PART = NT * TC_KC * 4;                          // bytes of the hi (or lo) image of a chunk

  if (warp == DF_WARPS) {
    // ===== MMA issuer warp =====
    if (n_my > 0) mbar_wait_bounded(wres_full, 0);
    auto mma_chunk = [&](uint32_t d_main, uint32_t d_corr, uint32_t a_base, uint32_t b_base, bool first) {   // one elected lane
#pragma unroll
      for (int kk = 0; kk < TC_KC / 8; ++kk) {
        const uint64_t dbh = umma_smem_desc(b_base + kk * 2 * (NT * 16), NT * 16, 128);
        const uint64_t dbl = umma_smem_desc(b_base + PART + kk * 2 * (NT * 16), NT * 16, 128);
        const uint32_t a_hi = a_base + kk * 8, a_lo = a_hi + 32;
        umma_ts_tf32(d_main, a_hi, dbh, IDESC, (first && kk == 0) ? 0u : 1u);
        umma_ts_tf32(d_corr, a_lo, dbh, IDESC, (first && kk == 0 && d_corr != d_main) ? 0u : 1u);
        umma_ts_tf32(d_corr, a_hi, dbl, IDESC, 1u);
      }
    };
    // stage 1 of block j (0: o0, 1: o1, 2: x) of tile `it`: D1[gj & 1] = A1 chunk(s) . W_j^T
    auto issue_s1 = [&](int it, int j) {
      const int gj = it * 3 + j, db = gj & 1;
      if (gj >= 2) mbar_wait_bounded(&d1_empty[db], (uint32_t)(((gj >> 1) - 1) & 1));
      const int c0 = j, nc = j == 2 ? 2 : 1;                       // A1 chunks of the block: tile-local index c0 .. c0 + nc - 1
      for (int c = 0; c < nc; ++c) {
        const int gc = it * 4 + c0 + c, slot = gc & 1;
        mbar_wait_bounded(&a1_full[slot], (uint32_t)((gc >> 1) & 1));
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t d1 = tmem_d + 128 + (uint32_t)(db * 64);
          mma_chunk(d1, d1, tmem_d + (uint32_t)(slot * 64), smem_u32(wres) + (uint32_t)((c0 + c) * TS_STAGE), c == 0);
          umma_commit(&a1_empty[slot]);
          if (c == nc - 1) umma_commit(&d1_full[db]);
        }
        __syncwarp();
      }
    };
    if (n_my > 0) issue_s1(0, 0);
    for (int it = 0; it < n_my; ++it) {
      for (int j = 0; j < 3; ++j) {
        const int gj = it * 3 + j;
        if (j < 2) issue_s1(it, j + 1);                  // overlaps the epilogue of block j
        else if (it + 1 < n_my) issue_s1(it + 1, 0);     // next tile's first block overlaps this tile's last epilogues
        mbar_wait_bounded(a2_full, (uint32_t)(gj & 1));
        if (j == 0 && it >= 1) mbar_wait_bounded(d2_empty, (uint32_t)((it - 1) & 1));
        for (int c = 0; c < 2; ++c) {
          const int wi = gj * 2 + c, st = wi & (DF_WF_STAGES - 1);
          mbar_wait_bounded(&wf_full[st], (uint32_t)((wi / DF_WF_STAGES) & 1));
          tc_fence_after();
          if (elect_one_sync()) {
            mma_chunk(tmem_d + 384, tmem_d + 448, tmem_d + 256 + (uint32_t)(c * 64), smem_u32(wfs) + (uint32_t)(st * TS_STAGE),
                      j == 0 && c == 0);
            umma_commit(&wf_empty[st]);
            if (c == 1) { umma_commit(a2_empty); if (j == 2) umma_commit(d2_full); }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == DF_WARPS + 1) {
    // ===== TMA warp: resident weight images once; then input boxes (two sets ahead) and down_feat weight chunks =====
    if (n_my > 0 && elect_one_sync()) {
      mbar_arrive_expect_tx(wres_full, 4 * TS_STAGE);
      bulk_g2s(wres, g.w0img, TS_STAGE, wres_full, policy_evict_last());
      bulk_g2s(wres + TS_STAGE, g.w1img, TS_STAGE, wres_full, policy_evict_last());
      bulk_g2s(wres + 2 * TS_STAGE, g.w2img, 2 * TS_STAGE, wres_full, policy_evict_last());
    }
    __syncwarp();
    const int total_in = n_my * 3, total_w = n_my * 6;
    int ii = 0, wi = 0;
    while (ii < total_in || wi < total_w) {
      if (ii < total_in) {
        const int it = ii / 3, k = ii - it * 3, set = it & 1;
        int ok = 1;
        if (it >= 2) ok = mbar_test_wait(&in_empty[set * 3 + k], (uint32_t)(((it >> 1) - 1) & 1)) ? 1 : 0;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok) {
          if (elect_one_sync()) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int img = tile / tiles_per_img, rem = tile - img * tiles_per_img;
            const int r = rem >> 1, c0 = rem & 1;
            unsigned char* dst = ins + set * DF_IN_SET + (k == 0 ? 0 : k == 1 ? 16384 : 32768);
            uint64_t* bar = &in_full[set * 3 + k];
            if (k < 2) {
              mbar_arrive_expect_tx(bar, 16384);
              tma_load_4d(dst, k == 0 ? &g.tm_o0 : &g.tm_o1, 64 * c0, 2 * r, 0, img, bar, policy_evict_first());
            } else {
              mbar_arrive_expect_tx(bar, 8192);
              tma_load_4d(dst, &g.tm_x, 32 * c0, r, 0, img, bar, policy_evict_first());
            }
          }
          __syncwarp();
          ++ii;
        }
      }
      if (wi < total_w) {
        const int st = wi & (DF_WF_STAGES - 1);
        int ok = 1;
        if (wi >= DF_WF_STAGES) ok = mbar_test_wait(&wf_empty[st], (uint32_t)(((wi / DF_WF_STAGES) - 1) & 1)) ? 1 : 0;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok) {
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&wf_full[st], TS_STAGE);
            bulk_g2s(wfs + st * TS_STAGE, reinterpret_cast<const unsigned char*>(g.wfimg) + (size_t)(wi % 6) * TS_STAGE, TS_STAGE,
                     &wf_full[st], policy_evict_last());
          }
          __syncwarp();
          ++wi;
        }
      }
    }
  } else {
    // ===== worker warps: lane quarter q = warp & 3 (TMEM lanes 32q.. = tile pixels), channel part p = warp >> 2 =====
    const int q = warp & 3, p = warp >> 2;
    const int m = q * 32 + lane;                       // pixel of the tile: row m >> 6, column m & 63
    const int ty = m >> 6, tx = m & 63;
    const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16);
    const int H2 = 2 * g.h;
    constexpr int W2 = 128;
    // A1 chunk c (0: o0, 1: o1, 2/3: x channels 0-31 / 32-63) of tile `it`: this thread's 8 channels p*8 .. p*8+7
    auto produce_a1 = [&](int it, int c) {
      const int set = it & 1, k = c < 2 ? c : 2;
      const int gc = it * 4 + c, slot = gc & 1;
      mbar_wait_bounded(&in_full[set * 3 + k], (uint32_t)((it >> 1) & 1));
      const unsigned char* base = ins + set * DF_IN_SET + (k == 0 ? 0 : k == 1 ? 16384 : 32768);
      float v[8];
      if (c < 2) {
        const float* sp = reinterpret_cast<const float*>(base) + (p * 8) * 128 + m;            // [32 ch][2 rows][64]
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = sp[i * 128];
      } else {
        const float* sp = reinterpret_cast<const float*>(base) + ((c - 2) * 32 + p * 8) * 32 + (tx >> 1);   // [64 ch][32]
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = sp[i * 32];
      }
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        hi[i] = __float_as_uint(v[i]) & 0xFFFFE000u;
        lo[i] = __float_as_uint(v[i] - __uint_as_float(hi[i]));
      }
      if (gc >= 2) { mbar_wait_bounded(&a1_empty[slot], (uint32_t)(((gc >> 1) - 1) & 1)); tc_fence_after(); }
      tmem_st8(lane_addr + (uint32_t)(slot * 64 + p * 8), hi);
      tmem_st8(lane_addr + (uint32_t)(slot * 64 + 32 + p * 8), lo);
      if (c != 2) {                                     // the x box serves two chunks: released after the second
        __syncwarp();
        if (lane == 0) mbar_arrive(&in_empty[set * 3 + k]);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a1_full[slot]);
    };
    // block epilogue, critical part: D1 -> + b_j -> ReLU -> split -> A2 of down_feat.  The values stay in `acc`; their
    // global stores are issued afterwards (block_store), off the MMA -> epilogue -> MMA dependency chain.
    auto block_epilogue = [&](int it, int j, float (&acc)[16]) {
      const int gj = it * 3 + j, db = gj & 1;
      mbar_wait_bounded(&d1_full[db], (uint32_t)((gj >> 1) & 1));
      tc_fence_after();
      tmem_ld_cols<16>(lane_addr + (uint32_t)(128 + db * 64 + p * 16), acc);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d1_empty[db]);          // D1 is in registers: the block after next may overwrite it
      const float4* bq = reinterpret_cast<const float4*>(bias_s + j * 64 + p * 16);
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 b4 = bq[i >> 2];
        acc[i] = fmaxf(acc[i] + b4.x, 0.f); acc[i + 1] = fmaxf(acc[i + 1] + b4.y, 0.f);
        acc[i + 2] = fmaxf(acc[i + 2] + b4.z, 0.f); acc[i + 3] = fmaxf(acc[i + 3] + b4.w, 0.f);
      }
      float hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        hi[i] = __uint_as_float(__float_as_uint(acc[i]) & 0xFFFFE000u);
        lo[i] = acc[i] - hi[i];
      }
      if (gj >= 1) { mbar_wait_bounded(a2_empty, (uint32_t)((gj - 1) & 1)); tc_fence_after(); }
      const uint32_t col = (uint32_t)(256 + (p >> 1) * 64 + (p & 1) * 16);
      tmem_st16(lane_addr + col, hi);
      tmem_st16(lane_addr + col + 32, lo);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a2_full);
    };
    auto block_store = [&](int it, int j, const float (&acc)[16]) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int img = tile / tiles_per_img, rem = tile - img * tiles_per_img;
      const int r = rem >> 1, c0 = rem & 1;
      if (j < 2) {
        const size_t o = (((size_t)img * 64 + p * 16) * H2 + (2 * r + ty)) * W2 + 64 * c0 + tx;
        const size_t plane = (size_t)H2 * W2;
        if (g.out_bf16) {                                   // bf16 mode of the head: f0 / f1 / f2 are stored as [B,H,W,64] bf16
          const size_t on = (((size_t)img * H2 + (2 * r + ty)) * W2 + 64 * c0 + tx) * 64 + p * 16;
          store_bf16x16(reinterpret_cast<__nv_bfloat16*>(j == 0 ? g.f0 : g.f1) + on, acc);
        } else {
          float* po = (j == 0 ? g.f0 : g.f1) + o;
#pragma unroll
          for (int i = 0; i < 16; ++i) po[(size_t)i * plane] = acc[i];
        }
      } else if (ty == 0 && (tx & 1) == 0) {              // f2 lives at half resolution: one writer per 2x2 block
        const size_t o = (((size_t)img * 64 + p * 16) * g.h + r) * 64 + 32 * c0 + (tx >> 1);
        const size_t plane = (size_t)g.h * 64;
        if (g.out_bf16) {
          const size_t on = (((size_t)img * g.h + r) * 64 + 32 * c0 + (tx >> 1)) * 64 + p * 16;
          store_bf16x16(reinterpret_cast<__nv_bfloat16*>(g.f2) + on, acc);
        } else {
          float* po = g.f2 + o;
#pragma unroll
          for (int i = 0; i < 16; ++i) po[(size_t)i * plane] = acc[i];
        }
      }
    };
    if (n_my > 0) { produce_a1(0, 0); produce_a1(0, 1); }
    float e0[16];
    if (n_my > 0) block_epilogue(0, 0, e0);
    for (int it = 0; it < n_my; ++it) {
      // (block 0's critical part already ran: before the loop / before the previous tile's final stores)
      produce_a1(it, 2); produce_a1(it, 3);
      block_store(it, 0, e0);
      float e1[16];
      block_epilogue(it, 1, e1);
      if (it + 1 < n_my) { produce_a1(it + 1, 0); produce_a1(it + 1, 1); }
      block_store(it, 1, e1);
      block_epilogue(it, 2, e1);
      block_store(it, 2, e1);
      // ---- final epilogue: feat_grid = relu(D2 + bf).  D2 is pulled into registers and released, then the next tile's
      //      first block goes through its critical part before this tile's feat_grid stores are issued ----
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int img = tile / tiles_per_img, rem = tile - img * tiles_per_img;
      const int r = rem >> 1, c0 = rem & 1;
      mbar_wait_bounded(d2_full, (uint32_t)(it & 1));
      tc_fence_after();
      float acc[16], part[16];
      tmem_ld_cols<16>(lane_addr + (uint32_t)(384 + p * 16), acc);
      tmem_ld_cols<16>(lane_addr + (uint32_t)(448 + p * 16), part);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_empty);
      if (it + 1 < n_my) block_epilogue(it + 1, 0, e0);
      const float4* bq = reinterpret_cast<const float4*>(bias_s + 3 * 64 + p * 16);
      const size_t fo = (((size_t)img * 64 + p * 16) * H2 + (2 * r + ty)) * W2 + 64 * c0 + tx;
      const size_t plane = (size_t)H2 * W2;
      if (g.fg_bf16) {            // bf16 mode: feat_grid as bf16 planes (the warp's TPSPP_SRC0_BF16 input)
        __nv_bfloat16* pb = reinterpret_cast<__nv_bfloat16*>(g.fg) + fo;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = bq[i >> 2];
          pb[(size_t)i * plane] = __float2bfloat16_rn(fmaxf(acc[i] + (part[i] + b4.x), 0.f));
          pb[(size_t)(i + 1) * plane] = __float2bfloat16_rn(fmaxf(acc[i + 1] + (part[i + 1] + b4.y), 0.f));
          pb[(size_t)(i + 2) * plane] = __float2bfloat16_rn(fmaxf(acc[i + 2] + (part[i + 2] + b4.z), 0.f));
          pb[(size_t)(i + 3) * plane] = __float2bfloat16_rn(fmaxf(acc[i + 3] + (part[i + 3] + b4.w), 0.f));
        }
      } else {
      float* po = g.fg + fo;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 b4 = bq[i >> 2];
        po[(size_t)i * plane] = fmaxf(acc[i] + (part[i] + b4.x), 0.f);
        po[(size_t)(i + 1) * plane] = fmaxf(acc[i + 1] + (part[i + 1] + b4.y), 0.f);
        po[(size_t)(i + 2) * plane] = fmaxf(acc[i + 2] + (part[i + 2] + b4.z), 0.f);
        po[(size_t)(i + 3) * plane] = fmaxf(acc[i + 3] + (part[i + 3] + b4.w), 0.f);
      }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 512);
}

// ------------------------------------------------------------------------------------------------
// Attention score chain in ONE kernel (reference tps_pp.py:258-261,293-312):
//     t = feat_linear.0(de')  (64 -> 32)    f = feat_linear.1(t)  (32 -> 128)    s = tanh(64^-1/2 . f . p1[b]^T)  (128 -> F = 32)
// As three launches the [rows, 32] and [rows, 128] intermediates went through HBM (0.17 GB written, 0.17 GB read back).
// Here a 128-pixel tile (2 rows x 64 columns of one image) walks the three GEMMs through tensor memory:
//   TMEM   A1 [0,128)   : split de' tile, K = 64 (two chunks of hi 32 | lo 32), thread = pixel = TMEM lane
//          D1 [128,160) : feat_linear.0 accumulator (N = 32, 3xTF32 corrections merged, K = 64)
//          A2 [192,256) : split (D1 + b0): one chunk, the A operand of feat_linear.1
//          D2 [256,320) : feat_linear.1 accumulator of one 64-column block (two blocks per tile; corrections merged, K = 32)
//          A3 [320,448) : split (D2 + b1): two chunks, the A operand of the score GEMM for that block
//          D3 [448,512) : score accumulator main 32 | corrections 32, summed over the two blocks (K = 128)
//   smem   feat_linear.0/.1 weight images resident (48 KB), the per-image score operand p1[b] (32 KB) and the de' tile
//          (32 KB) double-buffered
// ------------------------------------------------------------------------------------------------
struct ScoreFusedArgs {
  CUtensorMap tm_de;                   // de' [B,64,h,64] box 64 x 2 x 32
  const float *w0img, *w1img, *p1img;  // images: fl0 (2 chunks x 8 KB), fl1 (2 blocks x 16 KB), p1 per image (32 KB)
  const float *b0, *b1;
  float* score;                        // [B, h*64, 32]
  int B, h;
  float scale;
};
constexpr int SF_WARPS = 16;
constexpr int SF_THREADS = (SF_WARPS + 2) * 32;
constexpr int SF_SMEM = 16384 + 32768 + 2 * 32768 + 2 * 32768 + 1024 + 1024;

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(SF_THREADS, 1) score_fused_kernel(const __grid_constant__ ScoreFusedArgs g) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* w0s = smem;                      // fl0: chunk c at c * 8 KB (hi 4 KB | lo 4 KB)
  unsigned char* w1s = smem + 16384;              // fl1: block b at b * 16 KB (hi 8 KB | lo 8 KB)
  unsigned char* p1s = w1s + 32768;               // 2 sets x 32 KB: chunk c at c * 8 KB (hi 4 KB | lo 4 KB)
  unsigned char* ins = p1s + 2 * 32768;           // 2 sets x (chunk 0 16 KB | chunk 1 16 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ins + 2 * 32768);
  uint64_t* wres_full = bars;
  uint64_t* in_full = bars + 1;      // [2]
  uint64_t* in_empty = bars + 3;     // [2]
  uint64_t* p1_full = bars + 5;      // [2]
  uint64_t* p1_empty = bars + 7;     // [2]
  uint64_t* a1_full = bars + 9;
  uint64_t* a1_empty = bars + 10;
  uint64_t* d1_full = bars + 11;
  uint64_t* d1_empty = bars + 12;
  uint64_t* a2_full = bars + 13;
  uint64_t* a2_empty = bars + 14;
  uint64_t* d2_full = bars + 15;
  uint64_t* d2_empty = bars + 16;
  uint64_t* a3_full = bars + 17;
  uint64_t* a3_empty = bars + 18;
  uint64_t* d3_full = bars + 19;
  uint64_t* d3_empty = bars + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  float* bias_s = reinterpret_cast<float*>(bars + 22);             // b0 [32] | b1 [128]
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (tid == 0) {
    mbar_init(wres_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], SF_WARPS);
      mbar_init(&p1_full[i], 1); mbar_init(&p1_empty[i], 1);
    }
    mbar_init(a1_full, SF_WARPS); mbar_init(a1_empty, 1);
    mbar_init(d1_full, 1); mbar_init(d1_empty, SF_WARPS);
    mbar_init(a2_full, SF_WARPS); mbar_init(a2_empty, 1);
    mbar_init(d2_full, 1); mbar_init(d2_empty, SF_WARPS);
    mbar_init(a3_full, SF_WARPS); mbar_init(a3_empty, 1);
    mbar_init(d3_full, 1); mbar_init(d3_empty, SF_WARPS);
    fence_barrier_init();
  }
  if (tid < 160) bias_s[tid] = tid < 32 ? __ldg(g.b0 + tid) : __ldg(g.b1 + tid - 32);
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  pdl_trigger();          // this CTA holds its tensor memory: the next grid of the chain may be scheduled behind it
  pdl_wait();             // ... and nothing below runs before the previous grid has completed (no-op in a plain launch)
  const int tiles_per_img = g.h / 2;
  const int ntiles = g.B * tiles_per_img;
  const int n_my = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  constexpr uint32_t IDESC32 = umma_instr_desc(TC_TM, 32, 2), IDESC64 = umma_instr_desc(TC_TM, 64, 2);

  if (warp == SF_WARPS) {
    // ===== MMA issuer warp =====
    if (n_my > 0) mbar_wait_bounded(wres_full, 0);
    // one 32-wide K chunk: NTW = columns of the weight image (its hi part is NTW * 128 bytes, k-groups NTW * 16 bytes apart)
    auto mma_chunk = [&](uint32_t idesc, int NTW, uint32_t d_main, uint32_t d_corr, uint32_t a_base, uint32_t b_base, bool first) {
#pragma unroll
      for (int kk = 0; kk < TC_KC / 8; ++kk) {
        const uint64_t dbh = umma_smem_desc(b_base + kk * 2 * (NTW * 16), NTW * 16, 128);
        const uint64_t dbl = umma_smem_desc(b_base + NTW * 128 + kk * 2 * (NTW * 16), NTW * 16, 128);
        const uint32_t a_hi = a_base + kk * 8, a_lo = a_hi + 32;
        umma_ts_tf32(d_main, a_hi, dbh, idesc, (first && kk == 0) ? 0u : 1u);
        umma_ts_tf32(d_corr, a_lo, dbh, idesc, (first && kk == 0 && d_corr != d_main) ? 0u : 1u);
        umma_ts_tf32(d_corr, a_hi, dbl, idesc, 1u);
      }
    };
    for (int it = 0; it < n_my; ++it) {
      const int set = it & 1;
      // feat_linear.0: D1 = A1 . W0^T
      mbar_wait_bounded(a1_full, (uint32_t)(it & 1));
      if (it >= 1) mbar_wait_bounded(d1_empty, (uint32_t)((it - 1) & 1));
      tc_fence_after();
      if (elect_one_sync()) {
        for (int c = 0; c < 2; ++c)
          mma_chunk(IDESC32, 32, tmem_d + 128, tmem_d + 128, tmem_d + (uint32_t)(c * 64), smem_u32(w0s) + (uint32_t)(c * 8192), c == 0);
        umma_commit(a1_empty);
        umma_commit(d1_full);
      }
      __syncwarp();
      mbar_wait_bounded(a2_full, (uint32_t)(it & 1));
      mbar_wait_bounded(&p1_full[set], (uint32_t)((it >> 1) & 1));
      for (int b = 0; b < 2; ++b) {
        const int gb = it * 2 + b;
        // feat_linear.1, column block b: D2 = A2 . W1_b^T
        if (gb >= 1) mbar_wait_bounded(d2_empty, (uint32_t)((gb - 1) & 1));
        tc_fence_after();
        if (elect_one_sync()) {
          mma_chunk(IDESC64, 64, tmem_d + 256, tmem_d + 256, tmem_d + 192, smem_u32(w1s) + (uint32_t)(b * 16384), true);
          umma_commit(d2_full);
          if (b == 1) umma_commit(a2_empty);
        }
        __syncwarp();
        // score GEMM over the block's 64 features: D3 (+)= A3 . p1[img]^T chunks 2b, 2b+1
        mbar_wait_bounded(a3_full, (uint32_t)(gb & 1));
        if (b == 0 && it >= 1) mbar_wait_bounded(d3_empty, (uint32_t)((it - 1) & 1));
        tc_fence_after();
        if (elect_one_sync()) {
          for (int c = 0; c < 2; ++c)
            mma_chunk(IDESC32, 32, tmem_d + 448, tmem_d + 480, tmem_d + 320 + (uint32_t)(c * 64),
                      smem_u32(p1s) + (uint32_t)(set * 32768 + (2 * b + c) * 8192), b == 0 && c == 0);
          umma_commit(a3_empty);
          if (b == 1) { umma_commit(d3_full); umma_commit(&p1_empty[set]); }
        }
        __syncwarp();
      }
    }
  } else if (warp == SF_WARPS + 1) {
    // ===== TMA warp =====
    if (n_my > 0 && elect_one_sync()) {
      mbar_arrive_expect_tx(wres_full, 16384 + 32768);
      bulk_g2s(w0s, g.w0img, 16384, wres_full, policy_evict_last());
      bulk_g2s(w1s, g.w1img, 32768, wres_full, policy_evict_last());
    }
    __syncwarp();
    for (int it = 0; it < n_my; ++it) {
      const int set = it & 1;
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int img = tile / tiles_per_img, rp = tile - img * tiles_per_img;
      if (it >= 2) {
        mbar_wait_bounded(&in_empty[set], (uint32_t)(((it >> 1) - 1) & 1));
        mbar_wait_bounded(&p1_empty[set], (uint32_t)(((it >> 1) - 1) & 1));
      }
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&in_full[set], 32768);
        tma_load_4d(ins + set * 32768, &g.tm_de, 0, 2 * rp, 0, img, &in_full[set], policy_evict_first());
        tma_load_4d(ins + set * 32768 + 16384, &g.tm_de, 0, 2 * rp, 32, img, &in_full[set], policy_evict_first());
        mbar_arrive_expect_tx(&p1_full[set], 32768);
        bulk_g2s(p1s + set * 32768, g.p1img + (size_t)img * 8192, 32768, &p1_full[set], policy_evict_last());
      }
      __syncwarp();
    }
  } else {
    // ===== worker warps: lane quarter q = warp & 3 (TMEM lanes = tile pixels), part p = warp >> 2 =====
    const int q = warp & 3, p = warp >> 2;
    const int m = q * 32 + lane;
    const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16);
    auto produce_a1 = [&](int it) {
      const int set = it & 1;
      mbar_wait_bounded(&in_full[set], (uint32_t)((it >> 1) & 1));
      if (it >= 1) { mbar_wait_bounded(a1_empty, (uint32_t)((it - 1) & 1)); tc_fence_after(); }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float* sp = reinterpret_cast<const float*>(ins + set * 32768 + c * 16384) + (p * 8) * 128 + m;   // [32 ch][2][64]
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float v = sp[i * 128];
          hi[i] = __float_as_uint(v) & 0xFFFFE000u;
          lo[i] = __float_as_uint(v - __uint_as_float(hi[i]));
        }
        tmem_st8(lane_addr + (uint32_t)(c * 64 + p * 8), hi);
        tmem_st8(lane_addr + (uint32_t)(c * 64 + 32 + p * 8), lo);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&in_empty[set]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a1_full);
    };
    if (n_my > 0) produce_a1(0);
    for (int it = 0; it < n_my; ++it) {
      // ---- feat_linear.0 epilogue: D1 + b0 -> split -> A2 (8 of the 32 columns per thread) ----
      {
        mbar_wait_bounded(d1_full, (uint32_t)(it & 1));
        tc_fence_after();
        float acc[8];
        tmem_ld8(lane_addr + (uint32_t)(128 + p * 8), acc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d1_empty);
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float v = acc[i] + bias_s[p * 8 + i];
          hi[i] = __float_as_uint(v) & 0xFFFFE000u;
          lo[i] = __float_as_uint(v - __uint_as_float(hi[i]));
        }
        if (it >= 1) { mbar_wait_bounded(a2_empty, (uint32_t)((it - 1) & 1)); tc_fence_after(); }
        tmem_st8(lane_addr + (uint32_t)(192 + p * 8), hi);
        tmem_st8(lane_addr + (uint32_t)(192 + 32 + p * 8), lo);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a2_full);
      }
      if (it + 1 < n_my) produce_a1(it + 1);              // overlaps this tile's feat_linear.1 / score GEMMs
      // ---- feat_linear.1 epilogues: D2 + b1 -> split -> A3, one 64-column block at a time ----
#pragma unroll 1
      for (int b = 0; b < 2; ++b) {
        const int gb = it * 2 + b;
        mbar_wait_bounded(d2_full, (uint32_t)(gb & 1));
        tc_fence_after();
        float acc[16];
        tmem_ld_cols<16>(lane_addr + (uint32_t)(256 + p * 16), acc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d2_empty);
        float hi[16], lo[16];
        const float4* bq = reinterpret_cast<const float4*>(bias_s + 32 + b * 64 + p * 16);
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = bq[i >> 2];
          const float v0 = acc[i] + b4.x, v1 = acc[i + 1] + b4.y, v2 = acc[i + 2] + b4.z, v3 = acc[i + 3] + b4.w;
          hi[i] = __uint_as_float(__float_as_uint(v0) & 0xFFFFE000u); lo[i] = v0 - hi[i];
          hi[i + 1] = __uint_as_float(__float_as_uint(v1) & 0xFFFFE000u); lo[i + 1] = v1 - hi[i + 1];
          hi[i + 2] = __uint_as_float(__float_as_uint(v2) & 0xFFFFE000u); lo[i + 2] = v2 - hi[i + 2];
          hi[i + 3] = __uint_as_float(__float_as_uint(v3) & 0xFFFFE000u); lo[i + 3] = v3 - hi[i + 3];
        }
        if (gb >= 1) { mbar_wait_bounded(a3_empty, (uint32_t)((gb - 1) & 1)); tc_fence_after(); }
        const uint32_t col = (uint32_t)(320 + (p >> 1) * 64 + (p & 1) * 16);
        tmem_st16(lane_addr + col, hi);
        tmem_st16(lane_addr + col + 32, lo);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a3_full);
      }
      // ---- score epilogue: tanh(scale * D3) -> pc_score[row, 8p .. 8p+7] ----
      {
        mbar_wait_bounded(d3_full, (uint32_t)(it & 1));
        tc_fence_after();
        float acc[8], part[8];
        tmem_ld8(lane_addr + (uint32_t)(448 + p * 8), acc);
        tmem_ld8(lane_addr + (uint32_t)(480 + p * 8), part);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d3_empty);
        const size_t row = (size_t)((int)blockIdx.x + it * (int)gridDim.x) * TC_TM + m;
        float4 r0, r1;
        r0.x = tanhf((acc[0] + part[0]) * g.scale); r0.y = tanhf((acc[1] + part[1]) * g.scale);
        r0.z = tanhf((acc[2] + part[2]) * g.scale); r0.w = tanhf((acc[3] + part[3]) * g.scale);
        r1.x = tanhf((acc[4] + part[4]) * g.scale); r1.y = tanhf((acc[5] + part[5]) * g.scale);
        r1.z = tanhf((acc[6] + part[6]) * g.scale); r1.w = tanhf((acc[7] + part[7]) * g.scale);
        float4* po = reinterpret_cast<float4*>(g.score + row * 32 + p * 8);
        po[0] = r0; po[1] = r1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 512);
}

// weight image: for output row n (column block n / NT) and k = tap*Ctot + cin:
//   out[((blk*nchunks + k/32)*2 + part) * (NT*32) + ((k/4)%8) * (NT*4) + (n%NT)*4 + k%4]
struct WPrepArgs {
  WPrepLayer L[WPREP_MAX_LAYERS];
  int n;
};
__global__ void __launch_bounds__(256) wprep_kernel(WPrepArgs a) {
  const WPrepLayer L = a.L[blockIdx.y];
  const int Ktot = L.Ctot * L.taps;
  const int Npad = ((L.N + L.NT - 1) / L.NT) * L.NT;      // rows N .. Npad-1 of the last column block are zero
  const int total = Npad * Ktot;
  const int nchunks = Ktot >> 5;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    const int n = i / Ktot, k = i - n * Ktot;
    const int tap = k / L.Ctot, cin = k - tap * L.Ctot;
    float w = 0.f;
    if (n < L.N) {
      // forward: row n = output channel, k = tap * Ctot + cin.  dgrad (data gradient as a convolution over the output gradient):
      // row n = INPUT channel dg_ci0 + n of the forward layer, contraction over its Ctot OUTPUT channels, taps mirrored
      w = L.dg_cin == 0 ? __ldg(L.w + (size_t)n * Ktot + cin * L.taps + tap)
                        : __ldg(L.w + ((size_t)cin * L.dg_cin + L.dg_ci0 + n) * L.taps + (L.taps - 1 - tap));
    }
    if (L.scale != nullptr && n < L.N) w *= __ldg(L.scale + n);         // BatchNorm folded into the convolution (stage.cu)
    const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    if (L.bf16 == CM_MIX) {   // per chunk: hi tf32 [8 k-groups][NT n][4] | w bf16 [4 k-groups][NT n][8] | lo bf16 (same)
      // weights are prepared once, so the tf32 part is rounded to nearest (|lo| <= 2^-11 |w|)
      const float whi = __uint_as_float((__float_as_uint(w) + 0x1000u) & 0xFFFFE000u);
      unsigned char* oc = reinterpret_cast<unsigned char*>(L.out) + (size_t)(k >> 5) * (size_t)(L.NT * 256);
      reinterpret_cast<float*>(oc)[((k >> 2) & 7) * (L.NT * 4) + n * 4 + (k & 3)] = whi;
      const int o16 = ((k >> 3) & 3) * (L.NT * 8) + n * 8 + (k & 7);
      reinterpret_cast<__nv_bfloat16*>(oc + mix_b_bf(L.NT))[o16] = __float2bfloat16_rn(w);
      reinterpret_cast<__nv_bfloat16*>(oc + mix_b_lo(L.NT))[o16] = __float2bfloat16_rn(w - whi);
      continue;
    }
    if (L.bf16) {   // [chunk][4 k-groups][64 n][8 bf16]
      __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(L.out);
      ob[(size_t)(k >> 5) * (64 * 32) + ((k >> 3) & 3) * (64 * 8) + n * 8 + (k & 7)] = __float2bfloat16_rn(w);
      continue;
    }
    const int blk = n / L.NT, nn = n - blk * L.NT;
    const int ch = k >> 5, kg = (k >> 2) & 7, j = k & 3;
    float* o = L.out + ((size_t)(blk * nchunks + ch) * 2) * (L.NT * 32) + kg * (L.NT * 4) + nn * 4 + j;
    o[0] = hi;
    o[L.NT * 32] = w - hi;
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

size_t conv_tc_wprep_floats(int Ctot, int KS, int N) { return (size_t)2 * N * Ctot * KS * KS; }   // N already padded to NT

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

template <int KS, bool NHWC, int NT>
static int launch_tc(const ConvTcArgs& t, dim3 grid, cudaStream_t st) {
  static thread_local int attr_dev = -1;
  int dev = 0;
  TPSPP_CHECK_CUDA(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<KS, NHWC, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          tc_smem_bytes(NT)));
    attr_dev = dev;
  }
  conv_tc_kernel<KS, NHWC, NT><<<grid, TC_THREADS, tc_smem_bytes(NT), st>>>(t);
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());
  return TPSPP_OK;
}

// row-major [R, K] source, row-major output, R % 128 == 0, and either K <= 64 (A stays in TMEM across the column
// blocks) or a single column block
static bool lin_tma_plan(const ConvArgs& a, int NT, LinTmaArgs* g) {
  const ConvSrc& sc = a.src[0];
  if (!sc.nhwc || a.src[1].C != 0 || a.src[2].C != 0 || !a.out_nhwc || sc.uh != 1 || sc.uw != 1) return false;
  if (a.sh != 1 || a.sw != 1 || a.pad != 0 || a.Ctot != sc.C || a.Ctot % TC_KC || a.Cout % NT) return false;
  if (a.Cout > 256 && a.Ctot <= 2 * TC_KC) return false;      // the all-blocks-per-CTA form keeps <= 256 bias values in shared memory
  const long long R = (long long)a.B * a.Ho * a.Wo;
  if (R % TC_TM || R > 0x7fffffffLL || sc.H * sc.W * (long long)a.B != R) return false;
  const int nchunks = a.Ctot / TC_KC, nblocks = a.Cout / NT;
  g->col_split = (nchunks > 2 && nblocks != 1) ? 1 : 0;      // K > 64 and several column blocks: one block per blockIdx.y
  if (g->col_split && (nblocks > 65535 || a.wimg_stride != 0)) return false;
  if (a.splitk > 1 && (nchunks % a.splitk || nchunks / a.splitk < 1 || !(g->col_split || nblocks == 1) || a.wimg_stride != 0 ||
                       a.bias != nullptr || a.skip != nullptr || a.act != CONV_ACT_NONE || a.splitk > 64))
    return false;
  g->rows_per_img = 0;
  if (a.wimg_stride != 0) {
    const long long per = (long long)a.Ho * a.Wo;
    if (per % TC_TM) return false;
    g->rows_per_img = (int)per;
  }
  if (((uintptr_t)sc.ptr & 15) || ((uintptr_t)a.out & 15) || (a.skip && ((uintptr_t)a.skip & 15))) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)a.Ctot, (cuuint64_t)R};
  const cuuint64_t strides[1] = {(cuuint64_t)a.Ctot * 4};
  const cuuint32_t box[2] = {(cuuint32_t)TC_KC, (cuuint32_t)TC_TM};
  const cuuint32_t es[2] = {1, 1};
  return tmap_cached(&g->tmap, 2, sc.ptr, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
}

// fused Mlp launcher; returns 1 when the fused kernel does not apply (caller falls back to fc1 / fc2 launches)
int run_mlp_fused(const float* v, const float* x1, const float* w1img, const float* b1, const float* w2img, const float* b2,
                  float* out, long long R, cudaStream_t st) {
  if (R % TC_TM || R > 0x7fffffffLL || (((uintptr_t)v | (uintptr_t)x1 | (uintptr_t)out | (uintptr_t)b2) & 15)) return 1;
  MlpFusedArgs g;
  const cuuint64_t dims[2] = {64, (cuuint64_t)R};
  const cuuint64_t strides[1] = {64 * 4};
  const cuuint32_t box[2] = {(cuuint32_t)TC_KC, (cuuint32_t)TC_TM};
  const cuuint32_t es[2] = {1, 1};
  if (!tmap_cached(&g.tmap, 2, v, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  g.w1img = w1img; g.w2img = w2img; g.b1 = b1; g.b2 = b2; g.skip = x1; g.out = out; g.R = R;
  static thread_local int mf_dev = -1;
  int dev = 0;
  TPSPP_CHECK_CUDA(cudaGetDevice(&dev));
  if (mf_dev != dev) {
    TPSPP_CHECK_CUDA(cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MF_SMEM));
    mf_dev = dev;
  }
  const long long ntiles = R / TC_TM;
  dim3 grid((unsigned)min(ntiles, (long long)sm_count()));
  launch_k(mlp_fused_kernel, grid, dim3(MF_THREADS), MF_SMEM, st, g);
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());
  return TPSPP_OK;
}

// fused down0 + down1 + down2 + down_feat launcher; returns 1 when the fused kernel does not apply (caller runs the four
// separate convolutions instead: other widths, misaligned pointers)
int run_down_fused(const float* x, const float* o0, const float* o1, const float* w0img, const float* w1img, const float* w2img,
                   const float* wfimg, const float* b0, const float* b1, const float* b2, const float* bf, float* f0, float* f1,
                   float* f2, float* fg, int B, int h, int w, cudaStream_t st, int out_bf16, int fg_bf16) {
  if (w != 64 || h < 1 || B < 1) return 1;
  if ((((uintptr_t)x | (uintptr_t)o0 | (uintptr_t)o1 | (uintptr_t)w0img | (uintptr_t)w1img | (uintptr_t)w2img | (uintptr_t)wfimg) & 15) != 0)
    return 1;
  TPSPP_REQUIRE(tmap_encoder() != nullptr, "down_fused: cuTensorMapEncodeTiled is not available from this CUDA driver");
  DownFusedArgs g;
  {
    const cuuint64_t dims[4] = {128, (cuuint64_t)(2 * h), 32, (cuuint64_t)B};
    const cuuint64_t strides[3] = {128 * 4, (cuuint64_t)128 * 2 * h * 4, (cuuint64_t)128 * 2 * h * 32 * 4};
    const cuuint32_t box[4] = {64, 2, 32, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (!tmap_cached(&g.tm_o0, 4, o0, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
    if (!tmap_cached(&g.tm_o1, 4, o1, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  }
  {
    const cuuint64_t dims[4] = {64, (cuuint64_t)h, 64, (cuuint64_t)B};
    const cuuint64_t strides[3] = {64 * 4, (cuuint64_t)64 * h * 4, (cuuint64_t)64 * h * 64 * 4};
    const cuuint32_t box[4] = {32, 1, 64, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (!tmap_cached(&g.tm_x, 4, x, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  }
  g.w0img = w0img; g.w1img = w1img; g.w2img = w2img; g.wfimg = wfimg;
  g.b0 = b0; g.b1 = b1; g.b2 = b2; g.bf = bf;
  g.f0 = f0; g.f1 = f1; g.f2 = f2; g.fg = fg; g.B = B; g.h = h; g.out_bf16 = out_bf16; g.fg_bf16 = fg_bf16;
  static thread_local int df_dev = -1;
  int dev = 0;
  TPSPP_CHECK_CUDA(cudaGetDevice(&dev));
  if (df_dev != dev) {
    TPSPP_CHECK_CUDA(cudaFuncSetAttribute(down_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DF_SMEM));
    df_dev = dev;
  }
  const long long ntiles = (long long)B * h * 2;
  dim3 grid((unsigned)min(ntiles, (long long)sm_count()));
  launch_k(down_fused_kernel, grid, dim3(DF_THREADS), DF_SMEM, st, g);
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());
  return TPSPP_OK;
}

// fused feat_linear.0 -> feat_linear.1 -> tanh(QK^T / 8); returns 1 when not applicable (F != 32, odd geometry)
int run_score_fused(const float* de2, const float* w0img, const float* w1img, const float* p1img, const float* b0, const float* b1,
                    float* score, int B, int h, int w, int F, float scale, cudaStream_t st) {
  if (w != 64 || F != 32 || h < 2 || (h & 1) || B < 1) return 1;
  if ((((uintptr_t)de2 | (uintptr_t)w0img | (uintptr_t)w1img | (uintptr_t)p1img | (uintptr_t)score) & 15) != 0) return 1;
  TPSPP_REQUIRE(tmap_encoder() != nullptr, "score_fused: cuTensorMapEncodeTiled is not available from this CUDA driver");
  ScoreFusedArgs g;
  const cuuint64_t dims[4] = {64, (cuuint64_t)h, 64, (cuuint64_t)B};
  const cuuint64_t strides[3] = {64 * 4, (cuuint64_t)64 * h * 4, (cuuint64_t)64 * h * 64 * 4};
  const cuuint32_t box[4] = {64, 2, 32, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  if (!tmap_cached(&g.tm_de, 4, de2, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  g.w0img = w0img; g.w1img = w1img; g.p1img = p1img; g.b0 = b0; g.b1 = b1; g.score = score; g.B = B; g.h = h; g.scale = scale;
  static thread_local int sf_dev = -1;
  int dev = 0;
  TPSPP_CHECK_CUDA(cudaGetDevice(&dev));
  if (sf_dev != dev) {
    TPSPP_CHECK_CUDA(cudaFuncSetAttribute(score_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM));
    sf_dev = dev;
  }
  const long long ntiles = (long long)B * (h / 2);
  dim3 grid((unsigned)min(ntiles, (long long)sm_count()));
  launch_k(score_fused_kernel, grid, dim3(SF_THREADS), SF_SMEM, st, g);
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());
  return TPSPP_OK;
}

// NT: column tile (64, or 32 for narrow outputs); Cout / NT column blocks go to grid.y
int run_conv_tc(int KS, const ConvArgs& a, const float* wprep, int NT, cudaStream_t st, int mode) {
  const bool bf16 = mode == CM_BF16;
  TPSPP_REQUIRE(NT == 64 || NT == 32, "conv_tc: column tile must be 32 or 64");
  TPSPP_REQUIRE(a.Cout % NT == 0 || (NT == 64 && a.Cout == 32), "conv_tc: Cout %d is not a multiple of the column tile %d", a.Cout, NT);
  TPSPP_REQUIRE(KS == 1 || NT == 64 || (a.Cout == 32 && mode == CM_MIX && !a.out_nhwc),
                "conv_tc: 3x3 kernels with a 32-column tile exist for the 32-channel NCHW layers in the mixed operand mode only");
  // the TMA-staged kernels are the product path: without the driver's tensor-map encoder fail loudly instead of
  // silently dropping to the (2x slower) gather-fed kernels
  TPSPP_REQUIRE(tmap_encoder() != nullptr, "conv_tc: cuTensorMapEncodeTiled is not available from this CUDA driver");
  ConvTcArgs t;
  t.c = a;
  t.wprep = wprep;
  const long long M = (long long)a.B * a.Ho * a.Wo;
  dim3 grid((unsigned)((M + TC_TM - 1) / TC_TM), (unsigned)((a.Cout + NT - 1) / NT));
  const bool nhwc = a.src[0].nhwc != 0;
  if (!nhwc && a.wimg_stride == 0 && ((NT == 64 && (a.Cout == 64 || a.Cout == 32)) || (NT == 32 && a.Cout == 32 && !a.out_nhwc))) {
    // convolutions: A operand through TMEM
    int dev = 0;
    TPSPP_CHECK_CUDA(cudaGetDevice(&dev));
    ConvTmaArgs g;
    if (conv_tma_plan(KS, a, &g, mode)) {         // rectangular tiles: activations staged by TMA
      g.t = t;
      static thread_local int tma_dev = -1;
      if (tma_dev != dev) {
        TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_tma_kernel<1, CM_TF32X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tm_smem_bytes(1)));
        TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_tma_kernel<3, CM_TF32X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tm_smem_bytes(3)));
        TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_tma_kernel<1, CM_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, tm_smem_bytes(1)));
        TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_tma_kernel<3, CM_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, tm_smem_bytes(3)));
        TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_tma_kernel<1, CM_MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, tm_smem_bytes(1)));
        TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_tma_kernel<3, CM_MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, tm_smem_bytes(3)));
        TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_tma_kernel<1, CM_TF32X3, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, tm_smem_bytes(1)));
        TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_tma_kernel<3, CM_MIX, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, tm_smem_bytes(3)));
        tma_dev = dev;
      }
      dim3 pgrid((unsigned)min((long long)grid.x, 2LL * sm_count()));
      if (NT == 32) {       // 32-channel layers of the backbone stage: half the weight bytes and half the B-operand reads per chunk
        TPSPP_REQUIRE((KS == 1 && mode == CM_TF32X3) || (KS == 3 && mode == CM_MIX), "conv_tc: the 32-column tile exists for 1x1/3xTF32 and 3x3/mixed only");
        if (KS == 1) launch_k(conv_tma_kernel<1, CM_TF32X3, 32>, pgrid, dim3(TM_THREADS), tm_smem_bytes(1), st, g);
        else launch_k(conv_tma_kernel<3, CM_MIX, 32>, pgrid, dim3(TM_THREADS), tm_smem_bytes(3), st, g);
        count_launch();
        TPSPP_CHECK_CUDA(cudaGetLastError());
        return TPSPP_OK;
      }
      auto kern = KS == 1 ? (mode == CM_BF16 ? conv_tma_kernel<1, CM_BF16> : mode == CM_MIX ? conv_tma_kernel<1, CM_MIX> : conv_tma_kernel<1, CM_TF32X3>)
                          : (mode == CM_BF16 ? conv_tma_kernel<3, CM_BF16> : mode == CM_MIX ? conv_tma_kernel<3, CM_MIX> : conv_tma_kernel<3, CM_TF32X3>);
      launch_k(kern, pgrid, dim3(TM_THREADS), tm_smem_bytes(KS), st, g);
      count_launch();
      TPSPP_CHECK_CUDA(cudaGetLastError());
      return TPSPP_OK;
    }
    TPSPP_REQUIRE(NT == 64, "conv_tc: the 32-column convolution tile needs a TMA-stageable geometry");
    TPSPP_REQUIRE(!a.src[0].bf16 && !a.out_bf16 && !a.skip_bf16, "conv_tc: bf16-stored activations need the TMA-staged 3x3 bf16 kernel");
    static thread_local int ts_dev = -1;
    if (ts_dev != dev) {
      TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_ts_kernel<1, CM_TF32X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
      TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_ts_kernel<3, CM_TF32X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
      TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_ts_kernel<1, CM_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
      TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_ts_kernel<3, CM_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
      TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_ts_kernel<1, CM_MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
      TPSPP_CHECK_CUDA(cudaFuncSetAttribute(conv_ts_kernel<3, CM_MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
      ts_dev = dev;
    }
    auto kern = KS == 1 ? (mode == CM_BF16 ? conv_ts_kernel<1, CM_BF16> : mode == CM_MIX ? conv_ts_kernel<1, CM_MIX> : conv_ts_kernel<1, CM_TF32X3>)
                        : (mode == CM_BF16 ? conv_ts_kernel<3, CM_BF16> : mode == CM_MIX ? conv_ts_kernel<3, CM_MIX> : conv_ts_kernel<3, CM_TF32X3>);
    launch_k(kern, grid, dim3(TC_THREADS), TS_SMEM, st, t);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
    return TPSPP_OK;
  }
  TPSPP_REQUIRE(mode == CM_TF32X3, "conv_tc: the bf16 / mixed operand modes exist for the NCHW-source convolutions only");
  LinTmaArgs lg;
  const bool lin_ok = KS == 1 && nhwc && lin_tma_plan(a, NT, &lg);
  TPSPP_REQUIRE(a.splitk <= 1 || lin_ok, "conv_tc: split-K exists for the TMA-fed row-major linear kernel only");
  if (lin_ok) {
    lg.t = t;
    int dev = 0;
    TPSPP_CHECK_CUDA(cudaGetDevice(&dev));
    static thread_local int lin_dev = -1;
    if (lin_dev != dev) {
      TPSPP_CHECK_CUDA(cudaFuncSetAttribute(lin_tma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ln_smem_bytes()));
      TPSPP_CHECK_CUDA(cudaFuncSetAttribute(lin_tma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ln_smem_bytes()));
      lin_dev = dev;
    }
    dim3 pgrid((unsigned)min((long long)grid.x, 2LL * sm_count()), lg.col_split ? (unsigned)(a.Cout / NT) : 1u);
    if (lg.col_split) pgrid.x = (unsigned)min((long long)grid.x, max(1LL, 2LL * sm_count() / pgrid.y));
    if (a.splitk > 1) pgrid.z = (unsigned)a.splitk;
    if (NT == 64) launch_k(lin_tma_kernel<64>, pgrid, dim3(TC_THREADS), ln_smem_bytes(), st, lg);
    else launch_k(lin_tma_kernel<32>, pgrid, dim3(TC_THREADS), ln_smem_bytes(), st, lg);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
    return TPSPP_OK;
  }
  if (KS == 3) return nhwc ? launch_tc<3, true, 64>(t, grid, st) : launch_tc<3, false, 64>(t, grid, st);
  if (NT == 64) return nhwc ? launch_tc<1, true, 64>(t, grid, st) : launch_tc<1, false, 64>(t, grid, st);
  return nhwc ? launch_tc<1, true, 32>(t, grid, st) : launch_tc<1, false, 32>(t, grid, st);
}

}  // namespace tpspp
