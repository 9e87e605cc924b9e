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
//             gw[co][ci][t] = sum_{b,p} g[b,co,p] x[b,ci,p+t]  wgrad_kernel: warp-level tf32 MMAs with the 3xTF32 split (64 x 64 tile
//                                                             per tap and pixel split, deterministic two-stage reduction)
// Geometry: NCHW fp32, Cout = 64, Cin a multiple of 32, 1x1 (stride 1) or 3x3 (pad 1; stride 1, 2 or (2,1)).
#include "head.cuh"
#include "tc.cuh"

#include <stdlib.h>
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

// ---- weight gradient: one block = one (tap, 64-input-channel block, pixel split): a 64 x 64 tile gw[co][ci] += g[co][p] x[ci][p]
//      over its 32-pixel chunks, on the warp-level tensor-core path (mma.sync m16n8k8 tf32, fp32 accumulators in registers)
//      with the 3xTF32 split of both operands done on the fragments (hi*hi + lo*hi + hi*lo).  The first version (fp32 FMA outer
//      products, 4 x 4 register tiles) was bound by shared-memory loads: 5.6 ms of a 14.8 ms training step at batch 128.
//      Shared tiles are pixel-major [k][channel ^ f(k)] with f(k) = (k % 4) * 8 + k / 4: the staging stores of a warp (32 pixels
//      of one channel) and its fragment loads (8 channels x 4 pixels) both fall into 32 distinct banks.  tcgen05 is not used here on purpose: K = pixels would need the transposed A staging and its own
//      hand-off pipeline for 1 % of the training step's FLOP budget. ----
struct WgradArgs {
  const float *g, *x;          // g [B,64,Ho,Wo], x [B,Cin,H,W]
  float* part;                 // [splits][T * Cin * 64] as [tap][ci][co]
  int B, Cin, H, W, Ho, Wo, KS, sh, sw, splits;
  // conv mode, fused torch.cat / F.interpolate: nsrc > 1: 64-channel slice s of the input is its own tensor xs[s]; su/sv = log2 of
  // the slice's nearest-upsample factor (the tensor is stored [B, C, H >> su, W >> sv])
  const float* xs[3];
  int nsrc, su[3], sv[3];
  // row-major mode (wgrad_tc_kernel<true>, linear layers): g [R, N], x [R, K]; part [splits][K][Npad]; bpart [splits][2][Npad] or null
  long long R;
  int K, N, Npad;
  float* bpart;
};
constexpr int WG_PITCH = 64;
__device__ __forceinline__ int wg_swz(int k) { return ((k & 3) << 3) | (k >> 2); }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__global__ void __launch_bounds__(256, 3) wgrad_kernel(WgradArgs a) {
  __shared__ __align__(16) float Gs[32 * WG_PITCH];
  __shared__ __align__(16) float Xs[32 * WG_PITCH];
  const int T = a.KS * a.KS;
  const int tap = blockIdx.x % T, cib = blockIdx.x / T, sp = blockIdx.y;
  const int dy = tap / a.KS, dx = tap - dy * a.KS, pad = a.KS / 2;
  const int HoWo = a.Ho * a.Wo, HW = a.H * a.W;
  const long long nchunks = (long long)a.B * HoWo / 32;
  const long long c0 = nchunks * sp / a.splits, c1 = nchunks * (sp + 1) / a.splits;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cw = a.Cin - cib * 64 < 64 ? a.Cin - cib * 64 : 64;     // input channels of this block (32 for the 32-channel layers)
  // warp tile: 16 output channels x 32 input channels (four n-tiles of 8); fragment coordinates g = lane / 4, t = lane % 4
  const int co0 = (warp & 3) * 16, ci0 = (warp >> 2) * 32;
  const int fg = lane >> 2, ft = lane & 3;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // global loads of chunk ch + 1 are issued before the MMAs of chunk ch (registers), so their latency hides behind them
  float gr[8], xr[8];
  auto fetch = [&](long long ch) {
    const long long p0 = ch * 32;
    const int b = (int)(p0 / HoWo);
    const int p = (int)(p0 - (long long)b * HoWo) + lane;       // this lane's output pixel
    const int oy = p / a.Wo, ox = p - oy * a.Wo;
    const int iy = oy * a.sh + dy - pad, ix = ox * a.sw + dx - pad;
    const bool ok = iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
    const float* gp = a.g + (size_t)b * 64 * HoWo + p;
    const float* xp = a.x + ((size_t)b * a.Cin + cib * 64) * HW + (ok ? iy * a.W + ix : 0);
#pragma unroll
    for (int r = 0; r < 8; ++r) {                                // warp w stages channels w, w + 8, ...: lanes = pixels, coalesced
      const int c = warp + 8 * r;
      gr[r] = __ldg(gp + (size_t)c * HoWo);
      xr[r] = (ok && c < cw) ? __ldg(xp + (size_t)c * HW) : 0.f;
    }
  };
  if (c0 < c1) fetch(c0);
  for (long long ch = c0; ch < c1; ++ch) {
    __syncthreads();                                             // the previous chunk's tiles are no longer read
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int c = warp + 8 * r;
      Gs[lane * WG_PITCH + (c ^ wg_swz(lane))] = gr[r];
      Xs[lane * WG_PITCH + (c ^ wg_swz(lane))] = xr[r];
    }
    __syncthreads();
    if (ch + 1 < c1) fetch(ch + 1);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {                             // 8 pixels per mma
      const int k0 = ks * 8 + ft, k1 = k0 + 4;
      const int s0 = wg_swz(k0), s1 = wg_swz(k1);
      const float* g0 = Gs + k0 * WG_PITCH;
      const float* g1 = Gs + k1 * WG_PITCH;
      const float av[4] = {g0[(co0 + fg) ^ s0], g0[(co0 + fg + 8) ^ s0], g1[(co0 + fg) ^ s1], g1[(co0 + fg + 8) ^ s1]};   // (g,t) (g+8,t) (g,t+4) (g+8,t+4)
      uint32_t ah[4], al[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ah[i] = __float_as_uint(av[i]) & 0xFFFFE000u;
        al[i] = __float_as_uint(av[i] - __uint_as_float(ah[i]));
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int cc = ci0 + nt * 8 + fg;
        const float bv[2] = {Xs[k0 * WG_PITCH + (cc ^ s0)], Xs[k1 * WG_PITCH + (cc ^ s1)]};   // (k = t, n = g), (k = t + 4, n = g)
        uint32_t bh[2], bl[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          bh[i] = __float_as_uint(bv[i]) & 0xFFFFE000u;
          bl[i] = __float_as_uint(bv[i] - __uint_as_float(bh[i]));
        }
        mma_tf32(acc[nt], al, bh);
        mma_tf32(acc[nt], ah, bl);
        mma_tf32(acc[nt], ah, bh);
      }
    }
  }
  // accumulator fragment: c0 (row g, col 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1); rows = co, cols = ci
  float* o = a.part + ((size_t)sp * T + tap) * a.Cin * 64 + (size_t)(cib * 64) * 64;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int ci = ci0 + nt * 8 + 2 * ft, co = co0 + fg;
    if (ci < cw) { o[(size_t)ci * 64 + co] = acc[nt][0]; o[(size_t)ci * 64 + co + 8] = acc[nt][2]; }
    if (ci + 1 < cw) { o[(size_t)(ci + 1) * 64 + co] = acc[nt][1]; o[(size_t)(ci + 1) * 64 + co + 8] = acc[nt][3]; }
  }
}
// gw[co][ci][tap] = sum over splits of part[split][tap][ci][co]: threads walk the partials' layout (coalesced reads; the
// strided writes are 64 * Cin * T floats in total)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ gw, int Cin, int T, int splits) {
  // block = 32 outputs x 8 split lanes: with one thread per output the 1x1 layers (2048 outputs, 296 splits) were a serial chain of
  // 296 dependent loads on 8 blocks (40 us); fixed lane assignment + fixed tree = deterministic
  __shared__ float red[8][33];
  const int total = 64 * Cin * T;
  const int ox = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + ox;
  float a0 = 0.f, a1 = 0.f;
  if (i < total) {
    int s = sl;
    for (; s + 8 < splits; s += 16) { a0 += __ldg(part + (size_t)s * total + i); a1 += __ldg(part + (size_t)(s + 8) * total + i); }
    if (s < splits) a0 += __ldg(part + (size_t)s * total + i);
  }
  red[sl][ox] = a0 + a1;
  __syncthreads();
  if (sl == 0 && i < total) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += red[j][ox];
    const int tap = i / (Cin * 64), r = i - tap * (Cin * 64), ci = r >> 6, co = r & 63;
    gw[((size_t)co * Cin + ci) * T + tap] = acc;
  }
}

// ---- weight gradient on tcgen05 (the product path; the mma.sync kernel above is kept for geometries this one does not take).
//      gw[co][(tap, ci)] = sum over pixels: a GEMM whose contraction axis is the PIXEL axis, so the TS scheme of the forward
//      engine is used transposed: a CTA owns one pair of filter taps x 64 input channels = 128 A rows (TMEM lane = (tap, ci)),
//      a chunk's 32 pixels are the K columns, B = the 64 x 32 tile of the masked output gradient in the canonical shared-memory
//      layout.  3xTF32 on both operands (hi*hi into D_main, lo*hi + hi*lo into D_corr), accumulators persistent over the CTA's
//      pixel range, one epilogue at the end (partials per pixel split, reduced deterministically by wgrad_reduce_kernel).
//      Per chunk all 8 producer warps first stage x (two shifted rows per input channel) and g with coalesced loads
//      (lane = pixel) into padded shared tiles; then warps 0-3 read their A row (conflict-free, pitch 33), split it and store
//      it to tensor memory, warps 4-7 do the same for B into the operand images; one MMA warp issues. ----
constexpr int WT_THREADS = 288;
// raw-tile row pitch in floats.  Convolutions: 36 -- rows start 16-byte aligned, so g (and x of a 1x1 layer) arrive by 16-byte
// copies and are read back with LDS.128 (a quarter warp's rows fall into distinct bank groups: conflict-free).  Row-major
// linear layers: 33 -- the copies TRANSPOSE (lane = feature), which needs an odd pitch; reads are scalar.
// The x tile of a 3x3 layer keeps pitch 33 (its taps are misaligned: 4-byte copies, scalar reads; 36 would not fit two CTAs per SM);
// a 1x1 layer uses only 64 of the 128 rows, which leaves room for pitch 36 there.
template <bool ROWS> struct WtPitch { static constexpr int v = ROWS ? 33 : 36; };
constexpr int WT_XS = 128 * 33 * 4, WT_GS = 64 * 36 * 4;                          // raw staging tiles (bytes)
constexpr int WT_RING = 3;                                                         // raw-tile ring: chunk i+2 is in flight while chunk i is consumed
constexpr int WT_SMEM = 2 * 16384 /* B images */ + WT_RING * (WT_XS + WT_GS) + 64 + 1024;
// 4-byte asynchronous global -> shared copy; src_bytes = 0 writes a zero (padding, channels beyond the tensor)
__device__ __forceinline__ void cp_async4(void* dst, const void* src, bool ok) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(ok ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// ROWS = true: the same engine for a row-major linear layer y[R,N] = x[R,K] w^T: gw[n][k] = sum_r gy[r][n] x[r][k] -- the
// contraction axis is the ROW axis, A rows = 128 input features (blockIdx.x), B columns = 64 output features (blockIdx.z),
// a chunk = 32 rows; only the staging differs (lane = feature: coalesced 4-byte copies of a row's features into the
// transposed raw tiles).  The B-staging threads also keep the column sums of gy (bias gradient) when bpart is given.
template <bool ROWS>
__global__ void __launch_bounds__(WT_THREADS, 2) wgrad_tc_kernel(WgradArgs a, int npairs) {
  constexpr int WT_PITCH = WtPitch<ROWS>::v;              // g tile (and both tiles of the row-major mode)
  const int XP = (ROWS || a.KS != 1) ? 33 : 36;           // x tile
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* bimg = smem;                                             // 2 stages x (hi 8 KB | lo 8 KB)
  float* xs = reinterpret_cast<float*>(smem + 2 * 16384);                       // WT_RING x [128][pitch]
  float* gs = reinterpret_cast<float*>(smem + 2 * 16384 + WT_RING * WT_XS);     // WT_RING x [64][33]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * 16384 + WT_RING * (WT_XS + WT_GS));
  uint64_t* s_empty = bars;          // [2] the MMAs that read stage s completed
  uint64_t* s_full = bars + 2;       // [2] all 8 producer warps filled stage s (A in tensor memory, B images in shared memory)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) {
    mbar_init(&s_empty[0], 1); mbar_init(&s_empty[1], 1);
    mbar_init(&s_full[0], 8); mbar_init(&s_full[1], 8);
    fence_barrier_init();
  }
  if (!ROWS && a.KS == 1)                      // 1x1 layers copy only the valid A rows (16-byte path): the others stay zero
    for (int i = tid; i < WT_RING * (WT_XS / 4); i += WT_THREADS) xs[i] = 0.f;
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  const int T = a.KS * a.KS, pad = a.KS / 2;
  const int pair = ROWS ? 0 : blockIdx.x % npairs, cib = ROWS ? 0 : blockIdx.x / npairs, sp = blockIdx.y;
  const int cw = a.Cin - cib * 64 < 64 ? a.Cin - cib * 64 : 64;
  const int HoWo = a.Ho * a.Wo, HW = a.H * a.W;
  const long long nchunks = ROWS ? a.R / 32 : (long long)a.B * HoWo / 32;
  const long long c0 = nchunks * sp / a.splits, c1 = nchunks * (sp + 1) / a.splits;
  const int nmy = (int)(c1 - c0);
  constexpr uint32_t IDESC = umma_instr_desc(128, 64, 2);

  if (warp == 8) {
    // ===== MMA issuer =====
    for (int i = 0; i < nmy; ++i) {
      const int st = i & 1;
      mbar_wait_bounded(&s_full[st], (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t b_hi = smem_u32(bimg) + (uint32_t)(st * 16384), b_lo = b_hi + 8192;
        const uint32_t a_hi = tmem_d + 128 + (uint32_t)(st * 64), a_lo = a_hi + 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t dbh = umma_smem_desc(b_hi + j * 2 * (64 * 16), 64 * 16, 128);
          const uint64_t dbl = umma_smem_desc(b_lo + j * 2 * (64 * 16), 64 * 16, 128);
          umma_ts_tf32(tmem_d + 64, a_lo + j * 8, dbh, IDESC, (i | j) != 0 ? 1u : 0u);
          umma_ts_tf32(tmem_d + 64, a_hi + j * 8, dbl, IDESC, 1u);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t dbh = umma_smem_desc(b_hi + j * 2 * (64 * 16), 64 * 16, 128);
          umma_ts_tf32(tmem_d, a_hi + j * 8, dbh, IDESC, (i | j) != 0 ? 1u : 0u);
        }
        umma_commit(&s_empty[st]);
      }
      __syncwarp();
    }
  } else {
    // ===== producer warps =====
    // staging map: lane = pixel of the chunk; warp w loads A rows w, w + 8, ... (rows 0-63: first tap of the pair, 64-127: second)
    // and g rows w, w + 8, ...
    const int tap0 = 2 * pair, tap1 = 2 * pair + 1;
    const bool t1ok = tap1 < T;
    const int dy0 = tap0 / a.KS, dx0 = tap0 - dy0 * a.KS, dy1 = tap1 / a.KS, dx1 = tap1 - dy1 * a.KS;
    // asynchronous 4-byte copies (LDGSTS) straight into the raw-tile ring, two chunks ahead: a register prefetch of one chunk
    // left every chunk waiting for a full memory round trip (3300 cycles per chunk measured)
    auto fetch = [&](long long ch, int slot) {
      if (ROWS) {
        const long long r0 = ch * 32;
        const int k0 = blockIdx.x * 128, n0 = blockIdx.z * 64;
        float* xt = xs + slot * (WT_XS / 4);
        float* gt = gs + slot * (WT_GS / 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rl = warp + 8 * j;
          const float* xr = a.x + (size_t)(r0 + rl) * a.K;
          const float* gr = a.g + (size_t)(r0 + rl) * a.N;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = k0 + lane + 32 * q;
            cp_async4(xt + (lane + 32 * q) * WT_PITCH + rl, xr + (k < a.K ? k : 0), k < a.K);
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int n = n0 + lane + 32 * q;
            cp_async4(gt + (lane + 32 * q) * WT_PITCH + rl, gr + (n < a.N ? n : 0), n < a.N);
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        return;
      }
      const long long p0 = ch * 32;
      const int b = (int)(p0 / HoWo);
      const int p = (int)(p0 - (long long)b * HoWo) + lane;
      const int oy = p / a.Wo, ox = p - oy * a.Wo;
      const int iy0 = oy * a.sh + dy0 - pad, ix0 = ox * a.sw + dx0 - pad;
      const int iy1 = oy * a.sh + dy1 - pad, ix1 = ox * a.sw + dx1 - pad;
      const bool ok0 = iy0 >= 0 && iy0 < a.H && ix0 >= 0 && ix0 < a.W;
      const bool ok1 = t1ok && iy1 >= 0 && iy1 < a.H && ix1 >= 0 && ix1 < a.W;
      const int si = a.nsrc > 1 ? cib : 0, su = a.su[si], sv = a.sv[si], Ws = a.W >> sv, HWs = (a.H >> su) * Ws;
      const float* xb = a.nsrc > 1 ? a.xs[cib] + (size_t)b * 64 * HWs : a.x + ((size_t)b * a.Cin + cib * 64) * HWs;
      const float* x0 = xb + (ok0 ? (iy0 >> su) * Ws + (ix0 >> sv) : 0);
      const float* x1 = xb + (ok1 ? (iy1 >> su) * Ws + (ix1 >> sv) : 0);
      const float* gp = a.g + (size_t)b * 64 * HoWo + p;
      float* xt = xs + slot * (WT_XS / 4);
      float* gt = gs + slot * (WT_GS / 4);
      // g: the chunk's 32 pixels are contiguous and 16-byte aligned in every channel plane: 64 rows x 8 segments of 16 bytes
      const int ptid = warp * 32 + lane;
      const float* gch = gp - lane;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int idx = ptid + 256 * j, row = idx >> 3, seg = idx & 7;
        cp_async16(gt + row * WT_PITCH + seg * 4, gch + (size_t)row * HoWo + seg * 4);
      }
      if (a.KS == 1 && su == 0 && sv == 0) {           // 1x1, plain source: the same for the cw valid rows of x
        const float* xch = xb + (p - lane);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int idx = ptid + 256 * j, row = idx >> 3, seg = idx & 7;
          if (row < cw) cp_async16(xt + row * XP + seg * 4, xch + (size_t)row * HWs + seg * 4);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = warp + 8 * j;
          const bool cv = c < cw;
          cp_async4(xt + c * XP + lane, x0 + (size_t)(cv ? c : 0) * HWs, ok0 && cv);
          if (t1ok) cp_async4(xt + (64 + c) * XP + lane, x1 + (size_t)(cv ? c : 0) * HWs, ok1 && cv);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const uint32_t lane_addr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    float bsum = 0.f;                               // ROWS: column sum of gy over this CTA's rows (this thread's half of each chunk)
    if (nmy > 0) fetch(c0, 0);
    if (nmy > 1) fetch(c0 + 1, 1);
    for (int i = 0; i < nmy; ++i) {
      const int st = i & 1, slot = i % WT_RING;
      float* xt = xs + slot * (WT_XS / 4);
      float* gt = gs + slot * (WT_GS / 4);
      if (i + 1 < nmy) asm volatile("cp.async.wait_group 1;" ::: "memory");      // chunk i has landed (chunk i+1 may be in flight)
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      named_bar_sync(1, 256);                       // every producer thread's copies of chunk i are visible; chunk i-1 is consumed
      if (i + 2 < nmy) fetch(c0 + i + 2, (i + 2) % WT_RING);
      if (warp < 4) {
        // A: row = this thread's (tap, input channel), 32 pixels -> hi | lo columns of the stage in tensor memory
        const float* row = xt + (warp * 32 + lane) * XP;
        float v[32];
        if (ROWS || a.KS != 1) {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = row[k];
        } else if (warp < 2) {
#pragma unroll
          for (int k = 0; k < 8; ++k) *reinterpret_cast<float4*>(v + 4 * k) = *reinterpret_cast<const float4*>(row + 4 * k);
        } else {                 // rows 64-127 = the pair's second tap, which a 1x1 filter does not have
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = 0.f;
        }
        float hi[32], lo[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          hi[k] = __uint_as_float(__float_as_uint(v[k]) & 0xFFFFE000u);
          lo[k] = v[k] - hi[k];
        }
        if (i >= 2) {
          mbar_wait_bounded(&s_empty[st], (uint32_t)(((i >> 1) - 1) & 1));
          tc_fence_after();
        }
        const uint32_t col = lane_addr + 128u + (uint32_t)(st * 64);
        tmem_st16(col, hi); tmem_st16(col + 16, hi + 16);
        tmem_st16(col + 32, lo); tmem_st16(col + 48, lo + 16);
        tmem_st_wait();
        tc_fence_before();
      } else {
        // B: thread = (output channel, half of the chunk's pixels) -> canonical operand images [k-group of 4 pixels][co][16 B]
        const int t = tid - 128, co = t & 63, half = t >> 6;
        const float* row = gt + co * WT_PITCH + half * 16;
        float v[16];
        if (ROWS) {
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = row[k];
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(v + 4 * k) = *reinterpret_cast<const float4*>(row + 4 * k);
        }
        if (ROWS) {
#pragma unroll
          for (int k = 0; k < 16; ++k) bsum += v[k];
        }
        if (i >= 2) mbar_wait_bounded(&s_empty[st], (uint32_t)(((i >> 1) - 1) & 1));
        unsigned char* img = bimg + st * 16384;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 h, l;
          split_tf32(make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]), h, l);
          const int kg = half * 4 + q;
          *reinterpret_cast<float4*>(img + kg * 1024 + co * 16) = h;
          *reinterpret_cast<float4*>(img + 8192 + kg * 1024 + co * 16) = l;
        }
        fence_proxy_async();          // generic-proxy stores -> visible to the tensor core's async proxy
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_full[st]);
    }
    // ---- epilogue: D_main + D_corr -> partials [split][tap][ci][co] ----
    if (nmy > 0) {
      const int last = nmy - 1;
      mbar_wait_bounded(&s_empty[last & 1], (uint32_t)((last >> 1) & 1));
      tc_fence_after();
    }
    const int r = (warp & 3) * 32 + lane, half = warp >> 2;
    const int tap = 2 * pair + (r >> 6), ci = r & 63;
    const bool valid = ROWS ? ((int)blockIdx.x * 128 + r < a.K) : (tap < T && ci < cw);
    float* o = ROWS ? a.part + ((size_t)sp * a.K + (valid ? blockIdx.x * 128 + r : 0)) * a.Npad + blockIdx.z * 64 + half * 32
                    : a.part + (((size_t)sp * T + (valid ? tap : 0)) * a.Cin + cib * 64 + ci) * 64 + half * 32;
    if (ROWS && a.bpart != nullptr && blockIdx.x == 0 && warp >= 4) {
      const int t = tid - 128;                       // (column, half of the chunk) as in the B staging
      a.bpart[((size_t)sp * 2 + (t >> 6)) * a.Npad + blockIdx.z * 64 + (t & 63)] = bsum;
    }
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      float acc[16], part[16];
      if (nmy > 0) {
        tmem_ld_cols<16>(lane_addr + (uint32_t)(half * 32 + pass * 16), acc);
        tmem_ld_cols<16>(lane_addr + 64u + (uint32_t)(half * 32 + pass * 16), part);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) { acc[j] = 0.f; part[j] = 0.f; }
      }
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(o + pass * 16 + j) =
              make_float4(acc[j] + part[j], acc[j + 1] + part[j + 1], acc[j + 2] + part[j + 2], acc[j + 3] + part[j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 256);
}

// gw[n][k] = sum over splits of part[split][k][n];  gb[n] = sum over splits and halves of bpart (fixed order: deterministic)
__global__ void __launch_bounds__(256) wgrad_rows_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bpart,
                                                                float* __restrict__ gw, float* __restrict__ gb, int K, int N, int Npad,
                                                                int splits, int batched) {
  const int total = K * Npad;
  if (batched) {          // one split per weight batch (torch.bmm): gw[b][n][k] = part[b][k][n], nothing to sum
    for (long long i = blockIdx.x * 256 + threadIdx.x; i < (long long)total * splits; i += (long long)gridDim.x * 256) {
      const int b = (int)(i / total), r = (int)(i - (long long)b * total), k = r / Npad, n = r - k * Npad;
      if (n < N) gw[((size_t)b * N + n) * K + k] = __ldg(part + i);
    }
    return;
  }
  __shared__ float red[8][33];
  const int ox = threadIdx.x & 31, sl = threadIdx.x >> 5;
  for (int i0 = blockIdx.x * 32; i0 < total; i0 += gridDim.x * 32) {       // block = 32 outputs x 8 split lanes (see wgrad_reduce_kernel)
    const int i = i0 + ox;
    float a0 = 0.f;
    if (i < total)
      for (int s = sl; s < splits; s += 8) a0 += __ldg(part + (size_t)s * total + i);
    red[sl][ox] = a0;
    __syncthreads();
    if (sl == 0 && i < total) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += red[j][ox];
      const int k = i / Npad, n = i - k * Npad;
      if (n < N) gw[(size_t)n * K + k] = acc;
    }
    __syncthreads();
  }
  if (gb != nullptr && blockIdx.x * 32 < N) {
    const int n = blockIdx.x * 32 + ox;
    float a0 = 0.f;
    if (n < N)
      for (int s = sl; s < 2 * splits; s += 8) a0 += __ldg(bpart + (size_t)s * Npad + n);
    red[sl][ox] = a0;
    __syncthreads();
    if (sl == 0 && n < N) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += red[j][ox];
      gb[n] = acc;
    }
  }
}

int wgrad_rows_splits(long long R, int K, int N, int batches) {
  if (batches > 1) return batches;
  const long long blocks = (long long)((K + 127) / 128) * ((N + 63) / 64);
  long long sp = (2LL * sm_count()) / blocks;
  if (sp > R / 32 / 8) sp = R / 32 / 8;
  if (sp < 1) sp = 1;
  return (int)sp;
}
size_t wgrad_rows_ws_floats(long long R, int K, int N, int batches) {
  const int Npad = (N + 63) / 64 * 64;
  const int sp = wgrad_rows_splits(R, K, N, batches);
  return (size_t)sp * K * Npad + (size_t)sp * 2 * Npad;
}
// weight (and bias) gradient of a row-major linear layer on tcgen05; R % 32 == 0; ws >= wgrad_rows_ws_floats(R, K, N) floats
int run_wgrad_rows(const float* gy, const float* x, float* gw, float* gb, long long R, int K, int N, int batches, float* ws,
                   cudaStream_t st) {
  TPSPP_REQUIRE(R > 0 && R % 32 == 0, "wgrad_rows: rows must be a positive multiple of 32");
  TPSPP_REQUIRE(batches <= 1 || (R % batches == 0 && (R / batches) % 32 == 0 && gb == nullptr && batches <= 65535),
                "wgrad_rows: batched weights need rows per batch that are a multiple of 32 and no bias");
  const int Npad = (N + 63) / 64 * 64;
  const int sp = wgrad_rows_splits(R, K, N, batches);
  WgradArgs wa;
  memset(&wa, 0, sizeof(wa));
  wa.g = gy; wa.x = x; wa.part = ws; wa.KS = 1; wa.splits = sp; wa.R = R; wa.K = K; wa.N = N; wa.Npad = Npad;
  wa.Cin = 64; wa.Ho = wa.Wo = wa.H = wa.W = 1;
  wa.bpart = gb != nullptr ? ws + (size_t)sp * K * Npad : nullptr;
  static thread_local int wr_dev = -1;
  int dev = 0;
  TPSPP_CHECK_CUDA(cudaGetDevice(&dev));
  if (wr_dev != dev) {
    TPSPP_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM));
    wr_dev = dev;
  }
  wgrad_tc_kernel<true><<<dim3((K + 127) / 128, sp, Npad / 64), WT_THREADS, WT_SMEM, st>>>(wa, 1);
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());
  const long long rtotal = (long long)K * Npad * (batches > 1 ? sp : 1);
  wgrad_rows_reduce_kernel<<<(unsigned)min(batches > 1 ? (rtotal + 255) / 256 : (rtotal + 31) / 32, 4096LL), 256, 0, st>>>(ws, wa.bpart, gw, gb, K, N, Npad, sp, batches > 1);
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());
  return TPSPP_OK;
}

// data gradient of a nearest-upsampled source: gx[b][c][y][x] = sum over the uh x uw block of the full-resolution gradient
__global__ void __launch_bounds__(256) upsum_kernel(const float* __restrict__ gfull, float* __restrict__ gx, long long planes, int Hs,
                                                    int Ws, int uh, int uw) {
  const long long total = planes * Hs * Ws;
  const int Wf = Ws * uw;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long pl = i / (Hs * Ws);
    const int r = (int)(i - pl * (Hs * Ws)), y = r / Ws, x = r - y * Ws;
    const float* q = gfull + (pl * Hs * uh + (long long)y * uh) * Wf + (long long)x * uw;
    float acc = 0.f;
    for (int dy = 0; dy < uh; ++dy)
      for (int dx = 0; dx < uw; ++dx) acc += __ldg(q + (size_t)dy * Wf + dx);
    gx[i] = acc;
  }
}

struct ConvDims { int B, Cin, H, W, KS, sh, sw, Ho, Wo, relu, T, slices, splits, nsrc, uh[3], uw[3], anyup; };
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
  d->nsrc = c->nsrc < 1 ? 1 : c->nsrc;
  TPSPP_REQUIRE(d->nsrc <= 3 && (d->nsrc == 1 || d->Cin == 64 * d->nsrc), "nsrc must be 1..3 with 64 channels per concatenated source");
  d->anyup = 0;
  for (int s = 0; s < 3; ++s) {
    d->uh[s] = (s < d->nsrc && c->up_h[s] > 1) ? c->up_h[s] : 1;
    d->uw[s] = (s < d->nsrc && c->up_w[s] > 1) ? c->up_w[s] : 1;
    TPSPP_REQUIRE(d->uh[s] <= 2 && d->uw[s] <= 2 && d->H % d->uh[s] == 0 && d->W % d->uw[s] == 0, "upsample factors must be 1 or 2 and divide the input size");
    TPSPP_REQUIRE(((d->H / d->uh[s]) * (d->W / d->uw[s])) % 4 == 0, "source plane must be a multiple of 4 pixels");
    if (d->uh[s] > 1 || d->uw[s] > 1) d->anyup = 1;
  }
  TPSPP_REQUIRE(!d->anyup || (d->sh == 1 && d->sw == 1), "an upsampled source needs a stride-1 convolution");
  // pixel splits of the weight gradient: one resident wave of CTAs (two per SM), at least 8 chunks of 32 pixels per CTA
  long long chunks = (long long)d->B * d->Ho * d->Wo / 32;
  long long blocks_per_split = (long long)((d->T + 1) / 2) * d->slices;
  long long sp = (2LL * sm_count()) / blocks_per_split;
  if (sp > chunks / 8) sp = chunks / 8;
  if (sp < 1) sp = 1;
  if (sp > 296) sp = 296;
  d->splits = (int)sp;
  return TPSPP_OK;
}
enum { CW_WFWD = 0, CW_WDG, CW_G, CW_BPART, CW_WPART, CW_UPTMP, CW_COUNT };
static void conv_offsets(const ConvDims& d, size_t* off, size_t* total) {
  size_t sz[CW_COUNT];
  sz[CW_WFWD] = conv_tc_wprep_floats(d.Cin, d.KS, 64);
  sz[CW_WDG] = (size_t)d.slices * conv_tc_wprep_floats(64, d.KS, 64);
  sz[CW_G] = (size_t)d.B * 64 * d.Ho * d.Wo;
  sz[CW_BPART] = MB_SPLITS * 64;
  sz[CW_WPART] = (size_t)d.splits * d.T * d.Cin * 64;
  sz[CW_UPTMP] = d.anyup ? (size_t)d.B * (d.nsrc > 1 ? 64 : d.Cin) * d.H * d.W : 0;     // full-resolution data gradient of an upsampled source
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
  TPSPP_REQUIRE(cfg != nullptr && cfg->nsrc <= 1, "tpspp_conv_fwd: concatenated sources go through tpspp_convcat_fwd");
  const float* xs[3] = {x, nullptr, nullptr};
  return tpspp_convcat_fwd(cfg, xs, w, bias, y, workspace, stream);
}

extern "C" int tpspp_convcat_fwd(const tpspp_conv_cfg* cfg, const float* const* xs, const float* w, const float* bias, float* y,
                                 void* workspace, tpspp_stream_t stream) {
  reset_launch_count();
  ConvDims d;
  int rc = conv_dims(cfg, &d);
  if (rc != TPSPP_OK) return rc;
  if (d.B == 0) return TPSPP_OK;
  TPSPP_REQUIRE(xs && w && y && workspace, "tpspp_conv_fwd: null pointer");
  for (int s = 0; s < d.nsrc; ++s)
    TPSPP_REQUIRE(xs[s] != nullptr && ((uintptr_t)xs[s] & 15) == 0, "tpspp_conv_fwd: source %d is NULL or not 16-byte aligned", s);
  TPSPP_REQUIRE((((uintptr_t)y | (uintptr_t)workspace) & 15) == 0, "tpspp_conv_fwd: buffers must be 16-byte aligned");
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
  a.src[1].H = a.src[1].W = a.src[1].uh = a.src[1].uw = 1; a.src[2] = a.src[1];
  for (int s = 0; s < d.nsrc; ++s) {
    a.src[s].ptr = xs[s]; a.src[s].C = d.Cin / d.nsrc; a.src[s].H = d.H / d.uh[s]; a.src[s].W = d.W / d.uw[s];
    a.src[s].uh = d.uh[s]; a.src[s].uw = d.uw[s];
  }
  a.weight = w; a.bias = bias; a.out = y; a.B = d.B; a.Ho = d.Ho; a.Wo = d.Wo; a.Ctot = d.Cin; a.sh = d.sh; a.sw = d.sw;
  a.pad = d.KS / 2; a.act = d.relu ? CONV_ACT_RELU : CONV_ACT_NONE; a.act_scale = 1.f; a.Cout = 64;
  return run_conv_tc(d.KS, a, wimg, 64, st, mode);
}

extern "C" int tpspp_conv_bwd(const tpspp_conv_cfg* cfg, const float* x, const float* w, const float* y, const float* gy,
                              float* gx, float* gw, float* gb, void* workspace, tpspp_stream_t stream) {
  TPSPP_REQUIRE(cfg != nullptr && cfg->nsrc <= 1, "tpspp_conv_bwd: concatenated sources go through tpspp_convcat_bwd");
  const float* xs[3] = {x, nullptr, nullptr};
  float* gxs[3] = {gx, nullptr, nullptr};
  return tpspp_convcat_bwd(cfg, xs, w, y, gy, gxs, gw, gb, workspace, stream);
}

extern "C" int tpspp_convcat_bwd(const tpspp_conv_cfg* cfg, const float* const* xs, const float* w, const float* y, const float* gy,
                                 float* const* gxs, float* gw, float* gb, void* workspace, tpspp_stream_t stream) {
  reset_launch_count();
  ConvDims d;
  int rc = conv_dims(cfg, &d);
  if (rc != TPSPP_OK) return rc;
  if (d.B == 0) return TPSPP_OK;
  TPSPP_REQUIRE(xs && gxs && w && y && gy && workspace, "tpspp_conv_bwd: null pointer");
  for (int s = 0; s < d.nsrc; ++s)
    TPSPP_REQUIRE(xs[s] != nullptr && (((uintptr_t)xs[s] | (uintptr_t)gxs[s]) & 15) == 0, "tpspp_conv_bwd: source %d is NULL or misaligned", s);
  TPSPP_REQUIRE((((uintptr_t)y | (uintptr_t)gy | (uintptr_t)workspace) & 15) == 0, "tpspp_conv_bwd: buffers must be 16-byte aligned");
  const float* x = xs[0];
  bool any_gx = false;
  for (int s = 0; s < d.nsrc; ++s) any_gx = any_gx || gxs[s] != nullptr;
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
  if (any_gx) {
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
      const int si = d.nsrc > 1 ? s : 0;                       // the source this 64-channel slice belongs to
      float* gdst = gxs[si];
      if (gdst == nullptr) continue;
      const bool up = d.uh[si] > 1 || d.uw[si] > 1;
      // an upsampled source: full-resolution gradient into the scratch tensor, then the uh x uw block sums
      float* full = up ? W(CW_UPTMP) : gdst;
      ConvArgs a;
      memset(&a, 0, sizeof(a));
      a.src[0].ptr = W(CW_G); a.src[0].C = 64; a.src[0].H = d.Ho; a.src[0].W = d.Wo; a.src[0].uh = d.sh; a.src[0].uw = d.sw;
      a.src[1].H = a.src[1].W = a.src[1].uh = a.src[1].uw = 1; a.src[2] = a.src[1];
      a.zi = (d.sh == 2 || d.sw == 2) ? 1 : 0;
      if (d.nsrc > 1) { a.out = full; a.out_cstride = 0; }
      else { a.out = full + (size_t)s * 64 * d.H * d.W; a.out_cstride = d.Cin; }
      a.Cout = n;
      a.B = d.B; a.Ho = d.H; a.Wo = d.W; a.Ctot = 64; a.sh = 1; a.sw = 1; a.pad = d.KS / 2;
      a.act = CONV_ACT_NONE; a.act_scale = 1.f;
      rc = run_conv_tc(d.KS, a, W(CW_WDG) + (size_t)s * per, 64, st, mode);
      if (rc != TPSPP_OK) return rc;
      if (up && (d.nsrc > 1 || s == d.slices - 1)) {
        const long long planes = (long long)d.B * (d.nsrc > 1 ? 64 : d.Cin);
        const int Hs = d.H / d.uh[si], Ws = d.W / d.uw[si];
        const long long tot = planes * Hs * Ws;
        upsum_kernel<<<(unsigned)min((tot + 255) / 256, 8192LL), 256, 0, st>>>(full, gdst, planes, Hs, Ws, d.uh[si], d.uw[si]);
        count_launch();
        TPSPP_CHECK_CUDA(cudaGetLastError());
      }
    }
  }
  // 3. weight gradient
  if (gw != nullptr) {
    WgradArgs wa;
    memset(&wa, 0, sizeof(wa));
    wa.g = W(CW_G); wa.x = x; wa.part = W(CW_WPART); wa.B = d.B; wa.Cin = d.Cin; wa.H = d.H; wa.W = d.W; wa.Ho = d.Ho; wa.Wo = d.Wo;
    wa.KS = d.KS; wa.sh = d.sh; wa.sw = d.sw; wa.splits = d.splits; wa.nsrc = d.nsrc;
    for (int s = 0; s < 3; ++s) { wa.xs[s] = s < d.nsrc ? xs[s] : nullptr; wa.su[s] = d.uh[s] == 2; wa.sv[s] = d.uw[s] == 2; }
    static const bool env_mma_sync = getenv("TPSPP_WGRAD_MMASYNC") != nullptr;
    const bool use_mma_sync = env_mma_sync && d.nsrc == 1 && !d.anyup;
    if (!use_mma_sync) {
      static thread_local int wt_dev = -1;
      int dev = 0;
      TPSPP_CHECK_CUDA(cudaGetDevice(&dev));
      if (wt_dev != dev) {
        TPSPP_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM));
        wt_dev = dev;
      }
      const int npairs = (d.T + 1) / 2;
      wgrad_tc_kernel<false><<<dim3(npairs * d.slices, d.splits), WT_THREADS, WT_SMEM, st>>>(wa, npairs);
    } else {
      wgrad_kernel<<<dim3(d.T * d.slices, d.splits), 256, 0, st>>>(wa);
    }
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
    wgrad_reduce_kernel<<<(64 * d.Cin * d.T + 31) / 32, 256, 0, st>>>(W(CW_WPART), gw, d.Cin, d.T, d.splits);
    count_launch();
    TPSPP_CHECK_CUDA(cudaGetLastError());
  }
  return TPSPP_OK;
}
