// Single-query multi-head attention over a key/value cache: the per-step attention of `NRTRDecoder.forward_test`
// (reference decoders/nrtr_decoder.py:153-177 -> TFDecoderLayer -> MultiHeadAttention / ScaledDotProductAttention,
// common/modules/transformer_module.py:24-34,74-98) when the greedy decode keeps the keys / values of earlier positions instead
// of recomputing the whole prefix each step.
//   out[b, h, :] = sum_t softmax_t((q[b,h,:] / temperature) . K[b,t,h,:]) V[b,t,h,:],   t < len(b)
// q / out: [B, heads * 64] row-major; K / V: [B, capacity, heads * 64], or head-major [B, heads, capacity, 64]
// (cfg.kv_head_major) -- the self-attention cache, or the projected encoder memory for enc_attn; len(b) = kv_len, or kv_lens[b] (the reference's valid_ratio source mask, nrtr_decoder.py:111-123).
// One CTA per (image, head): four warps stride the keys in groups of four (lane = two of the 64 head dimensions,
// shuffle-reduced dot products, online softmax in fp32), their partial (max, sum, accumulator) triples are merged through
// shared memory.
#include "common.cuh"

namespace tpspp {

constexpr int AD_WARPS = 4, AD_DIM = 64;

template <bool STREAM>
__global__ void __launch_bounds__(AD_WARPS * 32) attn_decode_kernel(const float* __restrict__ q, float* __restrict__ k, float* __restrict__ v,
                                                                   float* __restrict__ out, const int* __restrict__ kv_lens, int kv_len,
                                                                   int capacity, int heads, float inv_temperature, int q_stride,
                                                                   const float* __restrict__ k_new, const float* __restrict__ v_new,
                                                                   int new_stride, int head_major) {
  pdl_trigger();          // no-ops in the plain launches the decode uses; they keep the kernel legal behind launch_k
  pdl_wait();
  const int b = blockIdx.x, h = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = heads * AD_DIM;
  int len = kv_lens != nullptr ? kv_lens[b] : kv_len;
  len = len < 0 ? 0 : (len > capacity ? capacity : len);
  if (k_new != nullptr && len > 0) {
    // append: this step's key / value rows (e.g. slices of a fused q|k|v projection) become cache position len - 1.  A CTA
    // touches only its own (image, head) slice, so the write needs no ordering against other CTAs.
    if (warp == 0) {
      const size_t dst = head_major ? (((size_t)b * heads + h) * capacity + (len - 1)) * AD_DIM + 2 * lane
                                    : ((size_t)b * capacity + (len - 1)) * D + h * AD_DIM + 2 * lane;
      const size_t src = (size_t)b * new_stride + h * AD_DIM + 2 * lane;
      *reinterpret_cast<float2*>(k + dst) = *reinterpret_cast<const float2*>(k_new + src);
      *reinterpret_cast<float2*>(v + dst) = *reinterpret_cast<const float2*>(v_new + src);
    }
    __syncthreads();
  }
  const float2 qv = *reinterpret_cast<const float2*>(q + (size_t)b * q_stride + h * AD_DIM + 2 * lane);
  const float q0 = qv.x * inv_temperature, q1 = qv.y * inv_temperature;       // the reference scales q, then multiplies
  // key t of this (image, head): head_major [B, heads, capacity, 64] -- one contiguous 256-byte row per key, 16 KB per CTA for
  // 64 keys (the [B, capacity, heads*64] layout interleaves the eight heads: 256-byte pieces 2 KB apart)
  const size_t tstride = head_major ? AD_DIM : D;
  const size_t base = head_major ? ((size_t)b * heads + h) * capacity * AD_DIM + 2 * lane : (size_t)b * capacity * D + h * AD_DIM + 2 * lane;
  const float* kb = k + base;
  const float* vb = v + base;
  float m = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f;
  // four keys per warp step: their key / value rows are loaded together and the four shuffle reductions interleave (one key
  // per step left a warp waiting on one 256-byte load and five dependent shuffles at a time: 16 us per call at B = 256)
  constexpr int U = 4;
  for (int t0 = warp * U; t0 < len; t0 += AD_WARPS * U) {
    float2 kv[U], vv[U];
    float s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u < len ? t0 + u : len - 1;                 // clamp: a valid row is loaded, its score is masked below
      // coherent loads (the cache may just have been appended to).  STREAM = this call alone reads more keys / values than the L2
      // holds (126 MB; the encoder memory at batch 1024: 268 MB per call): evict-first, so the dense layers' weight images stay
      // in L2 while it streams by (44.5 -> 43.3 ms per decode at batch 1024).  Below that size the hint costs time (batch 256,
      // 67 MB per call: 22.6 -> 23.8 ms), so smaller calls keep plain loads
      if (STREAM) {
        kv[u] = __ldcs(reinterpret_cast<const float2*>(kb + (size_t)t * tstride));
        vv[u] = __ldcs(reinterpret_cast<const float2*>(vb + (size_t)t * tstride));
      } else {
        kv[u] = *reinterpret_cast<const float2*>(kb + (size_t)t * tstride);
        vv[u] = *reinterpret_cast<const float2*>(vb + (size_t)t * tstride);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) s[u] = q0 * kv[u].x + q1 * kv[u].y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < U; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (t0 + u < len) {
        const float mn = fmaxf(m, s[u]);
        const float c = __expf(m - mn), p = __expf(s[u] - mn);          // m = -inf on the first key: c = 0
        l = l * c + p;
        a0 = a0 * c + p * vv[u].x;
        a1 = a1 * c + p * vv[u].y;
        m = mn;
      }
    }
  }
  __shared__ float sm[AD_WARPS], sl[AD_WARPS], sa[AD_WARPS][AD_DIM];
  if (lane == 0) { sm[warp] = m; sl[warp] = l; }
  sa[warp][2 * lane] = a0; sa[warp][2 * lane + 1] = a1;
  __syncthreads();
  if (warp == 0) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < AD_WARPS; ++w) M = fmaxf(M, sm[w]);
    float L = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int w = 0; w < AD_WARPS; ++w) {
      const float c = sm[w] == -INFINITY ? 0.f : __expf(sm[w] - M);
      L += sl[w] * c;
      o0 += sa[w][2 * lane] * c;
      o1 += sa[w][2 * lane + 1] * c;
    }
    const float r = L > 0.f ? 1.f / L : 0.f;                    // an empty key set (len = 0) yields zeros
    *reinterpret_cast<float2*>(out + (size_t)b * D + h * AD_DIM + 2 * lane) = make_float2(o0 * r, o1 * r);
  }
}

}  // namespace tpspp

using namespace tpspp;

extern "C" int tpspp_attn_decode(const tpspp_attn_cfg* cfg, const float* q, float* k, float* v, const int32_t* kv_lens,
                                 const float* k_new, const float* v_new, float* out, tpspp_stream_t stream) {
  reset_launch_count();
  TPSPP_REQUIRE(cfg != nullptr, "attn cfg is NULL");
  TPSPP_REQUIRE(cfg->batch >= 0 && cfg->heads > 0 && cfg->heads <= 65535, "attn: batch >= 0, 1 <= heads <= 65535");
  TPSPP_REQUIRE(cfg->head_dim == AD_DIM, "attn: head_dim must be 64 (d_k = d_v = 64 in the NRTR configs), got %d", cfg->head_dim);
  TPSPP_REQUIRE(cfg->kv_capacity > 0 && cfg->kv_len >= 0 && cfg->kv_len <= cfg->kv_capacity, "attn: 0 <= kv_len <= kv_capacity");
  TPSPP_REQUIRE(cfg->temperature > 0.f, "attn: temperature must be positive");
  if (cfg->batch == 0) return TPSPP_OK;
  TPSPP_REQUIRE(q && k && v && out, "tpspp_attn_decode: null pointer");
  TPSPP_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out | (uintptr_t)k_new | (uintptr_t)v_new) & 7) == 0,
                "tpspp_attn_decode: buffers must be 8-byte aligned");
  const int D = cfg->heads * AD_DIM;
  const int qs = cfg->q_stride > 0 ? cfg->q_stride : D, ns = cfg->new_stride > 0 ? cfg->new_stride : D;
  TPSPP_REQUIRE(qs >= D && ns >= D && qs % 2 == 0 && ns % 2 == 0, "tpspp_attn_decode: row strides must be even and >= heads * 64");
  TPSPP_REQUIRE((k_new == nullptr) == (v_new == nullptr), "tpspp_attn_decode: k_new and v_new come together");
  // (plain launch: inside the decode's CUDA graph programmatic edges bought nothing at batch 256 and cost 7 % at 1024)
  const bool stream_kv = (long long)cfg->batch * cfg->heads * cfg->kv_len * (2LL * AD_DIM * 4) > (126LL << 20);
  launch_k(stream_kv ? attn_decode_kernel<true> : attn_decode_kernel<false>, dim3(cfg->batch, cfg->heads), dim3(AD_WARPS * 32), 0, (cudaStream_t)stream,
           q, k, v, out, kv_lens, cfg->kv_len, cfg->kv_capacity, cfg->heads, 1.f / cfg->temperature, qs, k_new, v_new, ns,
           (int)(cfg->kv_head_major != 0));
  count_launch();
  TPSPP_CHECK_CUDA(cudaGetLastError());
  return TPSPP_OK;
}
