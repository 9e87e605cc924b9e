// Shared declarations of the head kernels (fp32 FFMA path in head.cu, tcgen05 path in head_tc.cu).
#pragma once
#include "common.cuh"

namespace tpspp {

struct ConvSrc {
  const float* ptr;
  int C, H, W, uh, uw;   // stored size and integer nearest-upsample factors (1 or 2)
  int nhwc;              // 0: [B,C,H,W] (the reference's layout, used at the boundary), 1: [B,H,W,C] (internal)
  int bf16;              // 1: the tensor is stored as bf16 (head-internal activations of the bf16 mode; 3x3 TMA kernel only)
};
struct ConvArgs {
  ConvSrc src[3];
  const float* weight;   // [64][Ctot][KS][KS]
  const float* bias;     // [64]
  const float* skip;     // [B,64,Ho,Wo] added after the ReLU, or null
  float* out;            // [B,64,Ho,Wo]
  int B, Ho, Wo, Ctot, sh, sw, pad;
  int out_nhwc;          // layout of out (and skip)
  // used by the tensor-core engine only (the FFMA kernel is 64-channel conv + bias + ReLU):
  int act;               // CONV_ACT_*
  float act_scale;       // tanh(scale * x)
  int Cout;              // channels of the output tensor (column blocks of NT go to grid.y)
  long long wimg_stride; // floats between per-image weight images (0: weights shared by all images)
  int skip_pre;          // 1: `skip` is added before the activation (residual block), 0: after it (MSFA decoder)
  int out_cstride;       // channels per image of the out / skip tensors when this launch writes a channel slice (0: = Cout)
  int zi;                // 1: up-sampled sources are ZERO-INSERTED instead of nearest (data gradient of a strided convolution)
  int out_bf16, skip_bf16; // 1: out / skip are stored as bf16 (bf16 mode of the head, 3x3 TMA kernel only)
  int splitk;              // row-major linear layers only (lin_tma_kernel): > 1 = the K chunks are split over blockIdx.z and `out` is a
                           // partial buffer [splitk][rows][Cout] (no bias / activation / skip: the caller reduces)
};
enum { CONV_ACT_RELU = 0, CONV_ACT_NONE = 1, CONV_ACT_GELU = 2, CONV_ACT_TANH = 3 };
// operand mode of a tensor-core convolution (head_tc.cu): 3xTF32 | single-pass bf16 | tf32 main term + bf16 corrections
enum { CM_TF32X3 = 0, CM_BF16 = 1, CM_MIX = 2 };


// tcgen05 implicit-GEMM convolution (head_tc.cu).  `wprep` is the layer's weight image produced by
// conv_tc_prepare_weights (3xTF32 hi/lo split, UMMA core-matrix order).
bool conv_tc_eligible(const ConvArgs& a, int KS);
int run_conv_tc(int KS, const ConvArgs& a, const float* wprep, int NT, cudaStream_t st, int mode = CM_TF32X3);
// DGAB Mlp fused (fc1 + GELU + fc2 + residual) on the weight images of fc1/fc2; returns 1 if not applicable
int run_mlp_fused(const float* v, const float* x1, const float* w1img, const float* b1, const float* w2img, const float* b2,
                  float* out, long long R, cudaStream_t st);
// down0 + down1 + down2 + down_feat in one kernel (tps_pp.py:538-540,548,560-562,581-585); returns 1 if not applicable
int run_down_fused(const float* x, const float* o0, const float* o1, const float* w0img, const float* w1img, const float* w2img,
                   const float* wfimg, const float* b0, const float* b1, const float* b2, const float* bf, float* f0, float* f1,
                   float* f2, float* fg, int B, int h, int w, cudaStream_t st, int out_bf16 = 0, int fg_bf16 = 0);
// feat_linear.0 -> feat_linear.1 -> tanh(QK^T / 8) in one kernel (tps_pp.py:258-261,293-312); returns 1 if not applicable
int run_score_fused(const float* de2, const float* w0img, const float* w1img, const float* p1img, const float* b0, const float* b1,
                    float* score, int B, int h, int w, int F, float scale, cudaStream_t st);
size_t conv_tc_wprep_floats(int Ctot, int KS, int N);    // floats needed for one layer's image (N output rows)

struct WPrepLayer {
  const float* w;   // [N][Ctot][KS*KS]
  float* out;
  int Ctot, taps, N, NT;
  const float* scale; // per-output-row factor applied before the split (BatchNorm folding), or null
  int dg_cin, dg_ci0; // dgrad image of a forward weight [Ctot out][dg_cin in][taps]: rows = input channels dg_ci0.. (dg_cin = 0: forward)
  int bf16;         // CM_*: 1 = single bf16 image [chunk][4 k-groups][64][8], 2 = tf32 hi | bf16 w | bf16 lo, 0 = fp32 hi|lo pair
};
constexpr int WPREP_MAX_LAYERS = 20;
// weight (+ bias) gradient of row-major linear layers on tcgen05 (conv_train.cu): gw[n][k] = sum_r gy[r][n] x[r][k]; batches > 1 =
// torch.bmm semantics (rows split evenly into groups with their own weight, gw [batches][N][K])
size_t wgrad_rows_ws_floats(long long R, int K, int N, int batches);
int run_wgrad_rows(const float* gy, const float* x, float* gw, float* gb, long long R, int K, int N, int batches, float* ws,
                   cudaStream_t st);
int conv_tc_prepare_weights(const WPrepLayer* layers, int nlayers, cudaStream_t st);

}  // namespace tpspp
