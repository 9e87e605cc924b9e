// Probe: which cp.async.bulk.tensor.4d box shapes / coordinates execute on this device (dev tool, not shipped).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, uint32_t bytes, float* out, int n) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(d), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  }
  __syncthreads();
  uint32_t ok = 0;
  for (int it = 0; it < (1 << 22) && !ok; ++it)
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(b) : "memory");
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = ok ? reinterpret_cast<float*>(sm)[i] : -777.f;
}
int main(int argc, char** argv) {
  typedef CUresult (*fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  fn_t enc = (fn_t)p;
  const int W = 64, H = 16, C = 64, B = 2;
  float* src; cudaMalloc(&src, sizeof(float) * W * H * C * B);
  float* h = (float*)malloc(sizeof(float) * W * H * C * B);
  for (int i = 0; i < W * H * C * B; ++i) h[i] = (float)(i % 1000) + 1.f;
  cudaMemcpy(src, h, sizeof(float) * W * H * C * B, cudaMemcpyHostToDevice);
  struct Case { int bw, bh, bc, x, y; } cases[] = {{72, 4, 32, -4, -1}, {68, 2, 32, -1, 0}, {68, 2, 32, 1, 0}, {68, 2, 32, 4, 0}, {72, 4, 32, 60, 13}, {40, 6, 32, -4, -1}, {72, 4, 32, -8, -1}};
  const int only = argc > 1 ? atoi(argv[1]) : -1; int idx = -1;
  for (auto cs : cases) {
    if (++idx != only && only >= 0) continue;
    CUtensorMap tm;
    cuuint64_t dims[4] = {W, H, C, B}, strides[3] = {W * 4, W * H * 4, (cuuint64_t)W * H * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)cs.bw, (cuuint32_t)cs.bh, (cuuint32_t)cs.bc, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int n = cs.bw * cs.bh * cs.bc;
    float* out; cudaMalloc(&out, n * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    probe<<<1, 128, n * 4 + 1024>>>(tm, cs.x, cs.y, 0, 1, n * 4, out, n);
    cudaError_t e = cudaDeviceSynchronize();
    float* ho = (float*)malloc(n * 4);
    cudaMemcpy(ho, out, n * 4, cudaMemcpyDeviceToHost);
    // expected value at box element (0,1,1): source (x, y+1, c=1, b=1)
    int bx = 1, by = 1, bc = 1;
    int sx = cs.x + bx, sy = cs.y + by;
    float exp = (sx < 0 || sy < 0 || sx >= W || sy >= H) ? 0.f : h[((1 * C + bc) * H + sy) * W + sx];
    printf("box %dx%dx%d at (%d,%d): encode=%d run=%s got=%g expect=%g\n", cs.bw, cs.bh, cs.bc, cs.x, cs.y, (int)r, cudaGetErrorString(e),
           ho[(bc * cs.bh + by) * cs.bw + bx], exp);
    if (e != cudaSuccess) { cudaDeviceReset(); cudaMalloc(&src, sizeof(float) * W * H * C * B); cudaMemcpy(src, h, sizeof(float) * W * H * C * B, cudaMemcpyHostToDevice); }
  }
  return 0;
}
