// Probe: does a TILED tensor map accept OVERLAPPING strides — dimension 0 = a 64-element (128-byte) window, dimension 1 =
// the window start advancing by 16 elements (32 bytes)? That is the sliding-window operand of the space-to-depth stem
// convolution (stem_s2d.cu): GEMM row = output pixel, K = 4 consecutive 16-channel s2d pixels of one filter row.
// Buffer: fp16 xs[N][Hs][Ws][16], value = a hash of its linear index. Box {64, 16 windows, 8 rows, 1 image}, SWIZZLE_128B.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_overlap_probe.exe tma_overlap_probe.cu
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../../multimodal-learning-with-alternating-unimodal-adaptation_b200/csrc/tc_common.cuh"

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, __half* out, int c1, int c2, int c3) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = (tc::smem_u32(raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    tc::mbar_init(tc::smem_u32(&bar), 1);
    tc::fence_mbar_init();
    tc::mbar_arrive_expect_tx(tc::smem_u32(&bar), 128 * 128);
    tc::tma_load_4d(base, &map, tc::smem_u32(&bar), 0, c1, c2, c3);
    tc::mbar_wait(tc::smem_u32(&bar), 0);
  }
  __syncthreads();
  const uint8_t* sm = raw + (base - tc::smem_u32(raw));
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int row = i / 64, e = i % 64;
    const int chunk = e / 8, within = e % 8;
    const uint32_t off = row * 128 + tc::swz16(chunk, row) + within * 2;
    out[i] = *reinterpret_cast<const __half*>(sm + off);
  }
}

int main() {
  const int N = 2, Hs = 11, Ws = 20;
  std::vector<__half> h((size_t)N * Hs * Ws * 16);
  for (size_t i = 0; i < h.size(); ++i) h[i] = __float2half((float)((i * 7 + 3) % 2039));
  __half *d, *o;
  cudaMalloc(&d, h.size() * 2);
  cudaMalloc(&o, 128 * 64 * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(f);
  CUtensorMap m;
  cuuint64_t dims[4] = {64, (cuuint64_t)(Ws - 3), (cuuint64_t)Hs, (cuuint64_t)N};
  cuuint64_t strides[3] = {32, (cuuint64_t)Ws * 32, (cuuint64_t)Hs * Ws * 32};
  cuuint32_t box[4] = {64, 16, 8, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode overlapping strides: %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  int bad_total = 0;
  const int cases[3][3] = {{0, 0, 0}, {3, 2, 1}, {8, 6, 1}};   // (window start, row start, image); the last one runs out of bounds
  for (auto& c : cases) {
    probe<<<1, 128, 34 * 1024>>>(m, o, c[0], c[1], c[2]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<__half> g(128 * 64);
    cudaMemcpy(g.data(), o, g.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int row = 0; row < 128; ++row)
      for (int e2 = 0; e2 < 64; ++e2) {
        const int j = row % 16, i = row / 16;
        const int w = c[0] + j, rr = c[1] + i;
        float want = 0.f;
        if (w < Ws - 3 && rr < Hs) want = __half2float(h[(((size_t)c[2] * Hs + rr) * Ws + w) * 16 + e2]);
        if (__half2float(g[row * 64 + e2]) != want) {
          if (bad < 4) printf("  case (%d,%d,%d) row %d e %d: got %g want %g\n", c[0], c[1], c[2], row, e2,
                              __half2float(g[row * 64 + e2]), want);
          ++bad;
        }
      }
    printf("case (%d,%d,%d): %d mismatches\n", c[0], c[1], c[2], bad);
    bad_total += bad;
  }
  printf(bad_total ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad_total != 0;
}
