// Probe: MN-major 16-bit (bf16) operands for tcgen05 kind::f16 — which (LBO, SBO, k-step, layout) does the hardware
// accept for a TMA SWIZZLE_128B panel image [K rows][64 elements along M/N = 128 B]? (What a 2-byte wgrad needs: both dy
// [pixel][co] and x [pixel][ci] are MN-major.)  D[m][n] = sum_k At[k][m] * Bt[k][n], K = 64, M = 128, N = 64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_mn16_probe.exe umma_mn16_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../../../multimodal-learning-with-alternating-unimodal-adaptation_b200/csrc/tc_common.cuh"

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                             float* out, uint32_t lbo, uint32_t sbo, uint32_t kstep, uint32_t layout) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar, done;
  __shared__ uint32_t slot;
  const uint32_t base = (tc::smem_u32(raw) + 1023u) & ~1023u;   // A: 2 panels x 8 KB; B: 1 panel 8 KB at +16 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { tc::mbar_init(tc::smem_u32(&bar), 1); tc::mbar_init(tc::smem_u32(&done), 1); tc::fence_mbar_init(); }
  if (warp == 0) { tc::tmem_alloc(tc::smem_u32(&slot), 64); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    tc::mbar_arrive_expect_tx(tc::smem_u32(&bar), 3 * 8192);
    tc::tma_load_2d(base, &mapA, tc::smem_u32(&bar), 0, 0);          // box {64 m, 64 k rows}
    tc::tma_load_2d(base + 8192, &mapA, tc::smem_u32(&bar), 64, 0);
    tc::tma_load_2d(base + 16384, &mapB, tc::smem_u32(&bar), 0, 0);
    tc::mbar_wait(tc::smem_u32(&bar), 0);
    tc::tc_fence_after();
    const uint32_t idesc = tc::make_idesc_f16(128, 64, 1, 1, 1, 1);   // bf16 x bf16, both MN-major
    for (int k = 0; k < 4; ++k) {                                     // K = 16 per instruction
      const uint64_t ad = tc::make_smem_desc(base + k * kstep, lbo, sbo, layout);
      const uint64_t bd = tc::make_smem_desc(base + 16384 + k * kstep, lbo, sbo, layout);
      tc::umma_f16(tmem, ad, bd, idesc, k != 0);
    }
    tc::umma_commit(tc::smem_u32(&done));
  }
  tc::mbar_wait(tc::smem_u32(&done), 0);
  tc::tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < 64; c += 32) {
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tc::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

int main() {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)f;
  const int K = 64, M = 128, N = 64;
  std::vector<__nv_bfloat16> hA(K * M), hB(K * N);
  std::vector<float> fA(K * M), fB(K * N), ref(M * N, 0.f);
  for (int k = 0; k < K; ++k) for (int m = 0; m < M; ++m) { fA[k * M + m] = (float)((k * 7 + m * 3) % 11 - 5); hA[k * M + m] = __float2bfloat16(fA[k * M + m]); }
  for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) { fB[k * N + n] = (float)((k * 5 + n * 2) % 7 - 3); hB[k * N + n] = __float2bfloat16(fB[k * N + n]); }
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += fA[k * M + m] * fB[k * N + n]; ref[m * N + n] = s; }
  __nv_bfloat16 *dA, *dB; float* dO;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, M * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap mA, mB;
  cuuint64_t dimsA[2] = {(cuuint64_t)M, (cuuint64_t)K}, dimsB[2] = {(cuuint64_t)N, (cuuint64_t)K};
  cuuint64_t strA[1] = {(cuuint64_t)M * 2}, strB[1] = {(cuuint64_t)N * 2};
  cuuint32_t box[2] = {64, 64}, es[2] = {1, 1};
  enc(&mA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, strA, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&mB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, strB, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> hO(M * N);
  const uint32_t lbos[] = {8192, 1024, 2048, 128, 16};
  const uint32_t sbos[] = {1024, 8192, 2048, 128, 512};
  const uint32_t ksteps[] = {2048, 1024, 256};
  for (uint32_t layout : {2u, 1u})
    for (uint32_t lbo : lbos) for (uint32_t sbo : sbos) for (uint32_t ks : ksteps) {
      cudaMemset(dO, 0, M * N * 4);
      probe<<<1, 128, 40 * 1024, 0>>>(mA, mB, dO, lbo, sbo, ks, layout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("layout %u lbo %u sbo %u kstep %u: CUDA error %s\n", layout, lbo, sbo, ks, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int i = 0; i < M * N; ++i) if (hO[i] != ref[i]) ++bad;
      if (bad == 0) printf("OK    layout %u lbo %5u sbo %5u kstep %4u\n", layout, lbo, sbo, ks);
      else if (bad < M * N / 2) printf("part  layout %u lbo %5u sbo %5u kstep %4u: %d mismatches\n", layout, lbo, sbo, ks, bad);
    }
  printf("done\n");
  return 0;
}
