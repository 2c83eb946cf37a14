// Probe: can a tcgen05 K-major SWIZZLE_128B shared-memory descriptor start at an arbitrary 128-byte ROW of a
// TMA-written (1024-byte aligned) tile — i.e. is "rows d .. d+127 of a 256-row strip" a valid A operand?
// (What a halo-resident implicit-GEMM needs: the 9 taps of a 3x3 filter become 9 shifted views of one strip.)
// A[256][32] = row + 256*(col/4) (exact in TF32), B[64][32] = identity on k<32 -> D[m][n] must equal A[m+d][n].
// Tries base_offset = 0 and base_offset = (start >> 7) & 7 for each shift d.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_offset_probe umma_offset_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../../multimodal-learning-with-alternating-unimodal-adaptation_b200/csrc/tc_common.cuh"

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint64_t d = tc::make_smem_desc(saddr, lbo, sbo, layout);
  d |= (uint64_t)(base_off & 7) << 49;
  return d;
}

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                             float* out, int shift, int use_bo) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar, done;
  __shared__ uint32_t slot;
  const uint32_t base = (tc::smem_u32(raw) + 1023u) & ~1023u;   // A strip: 256 rows x 128 B = 32 KB; B after it
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::mbar_init(tc::smem_u32(&bar), 1);
    tc::mbar_init(tc::smem_u32(&done), 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) { tc::tmem_alloc(tc::smem_u32(&slot), 64); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    tc::mbar_arrive_expect_tx(tc::smem_u32(&bar), 256 * 128 + 64 * 128);
    tc::tma_load_2d(base, &mapA, tc::smem_u32(&bar), 0, 0);
    tc::tma_load_2d(base + 128 * 128, &mapA, tc::smem_u32(&bar), 0, 128);
    tc::tma_load_2d(base + 256 * 128, &mapB, tc::smem_u32(&bar), 0, 0);
    tc::mbar_wait(tc::smem_u32(&bar), 0);
    tc::tc_fence_after();
    constexpr uint32_t idesc = tc::make_idesc_tf32(128, 64, 0, 0);
    const uint32_t a0 = base + shift * 128;
    for (int k = 0; k < 4; ++k) {
      const uint32_t sa = a0 + k * 32;
      const uint64_t ad = desc_bo(sa, 16, 1024, tc::kLayoutSw128, use_bo ? (sa >> 7) & 7 : 0);
      const uint64_t bd = tc::make_smem_desc(base + 256 * 128 + k * 32, 16, 1024, tc::kLayoutSw128);
      tc::umma_tf32(tmem, ad, bd, idesc, k != 0);
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
  std::vector<float> hA(256 * 32), hB(64 * 32, 0.f);
  for (int r = 0; r < 256; ++r) for (int c = 0; c < 32; ++c) hA[r * 32 + c] = (float)(r + 256 * (c / 4));   // <= 2047: exact in TF32; identifies the row and the 16-byte chunk
  for (int n = 0; n < 32; ++n) hB[n * 32 + n] = 1.f;
  float *dA, *dB, *dO;
  cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap mA, mB;
  cuuint64_t dimsA[2] = {32, 256}, dimsB[2] = {32, 64}, str[1] = {128};
  cuuint32_t boxA[2] = {32, 128}, boxB[2] = {32, 64}, es[2] = {1, 1};
  enc(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dimsA, str, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dimsB, str, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> hO(128 * 64);
  const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 58, 59, 117, 118, 119};
  for (int use_bo = 0; use_bo < 2; ++use_bo)
    for (int shift : shifts) {
      cudaMemset(dO, 0, 128 * 64 * 4);
      probe<<<1, 128, 42 * 1024 + 1024, 0>>>(mA, mB, dO, shift, use_bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d bo %d: CUDA error %s\n", shift, use_bo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n)
        if (hO[m * 64 + n] != hA[(m + shift) * 32 + n]) { if (first < 0) first = m * 64 + n; ++bad; }
      printf("shift %3d base_offset %s: %s (%d mismatches%s)\n", shift, use_bo ? "(addr>>7)&7" : "0          ",
             bad ? "WRONG" : "ok", bad, bad ? "" : "");
      if (bad && first >= 0)
        printf("    first mismatch at m=%d n=%d: got %.0f want %.0f\n", first / 64, first % 64, hO[first],
               hA[(first / 64 + shift) * 32 + first % 64]);
    }
  return 0;
}
