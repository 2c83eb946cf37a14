// Halo-resident ("strip") implicit GEMM with RESIDENT WEIGHTS for the 64 -> 64 channel 3x3 / stride 1 / pad 1 convolutions of
// the ResNet-18 encoders (models/backbone.py:39-50: layer1, the largest activations of the network) — fprop16 and dgrad16,
// fp16 operands, fp32 accumulation in TMEM.
//
// Why: the im2col-TMA kernel (conv_gemm.cu) re-fetches every input pixel once per filter tap and the weight tile once per
// output tile: 216 KB of L2 -> SM traffic per 128 x 64 output tile, and ncu shows it running at 82-88 % of the chip's L2 -> SM
// cap with the tensor pipe 80 % idle (profiles/r2_conv_pair.md). Here
//   * the 9 x [64 x 64] fp16 filter taps (72 KB) are loaded ONCE per CTA and stay in shared memory (the CTA is persistent);
//   * ONE tiled 4-D TMA box per output tile brings the tile's input strip — TR+2 image rows x (W+2) pixels x 64 channels, the
//     zero padding filled by the TMA's out-of-bounds handling — into shared memory as pixel-linear 128-byte rows, and the 9
//     taps are 9 SHIFTED VIEWS of it: a K-major SWIZZLE_128B UMMA descriptor may start at any 128-byte row of a TMA-written
//     strip (the swizzle is a function of the absolute shared-memory address; tests/tools/probes/umma_offset_probe.cu).
// L2 -> SM traffic per tile: ~30 KB instead of 216 KB.
//
// GEMM rows enumerate the PADDED pixel grid of the tile: m = lr * Wp + wp, lr < TR, wp < Wp = W + 2. For tap (r, s) the operand
// row of m is strip row m + r * Wp + s. Rows with wp >= W (2 per image row) and rows past TR * Wp are garbage: computed,
// never stored, masked out of the BatchNorm partial sums. dgrad is the same walk over dy with the taps flipped and the
// transposed filter [Cin][R][S][Cout] (so both directions are the same K-major GEMM).
// Warp roles: warp 4 = TMA producer (weights once, then one strip per tile into a 3-deep ring), warp 5 = MMA issuer (36
// tcgen05.mma kind::f16 128 x 64 x 16 per tile into one of two TMEM accumulators), warps 0-3 = epilogue of the previous tile.
// Epilogue: the accumulator rows are compacted (padded columns dropped) into a 128B-swizzled shared-memory staging tile and
// leave through TWO TMA STORES per tile (32 channels x W pixels x TR rows each, alternating between two 16 KB staging
// halves; rows past the image are clipped by the TMA) —
// cp.reduce.async.bulk.tensor .add for the accumulating dgrad, so the `+=` happens in L2 and the SM never reads dx. (Per-
// thread 16-byte row stores were the bound of this kernel: 2.4 us per tile against 0.6 us of MMAs.)
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kThreads = 192;
constexpr int kC = 64;                       // channels in and out
constexpr int kStages = 3;                   // strips in flight (one strip = one tile: the TMA round trip of ~2 us needs >= 2 ahead)
constexpr uint32_t kHalfOut = 128 * 128;     // staging of one 32-channel half of a tile: 128 rows x 128 B (fp32)
constexpr uint32_t kTapBytes = kC * 128;     // one filter tap: 64 rows (output channels) x 64 fp16
constexpr uint32_t kWBytes = 9 * kTapBytes;  // 72 KB

struct Strip16Params {
  int N, H, W;        // image geometry shared by the gathered tensor and the output (stride 1, pad 1)
  int Wp, TR;         // padded row pitch W + 2; output rows per tile
  int tiles_per_img;  // ceil(H / TR)
  int tiles;          // N * tiles_per_img
  int flip;           // dgrad: filter tap (r, s) reads strip shift (2 - r, 2 - s)
  uint32_t a_stage;   // bytes per strip stage (multiple of 1024, >= 128 * (130 + 2 * Wp))
  uint32_t strip_tx;  // bytes of one strip box = 128 * Wp * (TR + 2)
  float* out;
  int accumulate;
  float* stat_part;   // fprop: [tile][2][64] BatchNorm partial sums, or NULL
  const float* out_scale;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m), "r"(src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1) conv_strip16_kernel(const __grid_constant__ CUtensorMap tmap_w,
                                                                    const __grid_constant__ CUtensorMap tmap_x,
                                                                    const __grid_constant__ CUtensorMap tmap_o,
                                                                    Strip16Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_bar, full_bar[kStages], empty_bar[kStages], acc_full_bar[2], acc_empty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_stat[4 * 2 * kC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base;                       // 9 taps x 8 KB
  const uint32_t a_base = base + kWBytes;             // strip ring
  const uint32_t o_base = a_base + kStages * p.a_stage;   // two 16 KB output staging halves

  if (threadIdx.x == 0) {
    tc::mbar_init(tc::smem_u32(&w_bar), 1);
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(tc::smem_u32(&full_bar[s]), 1);
      tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(tc::smem_u32(&acc_full_bar[b]), 1);
      tc::mbar_init(tc::smem_u32(&acc_empty_bar[b]), 4);
    }
    tc::fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_w);
    tc::tma_prefetch_desc(&tmap_x);
    tc::tma_prefetch_desc(&tmap_o);
  }
  if (warp == 5) {
    tc::tmem_alloc(tc::smem_u32(&tmem_slot), 2 * kC);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(tc::smem_u32(&w_bar), kWBytes);
      for (int tap = 0; tap < 9; ++tap)        // box {64 k, 64 rows}: filter tap `tap` of every output channel
        tc::tma_load_2d(w_base + tap * kTapBytes, &tmap_w, tc::smem_u32(&w_bar), tap * kC, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int s = it % kStages;
        const int n = tile / p.tiles_per_img, h0 = (tile - n * p.tiles_per_img) * p.TR;
        tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ((it / kStages) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(tc::smem_u32(&full_bar[s]), p.strip_tx);
        // box {64 ch, Wp pixels from w = -1, TR + 2 rows from h0 - 1, 1 image}: out-of-bounds = zero padding
        tc::tma_load_4d(a_base + s * p.a_stage, &tmap_x, tc::smem_u32(&full_bar[s]), 0, -1, h0 - 1, n);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_f16(128, kC, 0, 0, 0, 0);
      tc::mbar_wait(tc::smem_u32(&w_bar), 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int s = it % kStages, buf = it & 1;
        tc::mbar_wait(tc::smem_u32(&acc_empty_bar[buf]), ((it >> 1) & 1) ^ 1);
        tc::mbar_wait(tc::smem_u32(&full_bar[s]), (it / kStages) & 1);
        tc::tc_fence_after();
        const uint32_t strip = a_base + s * p.a_stage;
        const uint32_t acc = tmem_base + (uint32_t)buf * kC;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int r = tap / 3, c = tap - r * 3;
          const int shift = p.flip ? (2 - r) * p.Wp + (2 - c) : r * p.Wp + c;   // strip row of GEMM row 0 for this tap
          const uint32_t a0 = strip + (uint32_t)shift * 128u;
          const uint32_t b0 = w_base + tap * kTapBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = tc::make_smem_desc(a0 + k * 32u, 16u, 1024u, tc::kLayoutSw128);
            const uint64_t bd = tc::make_smem_desc(b0 + k * 32u, 16u, 1024u, tc::kLayoutSw128);
            tc::umma_f16(acc, ad, bd, idesc, (tap | k) != 0 ? 1u : 0u);
          }
        }
        tc::umma_commit(tc::smem_u32(&empty_bar[s]));        // the strip may be overwritten
        tc::umma_commit(tc::smem_u32(&acc_full_bar[buf]));   // the accumulator is complete
      }
    }
  } else {
    // ===================== epilogue (warps 0-3) =====================
    const float oscale = p.out_scale != nullptr ? __ldg(p.out_scale) : 1.f;
    const int m = warp * 32 + lane;
    const int lr = m / p.Wp, wp = m - lr * p.Wp;
    const bool in_box = lr < p.TR && wp < p.W;           // this thread's row exists in the compacted TR x W staging tile
    const int crow = lr * p.W + wp;                      // its row there
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int n = tile / p.tiles_per_img, h0 = (tile - n * p.tiles_per_img) * p.TR;
      const bool valid = in_box && h0 + lr < p.H;
      tc::mbar_wait(tc::smem_u32(&acc_full_bar[buf]), (it >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t acc = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * kC;
#pragma unroll 1
      for (int c = 0; c < kC; c += 32) {
        const uint32_t stg = o_base + (uint32_t)(c >> 5) * kHalfOut;     // channels [c, c + 32) of every tile use this half
        // its previous occupant (the same half of the previous tile = the commit group before the latest one) must have
        // been read by its TMA store
        if (threadIdx.x == 0) bulk_wait_read<1>();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint32_t v[32];
        tc::tmem_ld32(acc + c, v);
        tc::tmem_ld_wait();
        if (c + 32 == kC) {                                               // the accumulator has been read completely
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty_bar[buf]));
        }
        if (in_box) {
          // row `crow` of 128 B, 16-byte chunk j at position j ^ (crow & 7) (SWIZZLE_128B)
          const uint32_t rowp = stg + (uint32_t)crow * 128u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                   __uint_as_float(v[4 * j + 3]));
            if (p.out_scale != nullptr) { o.x *= oscale; o.y *= oscale; o.z *= oscale; o.w *= oscale; }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + tc::swz16(j, crow)), "f"(o.x), "f"(o.y),
                         "f"(o.z), "f"(o.w)
                         : "memory");
          }
        }
        if (p.stat_part != nullptr) {
          float a[32], b[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { a[j] = valid ? __uint_as_float(v[j]) : 0.f; b[j] = a[j] * a[j]; }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {   // butterfly transpose-reduce: lane j ends with column c + j
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              const float sa = up ? a[i] : a[i + off], ka = up ? a[i + off] : a[i];
              const float sb = up ? b[i] : b[i + off], kb2 = up ? b[i + off] : b[i];
              a[i] = ka + __shfl_xor_sync(0xffffffffu, sa, off);
              b[i] = kb2 + __shfl_xor_sync(0xffffffffu, sb, off);
            }
          }
          s_stat[(warp * 2 + 0) * kC + c + lane] = a[0];
          s_stat[(warp * 2 + 1) * kC + c + lane] = b[0];
        }
        tc::fence_proxy_async();                             // staging writes -> visible to the TMA (async proxy)
        asm volatile("bar.sync 1, 128;" ::: "memory");       // this half is staged (second half: s_stat complete too)
        if (threadIdx.x == 0) {
          if (p.accumulate) tma_reduce_add_4d(&tmap_o, stg, c, 0, h0, n);
          else tma_store_4d(&tmap_o, stg, c, 0, h0, n);
          bulk_commit();
        }
      }
      if (p.stat_part != nullptr) {
        for (int t = threadIdx.x; t < kC; t += 128) {
          float sa = 0.f, sb = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) { sa += s_stat[(w * 2 + 0) * kC + t]; sb += s_stat[(w * 2 + 1) * kC + t]; }
          float* dstp = p.stat_part + (size_t)tile * 2 * kC + t;
          dstp[0] = sa;
          dstp[kC] = sb;
        }
        // (s_stat is rewritten only after the next tile's first bar.sync)
      }
    }
    if (threadIdx.x == 0) bulk_wait_all();                 // the last stores have left shared memory and are complete
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tmem_base, 2 * kC);
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// fp16 [rows][cols] row-major, box {64 cols (128 B), box_rows}
int make_weight_map16(CUtensorMap* m, const void* ptr, long long rows, long long cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return MLA_E_NODEVICE;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MLA_E_BADARG;
}

// NHWC fp16 tensor as {C, W, H, N}; box {64 channels, Wp pixels, rows image rows, 1 image}
int make_strip_map16(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int Wp, int rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return MLA_E_NODEVICE;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)Wp, (cuuint32_t)rows, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MLA_E_BADARG;
}

// NHWC fp32 output {C, W, H, N}: box {32 channels (128 B), W pixels, rows image rows, 1 image}, 128B-swizzled staging
int make_out_map(CUtensorMap* m, float* ptr, int N, int H, int W, int C, int rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return MLA_E_NODEVICE;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {32u, (cuuint32_t)W, (cuuint32_t)rows, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MLA_E_BADARG;
}

}  // namespace

namespace mla {

// Geometry of the strip path, or tiles == 0 when it does not apply: 64 -> 64 channels, 3x3 / stride 1 / pad 1, W + 2 <= 64
// (at least two image rows per 128-row tile) and a row utilisation of at least 60 %. MLA_CONV_STRIP16=0 disables it.
StripPlan strip16_plan(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad) {
  StripPlan pl{};
  static const bool off = [] { const char* e = getenv("MLA_CONV_STRIP16"); return e != nullptr && e[0] == '0'; }();
  if (off || Cin != kC || Cout != kC || R != 3 || S != 3 || stride != 1 || pad != 1 || W + 2 > 64 || H < 1 || N < 1) return pl;
  const int Wp = W + 2;
  int TR = 128 / Wp;
  if (TR > H) TR = H;
  if (TR < 1) return pl;
  const int tpi = (H + TR - 1) / TR;
  const double util = (double)H * W / ((double)tpi * 128.0);
  if (util < 0.6) return pl;
  pl.Wp = Wp; pl.TR = TR; pl.tiles_per_img = tpi; pl.tiles = N * tpi;
  return pl;
}

// mode 0: fprop  (src = x16 [N,H,W,64] fp16, w16 = filter [64][3][3][64] fp16,            out = y  fp32)
// mode 1: dgrad  (src = dy16 fp16 (scaled), w16 = TRANSPOSED filter [Cin][3][3][Cout] fp16, out = dx fp32 (+)= *out_scale * ...)
int conv_strip16_run(int mode, const void* src16, const void* w16, float* out, int N, int H, int W, int accumulate,
                     float* stat_part, const float* out_scale, const StripPlan& pl, void* stream) {
  const DeviceInfo& di = device_info();
  if (di.ok != 1) return di.ok;
  Strip16Params p{};
  p.N = N; p.H = H; p.W = W; p.Wp = pl.Wp; p.TR = pl.TR; p.tiles_per_img = pl.tiles_per_img; p.tiles = pl.tiles;
  p.flip = mode == 1 ? 1 : 0;
  p.a_stage = (uint32_t)align_up((size_t)128 * (130 + 2 * pl.Wp), 1024);
  p.strip_tx = (uint32_t)(128 * pl.Wp * (pl.TR + 2));
  p.out = out; p.accumulate = accumulate; p.stat_part = stat_part; p.out_scale = out_scale;
  CUtensorMap wmap, xmap, omap;
  int rc = make_weight_map16(&wmap, w16, kC, 9LL * kC, kC);
  if (rc) return rc;
  rc = make_strip_map16(&xmap, src16, N, H, W, kC, pl.Wp, pl.TR + 2);
  if (rc) return rc;
  rc = make_out_map(&omap, out, N, H, W, kC, pl.TR);
  if (rc) return rc;
  const size_t smem = 1024 + kWBytes + (size_t)kStages * p.a_stage + 2 * (size_t)kHalfOut;
  static std::atomic<size_t> configured{0};
  if (smem > configured.load(std::memory_order_acquire)) {
    const size_t want = 1024 + kWBytes + (size_t)kStages * 33 * 1024 + 2 * (size_t)kHalfOut;   // Wp <= 64: a_stage <= 33 KB
    MLA_CUDA_TRY(cudaFuncSetAttribute(conv_strip16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
    configured.store(want, std::memory_order_release);
  }
  const int grid = std::min(pl.tiles, di.sm_count);
  conv_strip16_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(wmap, xmap, omap, p);
  MLA_CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace mla
