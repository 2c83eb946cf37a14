// Halo-resident ("strip") implicit GEMM for the 3x3 / stride 1 / pad 1 convolutions of the ResNet-18 encoders
// (models/backbone.py:39-50: every BasicBlock conv except the three stride-2 ones) — fprop and dgrad.
//
// The im2col-TMA kernel (conv_gemm.cu) reloads every input pixel once per filter tap: 9 x 128 pixel rows of
// 128 B per 32-channel chunk and tile, and the TMA engine's im2col row rate (~15.6 ps per 128-byte row, chip-wide,
// measured) is what bounds it. Here ONE tiled 4-D TMA box per chunk brings the tile's input strip — TR+2 image
// rows x (W+2) pixels, zero padding filled by the TMA's out-of-bounds handling — into shared memory as
// pixel-linear 128-byte rows, and the 9 taps are 9 SHIFTED VIEWS of it: a K-major SWIZZLE_128B UMMA descriptor
// may start at any 128-byte row of a TMA-written strip (the swizzle is a function of the absolute shared-memory
// address; verified on B200 by tests/tools/probes/umma_offset_probe.cu, profiles/r1_umma_row_offset_probe.txt).
//
// GEMM rows enumerate the PADDED pixel grid of the tile: m = lr * Wp + wp, lr < TR, wp < Wp = W + 2. For tap (r, s)
// the operand row of m is strip row m + r*Wp + s — one linear shift for the whole tile. Rows with wp >= W (2 per
// image row) and rows past TR*Wp are garbage: computed, never stored, masked out of the BatchNorm partial sums.
//   strip rows per chunk : (TR+2) * Wp   (e.g. 4 x 58 = 232 for 56x56)   instead of 9 x 128 = 1152
//   MMA row utilisation  : TR * W / 128  (87.5 % at 56x56 and 28x28, 76.6 % at 14x14)
// dgrad (stride 1) is the same walk over dy with the taps flipped. Warp roles and the TMEM epilogue are those of
// conv_gemm.cu; two mbarrier rings: strips (one per chunk, 9 k-blocks each) and weight tiles (one per k-block).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kThreads = 192;

struct StripParams {
  int N, H, W;        // image geometry shared by the gathered tensor and the output (stride 1, pad 1)
  int Wp, TR;         // padded row pitch W + 2; output rows per tile
  int tiles_per_img;  // ceil(H / TR)
  int kcb;            // 32-channel chunks of the gathered tensor (Cin for fprop, Cout for dgrad)
  int CinW;           // Cin of the weight tensor (column stride of a tap in its 2-D view)
  int flip;           // dgrad: filter tap (r, s) reads strip shift (2-r, 2-s)
  uint32_t a_stage;   // bytes per strip stage (multiple of 1024, >= 128 * (130 + 2*Wp))
  uint32_t strip_tx;  // bytes of one strip box = 128 * Wp * (TR + 2)
  float* out;
  long long ldo;
  int accumulate;
  int Cout;           // channels of `out` (stride of the BN partial-sum rows)
  float* stat_part;   // fprop: [tile][2][Cout] BatchNorm partial sums, or NULL
  int dbg;            // timing experiments only (MLA_STRIP_DBG): 1 = loads without MMAs, 2 = MMAs without loads
};

template <int MODE, int BN, int SA, int SB>
__global__ void __launch_bounds__(kThreads) conv_strip_kernel(const __grid_constant__ CUtensorMap tmap_w,
                                                              const __grid_constant__ CUtensorMap tmap_x, StripParams p) {
  constexpr uint32_t kBBytes = BN * 128;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_a[SA], empty_a[SA], full_b[SB], empty_b[SB], tmem_full_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = base + SA * p.a_stage;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { tc::mbar_init(tc::smem_u32(&full_a[s]), 1); tc::mbar_init(tc::smem_u32(&empty_a[s]), 1); }
    for (int s = 0; s < SB; ++s) { tc::mbar_init(tc::smem_u32(&full_b[s]), 1); tc::mbar_init(tc::smem_u32(&empty_b[s]), 1); }
    tc::mbar_init(tc::smem_u32(&tmem_full_bar), 1);
    tc::fence_mbar_init();
  }
  if (warp == 4 && lane == 0) { tc::tma_prefetch_desc(&tmap_w); tc::tma_prefetch_desc(&tmap_x); }
  if (warp == 5) { tc::tmem_alloc(tc::smem_u32(&tmem_slot), BN); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  const int n = blockIdx.x / p.tiles_per_img;
  const int h0 = (blockIdx.x - n * p.tiles_per_img) * p.TR;
  const int n0 = blockIdx.y * BN;
  const int KB = p.kcb * 9;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int cb = 0; cb < (p.dbg >= 2 ? 0 : p.kcb); ++cb) {
        const int sa = cb % SA;
        tc::mbar_wait(tc::smem_u32(&empty_a[sa]), ((cb / SA) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(tc::smem_u32(&full_a[sa]), p.strip_tx);
        // box {32 ch, Wp pixels from w = -1, TR+2 rows from h0-1, 1 image}: out-of-bounds = zero padding
        tc::tma_load_4d(base + sa * p.a_stage, &tmap_x, tc::smem_u32(&full_a[sa]), cb * 32, -1, h0 - 1, n);
        for (int tap = 0; tap < 9; ++tap) {
          const int kb = cb * 9 + tap, sb = kb % SB;
          tc::mbar_wait(tc::smem_u32(&empty_b[sb]), ((kb / SB) & 1) ^ 1);
          const uint32_t bar = tc::smem_u32(&full_b[sb]);
          const uint32_t dst = b_base + sb * kBBytes;
          tc::mbar_arrive_expect_tx(bar, kBBytes);
          if (MODE == 0) {
            tc::tma_load_2d(dst, &tmap_w, bar, tap * p.CinW + cb * 32, n0);   // box {32 k, BN rows}, K-major
          } else {
#pragma unroll
            for (int pnl = 0; pnl < BN / 32; ++pnl)                          // box {32 ci, 32 co rows}, MN-major panels
              tc::tma_load_2d(dst + pnl * 4096, &tmap_w, bar, tap * p.CinW + n0 + pnl * 32, cb * 32);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_tf32(128, BN, 0, MODE == 1 ? 1 : 0);
      constexpr bool b_mn = (MODE == 1);
      constexpr uint32_t b_lbo = b_mn ? 4096u : 16u, b_sbo = b_mn ? 512u : 1024u;
      constexpr uint32_t b_lay = b_mn ? tc::kLayoutSw128Base32 : tc::kLayoutSw128;
      constexpr uint32_t b_kstep = b_mn ? 1024u : 32u;
      for (int cb = 0; cb < p.kcb; ++cb) {
        const int sa = cb % SA;
        if (p.dbg < 2) tc::mbar_wait(tc::smem_u32(&full_a[sa]), (cb / SA) & 1);
        const uint32_t strip = base + sa * p.a_stage;
        for (int tap = 0; tap < 9; ++tap) {
          const int kb = cb * 9 + tap, sb = kb % SB;
          if (p.dbg < 2) tc::mbar_wait(tc::smem_u32(&full_b[sb]), (kb / SB) & 1);
          tc::tc_fence_after();
          if (p.dbg == 1) { tc::mbar_arrive(tc::smem_u32(&empty_b[sb])); continue; }
          const int r = tap / 3, s = tap - r * 3;
          const int shift = p.flip ? (2 - r) * p.Wp + (2 - s) : r * p.Wp + s;   // strip row of GEMM row 0 for this tap
          const uint32_t a0 = strip + (uint32_t)shift * 128u;
          const uint32_t b0 = b_base + sb * kBBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = tc::make_smem_desc(a0 + k * 32, 16u, 1024u, tc::kLayoutSw128);
            const uint64_t bd = tc::make_smem_desc(b0 + k * b_kstep, b_lbo, b_sbo, b_lay);
            tc::umma_tf32(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc::umma_commit(tc::smem_u32(&empty_b[sb]));
        }
        if (p.dbg == 1) tc::mbar_arrive(tc::smem_u32(&empty_a[sa]));
        else tc::umma_commit(tc::smem_u32(&empty_a[sa]));
      }
      tc::umma_commit(tc::smem_u32(&tmem_full_bar));
    }
  } else {
    // ===================== epilogue (warps 0-3) =====================
    tc::mbar_wait(tc::smem_u32(&tmem_full_bar), 0);
    tc::tc_fence_after();
    const int m = warp * 32 + lane;
    const int lr = m / p.Wp, wp = m - lr * p.Wp;
    const bool valid = lr < p.TR && wp < p.W && h0 + lr < p.H;
    float* orow = valid ? p.out + (((long long)n * p.H + h0 + lr) * p.W + wp) * p.ldo + n0 : nullptr;
    // all MMAs have completed: the weight ring is free and serves as scratch for the per-warp column sums
    float* s_stat = reinterpret_cast<float*>(smem_raw + (b_base - tc::smem_u32(smem_raw)));   // [warp][2][BN]
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      tc::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
      tc::tmem_ld_wait();
      if (valid) {
        float4* dst = reinterpret_cast<float4*>(orow + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                 __uint_as_float(v[4 * j + 3]));
          if (p.accumulate) {
            const float4 old = dst[j];
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          dst[j] = o;
        }
      }
      if (MODE == 0 && p.stat_part != nullptr) {
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
        s_stat[(warp * 2 + 0) * BN + c + lane] = a[0];
        s_stat[(warp * 2 + 1) * BN + c + lane] = b[0];
      }
    }
    if (MODE == 0 && p.stat_part != nullptr) {
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the 4 epilogue warps only
      for (int t = threadIdx.x; t < BN; t += 128) {
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) { sa += s_stat[(w * 2 + 0) * BN + t]; sb += s_stat[(w * 2 + 1) * BN + t]; }
        float* dstp = p.stat_part + (size_t)blockIdx.x * 2 * p.Cout + n0 + t;
        dstp[0] = sa;
        dstp[p.Cout] = sb;
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tmem_base, BN);
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

int make_weight_map(CUtensorMap* m, const float* ptr, long long rows, long long cols, int box_rows, bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return MLA_E_NODEVICE;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MLA_E_BADARG;
}

// NHWC fp32 tensor as {C, W, H, N}; box {32 channels, Wp pixels, rows image rows, 1 image}
int make_strip_map(CUtensorMap* m, const float* ptr, int N, int H, int W, int C, int Wp, int rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return MLA_E_NODEVICE;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {32u, (cuuint32_t)Wp, (cuuint32_t)rows, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MLA_E_BADARG;
}

template <int MODE, int BN, int SA, int SB>
int launch(const CUtensorMap& wmap, const CUtensorMap& xmap, const StripParams& p, dim3 grid, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)SA * p.a_stage + (size_t)SB * BN * 128;
  static std::atomic<size_t> configured{0};
  if (smem > configured.load(std::memory_order_acquire)) {
    MLA_CUDA_TRY(cudaFuncSetAttribute(conv_strip_kernel<MODE, BN, SA, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(1024 + SA * 32768 + SB * BN * 128)));
    configured.store((size_t)(1024 + SA * 32768 + SB * BN * 128), std::memory_order_release);
  }
  conv_strip_kernel<MODE, BN, SA, SB><<<grid, kThreads, smem, st>>>(wmap, xmap, p);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}

}  // namespace

namespace mla {

// Geometry of the strip path for an H x W image, or tiles == 0 when it does not apply / does not pay:
// MLA_CONV_STRIP=0 disables it; small maps (row utilisation under 60 %) stay on the im2col kernel.
StripPlan strip_plan(int N, int H, int W, int R, int S, int stride, int pad) {
  StripPlan pl{};
  // Off by default: correct, but measured no faster than the im2col kernel on B200 (both are bound by the ~130 ns
  // a tcgen05.mma instruction occupies its CTA and by the ~2 us TMA round trip, not by operand bytes) — DESIGN.md 4.1.
  static const bool off = [] { const char* e = getenv("MLA_CONV_STRIP"); return e == nullptr || e[0] != '1'; }();
  if (off || R != 3 || S != 3 || stride != 1 || pad != 1 || W + 2 > 63 || H < 1) return pl;
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

int conv_strip_run(int mode, const float* src, const float* w, float* out, int N, int H, int W, int Cin, int Cout,
                   int accumulate, float* stat_part, const StripPlan& pl, void* stream) {
  // mode 0: fprop  (src = x [N,H,W,Cin],  out = y  [N,H,W,Cout]);  mode 1: dgrad (src = dy [N,H,W,Cout], out = dx [N,H,W,Cin])
  const int Cs = mode == 0 ? Cin : Cout;        // channels of the gathered tensor = GEMM K per tap
  const int Co = mode == 0 ? Cout : Cin;        // channels of the output = GEMM N
  StripParams p{};
  p.N = N; p.H = H; p.W = W; p.Wp = pl.Wp; p.TR = pl.TR; p.tiles_per_img = pl.tiles_per_img; p.kcb = Cs / 32;
  p.CinW = Cin; p.flip = mode == 1 ? 1 : 0;
  p.a_stage = (uint32_t)mla::align_up((size_t)128 * (130 + 2 * pl.Wp), 1024);
  p.strip_tx = (uint32_t)(128 * pl.Wp * (pl.TR + 2));
  p.out = out; p.ldo = Co; p.accumulate = accumulate; p.Cout = Co; p.stat_part = stat_part;
  static const int dbg = [] { const char* e = getenv("MLA_STRIP_DBG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg;
  const int BN = (Co % 128 == 0) ? 128 : 64;
  CUtensorMap wmap, xmap;
  int rc = mode == 0 ? make_weight_map(&wmap, w, Cout, 9LL * Cin, BN, false) : make_weight_map(&wmap, w, Cout, 9LL * Cin, 32, true);
  if (rc) return rc;
  rc = make_strip_map(&xmap, src, N, H, W, Cs, pl.Wp, pl.TR + 2);
  if (rc) return rc;
  dim3 grid(pl.tiles, Co / BN);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode == 0) return BN == 64 ? launch<0, 64, 2, 4>(wmap, xmap, p, grid, st) : launch<0, 128, 2, 3>(wmap, xmap, p, grid, st);
  return BN == 64 ? launch<1, 64, 2, 4>(wmap, xmap, p, grid, st) : launch<1, 128, 2, 3>(wmap, xmap, p, grid, st);
}

}  // namespace mla
