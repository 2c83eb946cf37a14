// The 7x7 / stride 2 / pad 3 stem convolution of the ResNet-18 encoders (models/backbone.py:78-83,149: Cin = 1 spectrogram or
// 3 RGB -> 64 channels) WITHOUT a materialised im2col matrix — fprop and wgrad, fp16 operands, fp32 accumulation in TMEM.
//
// The im2col route (encoder_ops.cu:stem_im2col + the generic GEMM) writes and re-reads an [M][192] fp16 matrix: 616 MB per
// visual batch, three trips through HBM (write, fprop read, wgrad read) = 22 % of the whole training step. Here:
//   * SPACE-TO-DEPTH. The zero-padded input is regrouped into 2x2 pixel cells: xs[n][r][q][(pr*2+pc)*Cin + c] =
//     x[n][c][2r+pr-3][2q+pc-3] (fp16, 16 channels per cell, the unused ones zero). The stride-2 7x7 convolution becomes a
//     stride-1 4x4 convolution over cells: output (oh, ow) reads cells (oh+a, ow+b), a, b < 4, with the filter regrouped the
//     same way (tap (2a+pr, 2b+pc); the 8th row / column of taps is zero). xs is 55 MB per visual batch and stays in L2.
//   * SLIDING-WINDOW TENSOR MAP. One filter row a of one output pixel is 4 consecutive cells = 64 fp16 = exactly one 128-byte
//     K-major operand row. A TILED tensor map with OVERLAPPING strides — dimension 0 = 64 contiguous elements, dimension 1 =
//     the window start advancing by one cell (32 bytes), dimension 2 = cell rows, dimension 3 = images — lets ONE TMA box
//     {64, 16 windows, 8 rows, 1} deliver the [128 pixels][64] k-block of a 16x8-pixel output tile, 128B-swizzled, zero filled
//     past the image (tests/tools/probes/tma_overlap_probe.cu). 4 boxes (a = 0..3) = K = 256 per tile.
//   * fprop: persistent CTAs, the regrouped filter (64 x 256 fp16 = 32 KB) resident in shared memory, an 8-slot ring of
//     k-blocks, two TMEM accumulators, 8 epilogue warps: BatchNorm partial sums from the fp32 accumulators, fp16 output rows
//     staged in 128B-swizzled shared memory and written by TMA stores (clipped at the image border).
//   * wgrad: dW2[256][64] = sum over pixels patch^T dy. Both operands are the SAME kind of [128 pixels][128 B] tile (MN-major,
//     K = pixels): the patch k-blocks above and the dy tile; every CTA accumulates all of its tiles in TMEM (2 x 128 x 64) and
//     writes one partial; a fixed-order reduction regroups the taps into [64][7][7][Cin] (deterministic).
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kCo = 64;                        // output channels
constexpr int kCell = 16;                      // fp16 channels per space-to-depth cell
constexpr int kTW = 16, kTH = 8;               // output tile: 16 x 8 pixels = 128 GEMM rows
constexpr uint32_t kBlk = 128 * 128;           // one k-block / dy tile / fp16 output tile: 128 rows x 128 B
constexpr uint32_t kW2Bytes = 4 * kCo * 128;   // regrouped filter: 4 k-blocks of [64 co][64 k] fp16

struct StemGeom {
  int N, OH, OW, Hs, Ws;        // Hs = OH + 3, Ws = OW + 3 cells
  int tiles_w, tiles_h, tiles;
};

StemGeom stem_geom(int N, int H, int W) {
  StemGeom g{};
  g.N = N;
  g.OH = (H + 6 - 7) / 2 + 1;
  g.OW = (W + 6 - 7) / 2 + 1;
  g.Hs = g.OH + 3;
  g.Ws = g.OW + 3;
  g.tiles_w = (g.OW + kTW - 1) / kTW;
  g.tiles_h = (g.OH + kTH - 1) / kTH;
  g.tiles = N * g.tiles_w * g.tiles_h;
  return g;
}

// ---------------------------------------------------------------------------------- pack / filter regrouping
// One thread per cell. Input element (n, c, h, w) at in[(n / T) * sB + (n % T) * sT + c * sC + h * W + w] (raw NCHW / NCTHW).
__global__ void __launch_bounds__(256) stem_s2d_pack_kernel(const float* __restrict__ in, uint4* __restrict__ xs, int T,
                                                            long long sB, long long sT, long long sC, int Cin, int H, int W,
                                                            int Hs, int Ws, long long cells) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % Ws);
    long long t = i / Ws;
    const int r = (int)(t % Hs);
    const int n = (int)(t / Hs);
    const float* src = in + (long long)(n / T) * sB + (long long)(n % T) * sT;
    __half v[kCell];
#pragma unroll
    for (int e = 0; e < kCell; ++e) v[e] = __float2half_rn(0.f);
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) {
      const int h = 2 * r + (pp >> 1) - 3, w = 2 * q + (pp & 1) - 3;
      if (h >= 0 && h < H && w >= 0 && w < W) {
        for (int c = 0; c < Cin; ++c) v[pp * Cin + c] = __float2half_rn(__ldg(src + c * sC + (long long)h * W + w));
      }
    }
    const uint4* pv = reinterpret_cast<const uint4*>(v);
    xs[2 * i] = pv[0];
    xs[2 * i + 1] = pv[1];
  }
}

// w [64][7][7][Cin] fp32 -> w2 [64][a 4][b 4][16] fp16, w2[co][a][b][(pr*2+pc)*Cin + c] = w[co][2a+pr][2b+pc][c]
__global__ void stem_s2d_weights_kernel(const float* __restrict__ w, __half* __restrict__ w2, int Cin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kCo * 256) return;
  const int e = i & 15, b = (i >> 4) & 3, a = (i >> 6) & 3, co = i >> 8;
  float v = 0.f;
  if (e < 4 * Cin) {
    const int pp = e / Cin, c = e - pp * Cin;
    const int kh = 2 * a + (pp >> 1), kw = 2 * b + (pp & 1);
    if (kh < 7 && kw < 7) v = w[((co * 7 + kh) * 7 + kw) * Cin + c];
  }
  w2[i] = __float2half_rn(v);
}

// dw[co][kh][kw][c] = *out_scale * sum over partials (in order) of part[p][a*64 + b*16 + (pr*2+pc)*Cin + c][co]
__global__ void stem_s2d_wgrad_reduce_kernel(const float* __restrict__ part, int nparts, const float* __restrict__ out_scale,
                                             float* __restrict__ dw, int Cin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_co = 49 * Cin;
  if (i >= kCo * per_co) return;
  const int co = i & 63, t = i >> 6;            // co fastest: coalesced reads of the partial rows
  const int c = t % Cin, kk = t / Cin;
  const int kh = kk / 7, kw = kk - kh * 7;
  const int row = (kh >> 1) * 64 + (kw >> 1) * 16 + ((kh & 1) * 2 + (kw & 1)) * Cin + c;
  const float* p = part + (size_t)row * kCo + co;
  float acc = 0.f;
  int s = 0;
  for (; s + 8 <= nparts; s += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(p + (size_t)(s + u) * 256 * kCo);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u];
  }
  for (; s < nparts; ++s) acc += __ldcg(p + (size_t)s * 256 * kCo);
  dw[(size_t)co * per_co + t] = acc * (out_scale != nullptr ? __ldg(out_scale) : 1.f);
}

// ---------------------------------------------------------------------------------- TMA store helpers
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m), "r"(src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------- fprop
constexpr int kEpiWarps = 8;
constexpr int kFThreads = (kEpiWarps + 2) * 32;     // 8 epilogue warps, TMA producer, MMA issuer
constexpr int kSlots = 8;                           // ring of 16 KB k-block slots (two tiles ahead)

struct StemFpropParams {
  int tiles_w, tiles_h, tiles, OH, OW;
  float* stat_part;       // [tile][2][64] BatchNorm partial sums of the fp32 accumulators, or NULL
};

__global__ void __launch_bounds__(kFThreads, 1) stem_s2d_fprop_kernel(const __grid_constant__ CUtensorMap tmap_w,
                                                                       const __grid_constant__ CUtensorMap tmap_x,
                                                                       const __grid_constant__ CUtensorMap tmap_o,
                                                                       StemFpropParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_bar, full_bar[kSlots], empty_bar[kSlots], acc_full_bar[2], acc_empty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_stat[kEpiWarps * 2 * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base;                            // 4 x 8 KB filter k-blocks
  const uint32_t a_base = base + kW2Bytes;                 // k-block ring
  const uint32_t o_base = a_base + kSlots * kBlk;          // two 16 KB fp16 output staging tiles

  if (threadIdx.x == 0) {
    tc::mbar_init(tc::smem_u32(&w_bar), 1);
    for (int s = 0; s < kSlots; ++s) {
      tc::mbar_init(tc::smem_u32(&full_bar[s]), 1);
      tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(tc::smem_u32(&acc_full_bar[b]), 1);
      tc::mbar_init(tc::smem_u32(&acc_empty_bar[b]), kEpiWarps);
    }
    tc::fence_mbar_init();
  }
  if (warp == kEpiWarps && lane == 0) {
    tc::tma_prefetch_desc(&tmap_w);
    tc::tma_prefetch_desc(&tmap_x);
    tc::tma_prefetch_desc(&tmap_o);
  }
  if (warp == kEpiWarps + 1) {
    tc::tmem_alloc(tc::smem_u32(&tmem_slot), 2 * kCo);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int per_img = p.tiles_w * p.tiles_h;

  if (warp == kEpiWarps) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(tc::smem_u32(&w_bar), kW2Bytes);
      for (int a = 0; a < 4; ++a) tc::tma_load_2d(w_base + a * (kCo * 128), &tmap_w, tc::smem_u32(&w_bar), a * 64, 0);
      int kb = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const int n = tile / per_img, rem = tile - n * per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        for (int a = 0; a < 4; ++a, ++kb) {
          const int s = kb % kSlots;
          tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ((kb / kSlots) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(tc::smem_u32(&full_bar[s]), kBlk);
          // filter row a of the tile's 16 x 8 pixels: windows ow0 .. ow0+15 of cell rows oh0+a .. oh0+a+7
          tc::tma_load_4d(a_base + s * kBlk, &tmap_x, tc::smem_u32(&full_bar[s]), 0, tw * kTW, th * kTH + a, n);
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_f16(128, kCo, 0, 0, 0, 0);
      tc::mbar_wait(tc::smem_u32(&w_bar), 0);
      int it = 0, kb = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        tc::mbar_wait(tc::smem_u32(&acc_empty_bar[buf]), ((it >> 1) & 1) ^ 1);
        const uint32_t acc = tmem_base + (uint32_t)buf * kCo;
        for (int a = 0; a < 4; ++a, ++kb) {
          const int s = kb % kSlots;
          tc::mbar_wait(tc::smem_u32(&full_bar[s]), (kb / kSlots) & 1);
          tc::tc_fence_after();
          const uint32_t a0 = a_base + s * kBlk, b0 = w_base + a * (kCo * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = tc::make_smem_desc(a0 + k * 32u, 16u, 1024u, tc::kLayoutSw128);
            const uint64_t bd = tc::make_smem_desc(b0 + k * 32u, 16u, 1024u, tc::kLayoutSw128);
            tc::umma_f16(acc, ad, bd, idesc, (a | k) != 0 ? 1u : 0u);
          }
          tc::umma_commit(tc::smem_u32(&empty_bar[s]));
        }
        tc::umma_commit(tc::smem_u32(&acc_full_bar[buf]));
      }
    }
  } else {
    // ===================== epilogue (warps 0-7): TMEM lane quadrant q, column half hc =====================
    const int q = warp & 3, hc = warp >> 2;
    const int m = q * 32 + lane;                      // GEMM row = pixel (lr, lc) of the tile
    const int lr = m >> 4, lc = m & 15;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int n = tile / per_img, rem = tile - n * per_img;
      const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
      const bool valid = th * kTH + lr < p.OH && tw * kTW + lc < p.OW;
      tc::mbar_wait(tc::smem_u32(&acc_full_bar[buf]), (it >> 1) & 1);
      tc::tc_fence_after();
      uint32_t v[32];
      tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kCo + hc * 32), v);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty_bar[buf]));
      // the staging tile of two tiles ago must have been read by its TMA store
      const uint32_t stg = o_base + (uint32_t)buf * kBlk;
      if (threadIdx.x == 0) bulk_wait_read<1>();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      {
        const uint32_t rowp = stg + (uint32_t)m * 128u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t h[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const __half2 hh = __floats2half2_rn(__uint_as_float(v[8 * j + 2 * u]), __uint_as_float(v[8 * j + 2 * u + 1]));
            h[u] = *reinterpret_cast<const uint32_t*>(&hh);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + tc::swz16(hc * 4 + j, m)), "r"(h[0]), "r"(h[1]),
                       "r"(h[2]), "r"(h[3])
                       : "memory");
        }
      }
      if (p.stat_part != nullptr) {
        float a[32], b[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { a[j] = valid ? __uint_as_float(v[j]) : 0.f; b[j] = a[j] * a[j]; }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {   // butterfly transpose-reduce: lane j ends with column hc*32 + j
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < off; ++i) {
            const float sa = up ? a[i] : a[i + off], ka = up ? a[i + off] : a[i];
            const float sb = up ? b[i] : b[i + off], kb2 = up ? b[i + off] : b[i];
            a[i] = ka + __shfl_xor_sync(0xffffffffu, sa, off);
            b[i] = kb2 + __shfl_xor_sync(0xffffffffu, sb, off);
          }
        }
        s_stat[(warp * 2 + 0) * 32 + lane] = a[0];
        s_stat[(warp * 2 + 1) * 32 + lane] = b[0];
      }
      tc::fence_proxy_async();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) {
        tma_store_4d(&tmap_o, stg, 0, tw * kTW, th * kTH, n);
        bulk_commit();
      }
      if (p.stat_part != nullptr && threadIdx.x < kCo) {
        const int t = threadIdx.x, h2 = t >> 5, cl = t & 31;
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {                  // the four lane quadrants, fixed order
          sa += s_stat[((h2 * 4 + w) * 2 + 0) * 32 + cl];
          sb += s_stat[((h2 * 4 + w) * 2 + 1) * 32 + cl];
        }
        float* dstp = p.stat_part + (size_t)tile * 2 * kCo + t;
        dstp[0] = sa;
        dstp[kCo] = sb;
      }
    }
    if (threadIdx.x == 0) bulk_wait_all();
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) tc::tmem_dealloc(tmem_base, 2 * kCo);
}

// ---------------------------------------------------------------------------------- wgrad
constexpr int kWThreads = 192;
constexpr int kWStages = 2;
constexpr uint32_t kWStage = 5 * kBlk;          // 4 patch k-blocks + the dy tile

struct StemWgradParams {
  int tiles_w, tiles_h, tiles;
  float* part;            // [gridDim.x][256][64]
};

__global__ void __launch_bounds__(kWThreads, 1) stem_s2d_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                       const __grid_constant__ CUtensorMap tmap_dy,
                                                                       StemWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWStages], empty_bar[kWStages], done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWStages; ++s) {
      tc::mbar_init(tc::smem_u32(&full_bar[s]), 1);
      tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1);
    }
    tc::mbar_init(tc::smem_u32(&done_bar), 1);
    tc::fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_x);
    tc::tma_prefetch_desc(&tmap_dy);
  }
  if (warp == 5) {
    tc::tmem_alloc(tc::smem_u32(&tmem_slot), 2 * kCo);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int per_img = p.tiles_w * p.tiles_h;

  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int s = it % kWStages;
        const int n = tile / per_img, rem = tile - n * per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ((it / kWStages) & 1) ^ 1);
        const uint32_t bar = tc::smem_u32(&full_bar[s]);
        tc::mbar_arrive_expect_tx(bar, kWStage);
        const uint32_t st = base + s * kWStage;
        for (int a = 0; a < 4; ++a) tc::tma_load_4d(st + a * kBlk, &tmap_x, bar, 0, tw * kTW, th * kTH + a, n);
        tc::tma_load_4d(st + 4 * kBlk, &tmap_dy, bar, 0, tw * kTW, th * kTH, n);   // pixels past the image: zero rows
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // D[h][m = (a & 1) * 64 + e][co] += sum over the tile's 128 pixels of patch[pix][a = 2h + (m >> 6)][e] * dy[pix][co]
      // A: two [128 pixels][128 B] panels (filter rows 2h, 2h+1) LBO = 16 KB apart, B: the dy panel; both MN-major,
      // 16 pixels (two 8-row swizzle atoms, SBO = 1 KB) per instruction
      constexpr uint32_t idesc = tc::make_idesc_f16(128, kCo, 0, 0, 1, 1);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int s = it % kWStages;
        tc::mbar_wait(tc::smem_u32(&full_bar[s]), (it / kWStages) & 1);
        tc::tc_fence_after();
        const uint32_t st = base + s * kWStage;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t ad = tc::make_smem_desc(st + (2 * h) * kBlk + k * 2048u, kBlk, 1024u, tc::kLayoutSw128);
            const uint64_t bd = tc::make_smem_desc(st + 4 * kBlk + k * 2048u, kBlk, 1024u, tc::kLayoutSw128);
            tc::umma_f16(tmem_base + (uint32_t)h * kCo, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
          }
        }
        tc::umma_commit(tc::smem_u32(&empty_bar[s]));
      }
      tc::umma_commit(tc::smem_u32(&done_bar));
    }
  } else {
    tc::mbar_wait(tc::smem_u32(&done_bar), 0);
    tc::tc_fence_after();
    const int m = warp * 32 + lane;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      float* orow = p.part + ((size_t)blockIdx.x * 256 + h * 128 + m) * kCo;
#pragma unroll 1
      for (int c = 0; c < kCo; c += 32) {
        uint32_t v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(h * kCo + c), v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          reinterpret_cast<float4*>(orow + c)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tmem_base, 2 * kCo);
}

// ---------------------------------------------------------------------------------- host side
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

int encode4(CUtensorMap* m, const void* ptr, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return MLA_E_NODEVICE;
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MLA_E_BADARG;
}

// sliding windows over the cell tensor: {64 elements of a window, OW window starts (32 B apart), Hs cell rows, N images}
int make_window_map(CUtensorMap* m, const void* xs, const StemGeom& g) {
  cuuint64_t dims[4] = {64, (cuuint64_t)g.OW, (cuuint64_t)g.Hs, (cuuint64_t)g.N};
  cuuint64_t strides[3] = {(cuuint64_t)kCell * 2, (cuuint64_t)g.Ws * kCell * 2, (cuuint64_t)g.Hs * g.Ws * kCell * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)kTW, (cuuint32_t)kTH, 1u};
  return encode4(m, xs, dims, strides, box);
}

// NHWC fp16 [N][OH][OW][64]: box {64 channels, 16 pixels, 8 rows, 1 image}
int make_pixel_map(CUtensorMap* m, const void* y, const StemGeom& g) {
  cuuint64_t dims[4] = {(cuuint64_t)kCo, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)g.N};
  cuuint64_t strides[3] = {(cuuint64_t)kCo * 2, (cuuint64_t)g.OW * kCo * 2, (cuuint64_t)g.OH * g.OW * kCo * 2};
  cuuint32_t box[4] = {(cuuint32_t)kCo, (cuuint32_t)kTW, (cuuint32_t)kTH, 1u};
  return encode4(m, y, dims, strides, box);
}

int make_w2_map(CUtensorMap* m, const void* w2) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return MLA_E_NODEVICE;
  cuuint64_t dims[2] = {256, (cuuint64_t)kCo};
  cuuint64_t strides[1] = {256 * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)kCo};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w2), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MLA_E_BADARG;
}

bool geom_ok(int N, int H, int W, int Cin) { return N >= 1 && H >= 1 && W >= 1 && Cin >= 1 && Cin <= 4; }

}  // namespace

extern "C" long long mla_stem_s2d_input_elems(int N, int H, int W) {
  if (N < 1 || H < 1 || W < 1) return 0;
  const StemGeom g = stem_geom(N, H, W);
  return (long long)N * g.Hs * g.Ws * kCell;
}

extern "C" int mla_stem_s2d_tiles(int N, int H, int W) {
  if (N < 1 || H < 1 || W < 1) return 0;
  return stem_geom(N, H, W).tiles;
}

extern "C" int mla_stem_s2d_pack(const float* in, void* xs16, int N, int T, long long sB, long long sT, long long sC, int Cin,
                                 int H, int W, void* stream) {
  if (!in || !xs16 || !mla::aligned16(xs16) || T < 1 || !geom_ok(N, H, W, Cin)) return MLA_E_BADARG;
  const StemGeom g = stem_geom(N, H, W);
  const long long cells = (long long)N * g.Hs * g.Ws;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  long long blocks = (cells + 255) / 256;
  if (blocks > (long long)di.sm_count * 16) blocks = (long long)di.sm_count * 16;
  stem_s2d_pack_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, static_cast<uint4*>(xs16), T, sB, sT, sC,
                                                                                 Cin, H, W, g.Hs, g.Ws, cells);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}

extern "C" int mla_stem_s2d_weights(const float* w, void* w2_16, int Cin, void* stream) {
  if (!w || !w2_16 || !mla::aligned16(w2_16) || Cin < 1 || Cin > 4) return MLA_E_BADARG;
  stem_s2d_weights_kernel<<<kCo * 256 / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, static_cast<__half*>(w2_16), Cin);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}

extern "C" int mla_stem_s2d_fprop(const void* xs16, const void* w2_16, void* y16, int N, int H, int W, float* stat_part,
                                  void* stream) {
  if (!xs16 || !w2_16 || !y16 || !mla::aligned16(xs16) || !mla::aligned16(w2_16) || !mla::aligned16(y16) || N < 1 || H < 1 ||
      W < 1)
    return MLA_E_BADARG;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  const StemGeom g = stem_geom(N, H, W);
  CUtensorMap wmap, xmap, omap;
  int rc = make_w2_map(&wmap, w2_16);
  if (rc) return rc;
  rc = make_window_map(&xmap, xs16, g);
  if (rc) return rc;
  rc = make_pixel_map(&omap, y16, g);
  if (rc) return rc;
  StemFpropParams p{};
  p.tiles_w = g.tiles_w; p.tiles_h = g.tiles_h; p.tiles = g.tiles; p.OH = g.OH; p.OW = g.OW; p.stat_part = stat_part;
  const size_t smem = 1024 + kW2Bytes + (size_t)kSlots * kBlk + 2 * (size_t)kBlk;
  static std::atomic<int> cfg{0};
  if (!cfg.load(std::memory_order_acquire)) {
    MLA_CUDA_TRY(cudaFuncSetAttribute(stem_s2d_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfg.store(1, std::memory_order_release);
  }
  const int grid = std::min(g.tiles, di.sm_count);
  stem_s2d_fprop_kernel<<<grid, kFThreads, smem, static_cast<cudaStream_t>(stream)>>>(wmap, xmap, omap, p);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}

extern "C" size_t mla_stem_s2d_wgrad_workspace_bytes(void) {
  const mla::DeviceInfo& di = mla::device_info();
  const int ctas = di.ok == 1 ? di.sm_count : 256;
  return (size_t)ctas * 256 * kCo * sizeof(float);
}

extern "C" int mla_stem_s2d_wgrad(const void* xs16, const void* dy16, const float* out_scale, float* dw, int N, int H, int W,
                                  int Cin, void* ws, size_t ws_bytes, void* stream) {
  if (!xs16 || !dy16 || !dw || !mla::aligned16(xs16) || !mla::aligned16(dy16) || !geom_ok(N, H, W, Cin)) return MLA_E_BADARG;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  const StemGeom g = stem_geom(N, H, W);
  const int grid = std::min(g.tiles, di.sm_count);
  if (!ws || !mla::aligned16(ws) || ws_bytes < (size_t)grid * 256 * kCo * sizeof(float)) return MLA_E_WORKSPACE;
  CUtensorMap xmap, dmap;
  int rc = make_window_map(&xmap, xs16, g);
  if (rc) return rc;
  rc = make_pixel_map(&dmap, dy16, g);
  if (rc) return rc;
  StemWgradParams p{};
  p.tiles_w = g.tiles_w; p.tiles_h = g.tiles_h; p.tiles = g.tiles; p.part = static_cast<float*>(ws);
  const size_t smem = 1024 + (size_t)kWStages * kWStage;
  static std::atomic<int> cfg{0};
  if (!cfg.load(std::memory_order_acquire)) {
    MLA_CUDA_TRY(cudaFuncSetAttribute(stem_s2d_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfg.store(1, std::memory_order_release);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  stem_s2d_wgrad_kernel<<<grid, kWThreads, smem, st>>>(xmap, dmap, p);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  const int outs = kCo * 49 * Cin;
  stem_s2d_wgrad_reduce_kernel<<<(outs + 127) / 128, 128, 0, st>>>(p.part, grid, out_scale, dw, Cin);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}
