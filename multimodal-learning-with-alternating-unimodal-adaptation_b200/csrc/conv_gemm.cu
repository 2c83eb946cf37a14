// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma kind::tf32, fp32
// accumulators in TMEM) for the ResNet-18 encoders — reference models/backbone.py:39-50
// (BasicBlock convs), :126-129 (1x1 downsample) and their autograd backward.
//
// Activations are NHWC fp32, weights are [Cout][R][S][Cin] (= torch channels_last memory of the
// reference's OIHW parameter, so state dicts stay compatible and no weight transform is needed).
//
//   MODE 0  fprop   y[m, co]  = sum_{r,s,ci} x[pix(m,r,s), ci] * w[co, r,s,ci]     M = N*OH*OW
//   MODE 1  dgrad   dx[m, ci] = sum_{r,s,co} dy[pix'(m,r,s), co] * w[co, r,s,ci]   M = N*H*W
//   MODE 2  wgrad   dw[co, r,s,ci] = sum_m dy[m, co] * x[pix(m,r,s), ci]           K = N*OH*OW (split)
//
// One CTA computes a 128 x BN accumulator tile. Warp roles (192 threads):
//   warp 4     TMA producer. The dense operand (weights, or dy for wgrad) comes through a 2-D tiled
//              tensor map; the GATHERED operand (the im2col view of x or dy) comes through an
//              im2col-mode tensor map (cuTensorMapEncodeIm2col): one cp.async.bulk.tensor.4d...im2col
//              per k-block loads 128 (or 32) pixels x 32 channels of one filter tap, zero-filling
//              padding and tails in hardware, straight into the 128B-swizzled layout the UMMA
//              descriptors expect. Both complete on the stage's mbarrier (complete_tx).
//   warps 0-3  epilogue (tcgen05.ld 32x32b -> registers -> global). Only for the stride-2 dgrad,
//              whose gather has holes the im2col walk cannot express, the same warps first act as
//              cp.async gather producers (one 128-byte channel segment per thread and k-block).
//   warp 5     TMEM allocation + the single MMA-issuing thread: 4 x tcgen05.mma (K = 8 each)
//              per 32-wide k-block, tcgen05.commit releases the smem stage / publishes TMEM.
// K-major operands are [rows][32 tf32]; MN-major operands (weights in dgrad, both operands in
// wgrad) are panels of [k rows][32 elements along M/N] — both are the same physical image, which
// is why one gather routine serves all three modes.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kGatherThreads = 128;
constexpr int kThreads = 192;
constexpr uint32_t kABytes = 128 * 128;  // 128 rows x 128 B

struct ConvGemmParams {
  const float* src;  // gathered tensor: x (fprop, wgrad) or dy (dgrad), NHWC
  int Hs, Ws, Cs;    // its spatial size and channel count
  int OH, OW;        // the pixel space that indexes GEMM rows (fprop/dgrad) or GEMM K (wgrad)
  int M;             // number of such pixels = N * OH * OW
  int R, S;
  int mul, sgn, off, div;  // source row: (oh*mul + sgn*r + off) / div, must divide and be in range
  int kcb;           // 32-channel blocks per tap of the gathered tensor (fprop, dgrad)
  int KB;            // k-blocks this launch iterates (fprop/dgrad: R*S*kcb)
  int CinW;          // Cin of the weight tensor (column offset of a tap in the 2-D weight view)
  float* out;
  long long ldo;     // output row stride in elements
  int accumulate;    // out += acc instead of out = acc (dgrad into an existing gradient)
  // wgrad
  int Cout;
  int kb_per_split;  // k-blocks (of 32 pixels) per split
  int KBtot;         // total k-blocks = ceil(M / 32)
  int splits;
  long long split_stride;  // elements between split partials
  // im2col-mode TMA gather (IM2COL kernels): base pixel of GEMM row / K index (oh, ow) is
  // (ow * mul + g_base_w, oh * mul + g_base_h). The launch iterates nr x ns filter taps: tap (i, j) uses the
  // weights of filter position (tap_r[i], tap_s[j]) and the instruction offset (off_s[j], off_r[i]).
  int g_base_w, g_base_h;
  int nr, ns;
  signed char tap_r[8], tap_s[8], off_r[8], off_s[8];
  // output row map (stride-s dgrad, one launch per output parity class): GEMM row (n, i, j) over the OH x OW
  // sub-grid is written to dx pixel (n, i * o_mul + o_ph, j * o_mul + o_pw) of an o_H x o_W image.
  int o_mul, o_ph, o_pw, o_H, o_W;
  // fprop only: per-M-tile BatchNorm partial sums of the OUTPUT, stat_part[m tile][2][Cout] = (sum y, sum y*y)
  // over the tile's 128 rows, taken from the fp32 accumulators (NULL = off)
  float* stat_part;
  // fprop-type launches only (Linear layers): out = acc + bias[column] + resid[row][column] (either may be NULL)
  const float* bias;
  const float* resid;
  // every mode: the accumulator is multiplied by *out_scale (a device scalar, NULL = 1) before anything is added —
  // undoes the power-of-two scale of fp16 gradient operands
  const float* out_scale;
};

template <bool MN_MAJOR>
__device__ __forceinline__ void gather_segment(const ConvGemmParams& p, uint32_t dst_row, int row, int m, int r, int s,
                                               int c0) {
  // dst_row: smem address of the 128-byte row `row` of its tile/panel
  const float* src = p.src;
  uint32_t bytes = 0;
  if (m < p.M) {
    const int ow = m % p.OW;
    const int t = m / p.OW;
    const int oh = t % p.OH;
    const int n = t / p.OH;
    const int hn = oh * p.mul + p.sgn * r + p.off;
    const int wn = ow * p.mul + p.sgn * s + p.off;
    if (hn >= 0 && wn >= 0) {
      int hs = hn, ws = wn;
      bool ok = true;
      if (p.div != 1) {
        hs = hn / p.div;
        ws = wn / p.div;
        ok = (hs * p.div == hn) && (ws * p.div == wn);
      }
      if (ok && hs < p.Hs && ws < p.Ws) {
        src = p.src + (((long long)n * p.Hs + hs) * p.Ws + ws) * p.Cs + c0;
        bytes = 16;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    tc::cp_async16(dst_row + (MN_MAJOR ? tc::swz32(j, row) : tc::swz16(j, row)), src + (bytes ? 4 * j : 0), bytes);
}

// NT > 1 (wgrad only, Cin == BN): one CTA accumulates NT consecutive filter taps side by side — the N tile is
// NT x BN columns of the flattened (tap, ci) axis of dw, so dy (the A operand) is loaded once for NT taps.
// CL > 1 (fprop / dgrad, im2col): CL CTAs with consecutive M tiles and the SAME weight tile form a thread-block
// cluster; each loads 1/CL of the weight tile per k-block and TMA-multicasts it into all CL shared memories, so
// the weights cross L2->SM once per cluster instead of once per CTA. A stage is released cluster-wide: every
// CTA's tcgen05.commit arrives on the stage's empty barrier of all CL CTAs.
// PAIR (fprop / dgrad, im2col): two CTAs with consecutive M tiles form a tcgen05 CTA pair (cta_group::2): one
// 256 x BN MMA per k-step, issued by the leader, reads each CTA's own 128 A rows and HALF of the weight tile from
// each CTA's shared memory — every SM ingests BN/2 weight rows instead of BN. Loads of both CTAs complete on the
// leader's full barrier; its tcgen05.commit releases the stage in both CTAs.
// ET != 0 (MODE 0, im2col): 2-byte operands through kind::f16 — ET 1: A fp16, B fp16; 2: A bf16, B bf16; 3: A bf16,
// B fp16. A 128-byte operand row then holds 64 channels, so a k-block is 64 channels deep: half the k-blocks, MMA
// instructions and bytes of the TF32 path for the same convolution. Layouts, swizzles and k-steps are byte-identical.
template <int MODE, int BN, int STAGES, bool IM2COL, int NT = 1, int CL = 1, bool PAIR = false, int ET = 0>
__global__ void __launch_bounds__(kThreads) conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap,
                                                             const __grid_constant__ CUtensorMap tmap_g, ConvGemmParams p) {
  static_assert(NT == 1 || (MODE == 2 && IM2COL), "multi-tap tiles exist for the im2col wgrad only");
  static_assert(CL == 1 || (MODE != 2 && IM2COL && NT == 1), "weight multicast exists for the im2col fprop / dgrad only");
  static_assert(CL == 1 || MODE == 0 || (BN / 32) % CL == 0, "dgrad splits whole 32-column weight panels");
  static_assert(!PAIR || (CL == 1 && NT == 1 && IM2COL && MODE != 2), "CTA pairs exist for the im2col fprop / dgrad only");
  static_assert(ET == 0 || (MODE != 1 && IM2COL && CL == 1), "2-byte operands: im2col fprop-type and wgrad only");
  static_assert(ET == 0 || !PAIR || MODE == 0, "2-byte CTA pairs: fprop-type only");
  static_assert(ET == 0 || MODE == 0 || ET == 2 || ET == 1, "the 2-byte wgrad takes bf16 x bf16 or fp16 x fp16");
  // MN-major panels (wgrad): [K rows = pixels][128 B along M/N]; 32 tf32 or 64 2-byte elements wide, 32 / 64 pixels deep
  constexpr int PW = ET == 0 ? 32 : 64;
  constexpr uint32_t kPanel = ET == 0 ? 4096u : 8192u;
  constexpr int KE = ET == 0 ? 32 : 64;   // elements of K per 128-byte operand row = per k-block
  constexpr uint16_t kClMask = (uint16_t)((1u << CL) - 1);
  constexpr int kCluster = PAIR ? 2 : CL;
  constexpr int NTOT = BN * NT;                                     // accumulator columns
  constexpr uint32_t kTmemCols = NTOT <= 64 ? 64 : (NTOT <= 128 ? 128 : 256);
  static_assert(NTOT <= 256, "one UMMA covers at most 256 columns");
  constexpr uint32_t kBBytes = (PAIR ? NTOT / 2 : NTOT) * 128;      // weight bytes per stage held by THIS CTA
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr int LAG = STAGES - 1;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_stat[MODE == 0 ? 4 * 2 * BN : 1];   // [warp][sum | sumsq][column] (fprop BN statistics)
  static_assert(BN <= 256, "tile width");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(tc::smem_u32(&full_bar[s]), IM2COL ? 1 : kGatherThreads + 1);
      tc::mbar_init(tc::smem_u32(&empty_bar[s]), CL);
    }
    tc::mbar_init(tc::smem_u32(&tmem_full_bar), 1);
    tc::fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tc::tma_prefetch_desc(&tmap);
    if (IM2COL) tc::tma_prefetch_desc(&tmap_g);
  }
  if (warp == 5) {
    if (PAIR) {
      tc::tmem_alloc2(tc::smem_u32(&tmem_slot), kTmemCols);
      tc::tmem_relinquish2();
    } else {
      tc::tmem_alloc(tc::smem_u32(&tmem_slot), kTmemCols);
      tc::tmem_relinquish();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (kCluster > 1) tc::cluster_sync();   // every CTA's barriers are initialised before any remote arrive / multicast
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t cl_rank = kCluster > 1 ? tc::cluster_ctarank() : 0u;

  // tile coordinates
  int m0 = 0, n0 = 0, tap_r = 0, tap_s = 0, kb_begin = 0, KB = p.KB, split = 0;
  if (MODE == 2) {
    n0 = blockIdx.x * BN;    // ci tile
    m0 = blockIdx.y * 128;   // co tile
    const int tap = (blockIdx.z / p.splits) * NT;   // first of the NT taps of this CTA
    split = blockIdx.z - (blockIdx.z / p.splits) * p.splits;
    tap_r = tap / p.S;
    tap_s = tap - tap_r * p.S;
    kb_begin = split * p.kb_per_split;
    KB = min(p.kb_per_split, p.KBtot - kb_begin);
    if (KB < 0) KB = 0;
  } else {
    m0 = blockIdx.x * 128;
    n0 = blockIdx.y * BN;
  }

  if (warp < 4) {
    // ===================== gather producers =====================
    const int t = threadIdx.x;
    for (int kb = 0; kb < (IM2COL ? 0 : KB); ++kb) {
      const int s = kb % STAGES;
      tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ((kb / STAGES) & 1) ^ 1);
      const uint32_t stage = tiles + s * kStageBytes;
      if (MODE == 2) {
        // B operand: BN/32 panels x 32 pixel rows; segment q = (panel, pixel row)
        const int pix0 = (kb_begin + kb) * 32;
        for (int q = t; q < BN; q += kGatherThreads) {
          const int i = q & 31, pnl = q >> 5;
          gather_segment<true>(p, stage + kABytes + pnl * 4096 + i * 128, i, pix0 + i, tap_r, tap_s, n0 + pnl * 32);
        }
      } else {
        const int tap = kb / p.kcb;
        const int cb = kb - tap * p.kcb;
        const int r = tap / p.S;
        gather_segment<false>(p, stage + t * 128, t, m0 + t, r, tap - r * p.S, cb * 32);
      }
      tc::cp_async_commit();
      if (kb >= LAG) {
        tc::cp_async_wait<LAG>();
        tc::fence_proxy_async();
        tc::mbar_arrive(tc::smem_u32(&full_bar[(kb - LAG) % STAGES]));
      }
    }
    if (!IM2COL) {
      tc::cp_async_wait<0>();
      tc::fence_proxy_async();
      for (int kb = max(KB - LAG, 0); kb < KB; ++kb) tc::mbar_arrive(tc::smem_u32(&full_bar[kb % STAGES]));
    }

    // ===================== epilogue =====================
    if (KB > 0) {
      tc::mbar_wait(tc::smem_u32(&tmem_full_bar), 0);
      tc::tc_fence_after();
    }
    const int row = warp * 32 + lane;
    float* orow = nullptr;
    if (MODE == 2) {
      const int co = m0 + row;
      if (co < p.Cout)
        orow = p.out + (long long)split * p.split_stride + (long long)co * p.ldo + (tap_r * p.S + tap_s) * p.CinW + n0;
    } else {
      const int m = m0 + row;
      if (m < p.M) {
        long long orow_idx = m;
        if (p.o_mul > 1) {
          const int j = m % p.OW;
          const int t = m / p.OW;
          orow_idx = ((long long)(t / p.OH) * p.o_H + (t % p.OH) * p.o_mul + p.o_ph) * p.o_W + j * p.o_mul + p.o_pw;
        }
        orow = p.out + orow_idx * p.ldo + n0;
      }
    }
    const float* rrow = (MODE == 0 && p.resid != nullptr && orow != nullptr) ? p.resid + (long long)(m0 + row) * p.ldo + n0 : nullptr;
    const float oscale = p.out_scale != nullptr ? __ldg(p.out_scale) : 1.f;
#pragma unroll 1
    for (int c = 0; c < NTOT; c += 32) {
      if (MODE == 2 && NT == 1 && n0 + c >= p.CinW) break;   // partial last ci tile (Cin % BN == 32)
      uint32_t v[32];
      if (KB > 0) {
        tc::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
        tc::tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (orow != nullptr) {
        // every optional addend is loaded for the whole 32-column chunk BEFORE the first store, so the loads are in flight
        // together (interleaving load / add / store per float4 serialises eight global round trips per chunk)
        float4* dst = reinterpret_cast<float4*>(orow + c);
        float4 o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                             __uint_as_float(v[4 * j + 3]));
        if (p.out_scale != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { o[j].x *= oscale; o[j].y *= oscale; o[j].z *= oscale; o[j].w *= oscale; }
        }
        if (p.accumulate) {
          float4 old[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) old[j] = dst[j];
#pragma unroll
          for (int j = 0; j < 8; ++j) { o[j].x += old[j].x; o[j].y += old[j].y; o[j].z += old[j].z; o[j].w += old[j].w; }
        }
        if (MODE == 0 && rrow != nullptr) {
          float4 rv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) rv[j] = reinterpret_cast<const float4*>(rrow + c)[j];
#pragma unroll
          for (int j = 0; j < 8; ++j) { o[j].x += rv[j].x; o[j].y += rv[j].y; o[j].z += rv[j].z; o[j].w += rv[j].w; }
        }
        if (MODE == 0 && p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c) + j);
            o[j].x += bv.x; o[j].y += bv.y; o[j].z += bv.z; o[j].w += bv.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = o[j];
      }
      if (MODE == 0 && p.stat_part != nullptr) {
        // column sums over this warp's 32 rows (rows past M are exact zeros): butterfly transpose-reduce, lane j
        // ends up with column c + j. Fixed shuffle tree -> deterministic.
        float a[32], b[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { a[j] = __uint_as_float(v[j]); b[j] = a[j] * a[j]; }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
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
      for (int t = threadIdx.x; t < BN && m0 < p.M; t += 128) {
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) { sa += s_stat[(w * 2 + 0) * BN + t]; sb += s_stat[(w * 2 + 1) * BN + t]; }
        float* dstp = p.stat_part + (size_t)blockIdx.x * 2 * p.Cout + n0 + t;
        dstp[0] = sa;
        dstp[p.Cout] = sb;
      }
    }
  } else if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // base pixel of this tile's first GEMM row (fprop / dgrad) in the gathered tensor's coordinates
      int gw = 0, gh = 0, gn = 0;
      if (IM2COL && MODE != 2) {
        const int ow = m0 % p.OW;
        const int t = m0 / p.OW;
        gw = ow * p.mul + p.g_base_w;
        gh = (t % p.OH) * p.mul + p.g_base_h;
        gn = t / p.OH;
      }
      const uint32_t tx_gather = IM2COL ? (MODE == 2 ? kBBytes : kABytes) : 0u;
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ((kb / STAGES) & 1) ^ 1);
        const uint32_t stage = tiles + s * kStageBytes;
        const uint32_t bar = tc::smem_u32(&full_bar[s]);
        if (MODE == 0 || MODE == 1) {
          int tap = kb / p.kcb;
          const int cb = kb - tap * p.kcb;
          if (PAIR) {
            if (cl_rank == 0) tc::mbar_arrive_expect_tx(bar, 2u * (kABytes + kBBytes));   // both CTAs' bytes land here
          } else {
            tc::mbar_arrive_expect_tx(bar, kBBytes + tx_gather);
          }
          if (IM2COL) {
            const int ti = tap / p.ns, tj = tap - ti * p.ns;
            if (PAIR)
              tc::tma_load_im2col_4d_2sm(stage, &tmap_g, bar, cb * KE, gw, gh, gn, (uint16_t)p.off_s[tj],
                                         (uint16_t)p.off_r[ti]);
            else
              tc::tma_load_im2col_4d(stage, &tmap_g, bar, cb * KE, gw, gh, gn, (uint16_t)p.off_s[tj], (uint16_t)p.off_r[ti]);
            tap = p.tap_r[ti] * p.S + p.tap_s[tj];   // filter position whose weights this k-block multiplies
          }
          if (MODE == 0) {
            if (PAIR) {   // box {KE k, BN / 2 rows}: this CTA's half of the weight tile
              tc::tma_load_2d_2sm(stage + kABytes, &tmap, bar, tap * p.CinW + cb * KE, n0 + (int)cl_rank * (BN / 2));
            } else if (CL == 1) {
              tc::tma_load_2d(stage + kABytes, &tmap, bar, tap * p.CinW + cb * KE, n0);  // box {KE k, BN rows}
            } else {   // box {32 k, BN / CL rows}: this CTA's slice of the weight tile, delivered to the whole cluster
              tc::tma_load_2d_mc(stage + kABytes + cl_rank * (BN / CL) * 128, &tmap, bar, tap * p.CinW + cb * 32,
                                 n0 + (int)cl_rank * (BN / CL), kClMask);
            }
          } else {
            if (PAIR) {
#pragma unroll
              for (int pnl = 0; pnl < BN / 64; ++pnl)  // this CTA's half of the 32-column weight panels
                tc::tma_load_2d_2sm(stage + kABytes + pnl * 4096, &tmap, bar,
                                    tap * p.CinW + n0 + ((int)cl_rank * (BN / 64) + pnl) * 32, cb * 32);
            } else {
#pragma unroll
              for (int pnl = 0; pnl < BN / 32; ++pnl) {  // box {32 ci, 32 co rows}
                if (CL == 1)
                  tc::tma_load_2d(stage + kABytes + pnl * 4096, &tmap, bar, tap * p.CinW + n0 + pnl * 32, cb * 32);
                else if ((pnl % CL) == (int)cl_rank)
                  tc::tma_load_2d_mc(stage + kABytes + pnl * 4096, &tmap, bar, tap * p.CinW + n0 + pnl * 32, cb * 32, kClMask);
              }
            }
          }
        } else {
          // panels of a partial last ci tile / of co rows beyond Cout are neither loaded nor stored (their
          // accumulator columns / rows are garbage that nobody reads)
          const int npnl = IM2COL ? min(BN / PW, (p.CinW - n0) / PW) : BN / PW;
          const int napnl = IM2COL ? min(128 / PW, (p.Cout - m0 + PW - 1) / PW) : 128 / PW;
          tc::mbar_arrive_expect_tx(bar, (uint32_t)napnl * kPanel + (IM2COL ? (uint32_t)(npnl * NT) * kPanel : 0u));
#pragma unroll
          for (int pnl = 0; pnl < 128 / PW; ++pnl)  // box {PW co, PW pixel rows}
            if (pnl < napnl) tc::tma_load_2d(stage + pnl * kPanel, &tmap, bar, m0 + pnl * PW, (kb_begin + kb) * PW);
          if (IM2COL) {
            const int pix = (kb_begin + kb) * PW;
            const int ow = pix % p.OW;
            const int t = pix / p.OW;
            const int w = ow * p.mul + p.g_base_w, h = (t % p.OH) * p.mul + p.g_base_h, n = t / p.OH;
#pragma unroll
            for (int tp = 0; tp < NT; ++tp) {
              const int tr = NT == 1 ? tap_r : (tap_r * p.S + tap_s + tp) / p.S;
              const int ts = NT == 1 ? tap_s : (tap_r * p.S + tap_s + tp) - tr * p.S;
#pragma unroll
              for (int pnl = 0; pnl < BN / PW; ++pnl)  // PW pixels x PW ci of filter tap (tr, ts)
                if (pnl < npnl)
                  tc::tma_load_im2col_4d(stage + kABytes + (tp * (BN / PW) + pnl) * kPanel, &tmap_g, bar, n0 + pnl * PW, w, h, n,
                                         (uint16_t)ts, (uint16_t)tr);
            }
          }
        }
      }
    }
  } else {
    // ===================== MMA issuer =====================
    if (lane == 0 && KB > 0 && (!PAIR || cl_rank == 0)) {
      constexpr uint32_t idesc = ET == 0 ? tc::make_idesc_tf32(PAIR ? 256 : 128, NTOT, MODE == 2 ? 1 : 0, MODE != 0 ? 1 : 0)
                                         : tc::make_idesc_f16(PAIR ? 256 : 128, NTOT, ET >= 2 ? 1 : 0, ET == 2 ? 1 : 0,
                                                              MODE == 2 ? 1 : 0, MODE == 2 ? 1 : 0);
      constexpr bool a_mn = (MODE == 2), b_mn = (MODE != 0);
      // MN-major operands: tf32 -> SWIZZLE_128B_BASE32B panels (4-row atoms); 2-byte -> plain SWIZZLE_128B panels (8-row
      // atoms, LBO = panel stride, SBO = 1024, 16 K rows per instruction; profiles/r1_umma_mn16_probe.txt)
      constexpr uint32_t a_lbo = a_mn ? kPanel : 16u, b_lbo = b_mn ? kPanel : 16u;
      constexpr uint32_t mn_sbo = ET == 0 ? 512u : 1024u, mn_kstep = ET == 0 ? 1024u : 2048u;
      constexpr uint32_t mn_lay = ET == 0 ? tc::kLayoutSw128Base32 : tc::kLayoutSw128;
      constexpr uint32_t a_sbo = a_mn ? mn_sbo : 1024u, b_sbo = b_mn ? mn_sbo : 1024u;
      constexpr uint32_t a_lay = a_mn ? mn_lay : tc::kLayoutSw128;
      constexpr uint32_t b_lay = b_mn ? mn_lay : tc::kLayoutSw128;
      constexpr uint32_t a_kstep = a_mn ? mn_kstep : 32u, b_kstep = b_mn ? mn_kstep : 32u;
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        tc::mbar_wait(tc::smem_u32(&full_bar[s]), (kb / STAGES) & 1);
        tc::tc_fence_after();
        const uint32_t stage = tiles + s * kStageBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = tc::make_smem_desc(stage + k * a_kstep, a_lbo, a_sbo, a_lay);
          const uint64_t bd = tc::make_smem_desc(stage + kABytes + k * b_kstep, b_lbo, b_sbo, b_lay);
          if (PAIR && ET != 0) tc::umma_f16_2sm(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          else if (PAIR) tc::umma_tf32_2sm(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          else if (ET != 0) tc::umma_f16(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          else tc::umma_tf32(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        if (PAIR) tc::umma_commit_2sm(tc::smem_u32(&empty_bar[s]), (uint16_t)3);
        else if (CL == 1) tc::umma_commit(tc::smem_u32(&empty_bar[s]));
        else tc::umma_commit_mc(tc::smem_u32(&empty_bar[s]), kClMask);
      }
      if (PAIR) tc::umma_commit_2sm(tc::smem_u32(&tmem_full_bar), (uint16_t)3);
      else tc::umma_commit(tc::smem_u32(&tmem_full_bar));
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (kCluster > 1) tc::cluster_sync();   // no CTA exits while a peer can still signal its barriers
  if (warp == 5) {
    if (PAIR) tc::tmem_dealloc2(tmem_base, kTmemCols);
    else tc::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Fixed-order reduction of split-K partials: dw[i] = sum_s part[s][i]. A thread owns two float4 columns (i and i + n4 / 2
// rounded up) and loads up to four partials of each before adding them (memory-level parallelism: with 2-4 splits one
// column per thread left two loads in flight); the additions keep the order s = 0, 1, 2, ... (deterministic).
__global__ void __launch_bounds__(128) splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                             long long n4, int splits, long long stride4) {
  const long long half = (n4 + 1) / 2;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= half) return;
  const long long i1 = i0 + half;
  const bool two = i1 < n4;
  const float4* p0 = reinterpret_cast<const float4*>(part) + i0;
  const float4* p1 = reinterpret_cast<const float4*>(part) + (two ? i1 : i0);
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  int s = 0;
  for (; s + 4 <= splits; s += 4) {
    float4 v[4], w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { v[u] = __ldcs(p0 + (long long)(s + u) * stride4); w[u] = __ldcs(p1 + (long long)(s + u) * stride4); }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a0.x += v[u].x; a0.y += v[u].y; a0.z += v[u].z; a0.w += v[u].w;
      a1.x += w[u].x; a1.y += w[u].y; a1.z += w[u].z; a1.w += w[u].w;
    }
  }
  {
    float4 v[3], w[3];
    const int r = splits - s;                                    // 0 .. 3 remaining partials, all loaded first
#pragma unroll
    for (int u = 0; u < 3; ++u)
      if (u < r) { v[u] = __ldcs(p0 + (long long)(s + u) * stride4); w[u] = __ldcs(p1 + (long long)(s + u) * stride4); }
#pragma unroll
    for (int u = 0; u < 3; ++u)
      if (u < r) {
        a0.x += v[u].x; a0.y += v[u].y; a0.z += v[u].z; a0.w += v[u].w;
        a1.x += w[u].x; a1.y += w[u].y; a1.z += w[u].z; a1.w += w[u].w;
      }
  }
  reinterpret_cast<float4*>(out)[i0] = a0;
  if (two) reinterpret_cast<float4*>(out)[i1] = a1;
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

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

EncodeIm2colFn encode_im2col_fn() {
  static EncodeIm2colFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeIm2colFn>(f);
  }();
  return fn;
}

// im2col-mode map over an NHWC fp32 tensor: 32 channels x `pixels` base pixels per load. Base pixels walk
// [lower, dim - 1 + upper] in w and h with the traversal stride `stride` (= one GEMM row / K index each);
// filter taps are added as instruction offsets. Out-of-bounds reads return zero.
int make_map_im2col(CUtensorMap* m, const float* ptr, int N, int H, int W, int C, int lower_w, int lower_h, int upper_w,
                    int upper_h, int stride, int pixels, bool mn_major) {
  EncodeIm2colFn fn = encode_im2col_fn();
  if (!fn) return MLA_E_NODEVICE;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  int lo[2] = {lower_w, lower_h}, up[2] = {upper_w, upper_h};   // {W, H} order (as the instruction's offsets)
  cuuint32_t estr[4] = {1u, (cuuint32_t)stride, (cuuint32_t)stride, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, lo, up, 32u,
                  (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MLA_E_BADARG;
}

// 2-byte variants (fp16 / bf16: `bf16` selects the TMA data type; the bytes move identically): 64 elements per row.
int make_map_im2col16(CUtensorMap* m, const void* ptr, bool bf16, int N, int H, int W, int C, int lower_w, int lower_h,
                      int upper_w, int upper_h, int stride, int pixels) {
  EncodeIm2colFn fn = encode_im2col_fn();
  if (!fn) return MLA_E_NODEVICE;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lo[2] = {lower_w, lower_h}, up[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1u, (cuuint32_t)stride, (cuuint32_t)stride, 1u};
  CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims,
                  strides, lo, up, 64u, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MLA_E_BADARG;
}

// MLA_CONV_CLUSTER = 1 | 2 | 4: CTAs per weight-multicast cluster in fprop / dgrad.
int conv_cluster() {
  static const int v = [] {
    const char* e = getenv("MLA_CONV_CLUSTER");
    const int c = e ? atoi(e) : 1;
    return (c == 2 || c == 4) ? c : 1;
  }();
  return v;
}

// MLA_CONV_PAIR = 0 | 1 | 2: tcgen05 CTA pairs (cta_group::2) in fprop / dgrad; 2 = also 256-column tiles where the
// channel count allows.
// MLA_CONV_GATHER=1 forces the cp.async gather kernels everywhere (A/B comparison, bring-up).
bool force_gather() {
  static const bool v = [] {
    const char* e = getenv("MLA_CONV_GATHER");
    return e != nullptr && e[0] == '1';
  }();
  return v;
}

// MLA_CONV_PAIR = 1 | 2 | 3 enables tcgen05 CTA pairs (cta_group::2) in fprop / dgrad: 1 = same tile widths, 2 = also
// 256-column tiles, 3 = the measured-best mix (profiles/r1_conv_variants.md: pairs for N tiles >= 128 columns, 256-column
// tiles for exactly 256 channels). OFF BY DEFAULT: the host runtime runs the two encoders (and their weight
// gradients) on concurrent CUDA streams, and two different pair kernels co-resident on one SM pair can dead-lock in
// tcgen05.alloc.cta_group::2 (each holds one SM's allocation while waiting for the other's) — observed once as a hung
// bench on B200. Single-CTA allocations cannot form such a cycle. Pairs are only safe on a single stream.
int conv_pair_env() {
  static const int v = [] {
    const char* e = getenv("MLA_CONV_PAIR");
    return e ? atoi(e) : 0;
  }();
  return v;
}
// tile width for `nch` output columns of the GEMM, or 0 = no pairing
// MLA_LINEAR_PAIR=0 disables CTA pairs in the Linear-layer entry points (default on: they run on a single stream).
bool linear_pair() {
  static const bool v = [] {
    const char* e = getenv("MLA_LINEAR_PAIR");
    return e == nullptr || e[0] != '0';
  }();
  return v && !force_gather();
}

int conv_pair_bn(int nch) {
  const int e = conv_pair_env();
  if (e == 0 || force_gather()) return 0;
  const int bn = (nch % 128 == 0) ? 128 : 64;
  if (e == 1) return bn;
  if (e == 2) return nch % 256 == 0 ? 256 : bn;
  if (nch % 128 != 0) return 0;          // 3: measured-best mix
  return nch == 256 ? 256 : 128;
}

// 2-D fp32 row-major [rows][cols] tensor map with a {32 cols (128 B), box_rows} box; 128B swizzle with
// 16 B chunks (K-major operand tiles) or 32 B chunks (MN-major operand panels).
int make_map_2d(CUtensorMap* m, const float* ptr, long long rows, long long cols, int box_rows, bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return MLA_E_NODEVICE;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MLA_E_BADARG;
}

int make_map_2d16(CUtensorMap* m, const void* ptr, bool bf16, long long rows, long long cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return MLA_E_NODEVICE;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MLA_E_BADARG;
}

template <int MODE, int BN, int STAGES, bool IM2COL, int NT = 1, int CL = 1, bool PAIR = false, int ET = 0>
int launch(const CUtensorMap& map, const CUtensorMap& gmap, const ConvGemmParams& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (kABytes + (PAIR ? BN / 2 : BN * NT) * 128) + 1024;
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    MLA_CUDA_TRY(cudaFuncSetAttribute(conv_gemm_kernel<MODE, BN, STAGES, IM2COL, NT, CL, PAIR, ET>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.store(1, std::memory_order_release);
  }
  constexpr int kCluster = PAIR ? 2 : CL;
  if (kCluster > 1) {
    grid.x = (grid.x + kCluster - 1) / kCluster * kCluster;   // whole clusters; padding CTAs load like the others, store nothing
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    MLA_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<MODE, BN, STAGES, IM2COL, NT, CL, PAIR, ET>, map, gmap, p));
    mla::count_launch();
    return 0;
  }
  conv_gemm_kernel<MODE, BN, STAGES, IM2COL, NT, CL, PAIR, ET><<<grid, kThreads, smem, st>>>(map, gmap, p);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}

// all R x S taps: weights of filter position (i, j), instruction offset (i, j) or the flipped one
void full_taps(ConvGemmParams& p, int R, int S, bool flip) {
  p.nr = R; p.ns = S;
  for (int i = 0; i < R; ++i) { p.tap_r[i] = (signed char)i; p.off_r[i] = (signed char)(flip ? R - 1 - i : i); }
  for (int j = 0; j < S; ++j) { p.tap_s[j] = (signed char)j; p.off_s[j] = (signed char)(flip ? S - 1 - j : j); }
}

// cin_mult: 64 where Cin is tiled as a GEMM N dimension without a tail guard (dgrad), 32 otherwise
bool conv_shape_ok(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int cin_mult = 64) {
  return N > 0 && H > 0 && W > 0 && Cin >= cin_mult && Cin % cin_mult == 0 && Cout >= 64 && Cout % 64 == 0 && R == S &&
         (R == 1 || R == 3 || R == 5 || R == 7) && (stride == 1 || stride == 2) && pad >= 0 && pad <= R / 2;
}

int out_size(int x, int k, int stride, int pad) { return (x + 2 * pad - k) / stride + 1; }

}  // namespace

static int conv2d_fprop_impl(const float* x, const float* w, float* y, int N, int H, int W, int Cin, int Cout, int R, int S,
                            int stride, int pad, float* stat_part, void* stream);

extern "C" int mla_conv2d_fprop(const float* x, const float* w, float* y, int N, int H, int W, int Cin, int Cout, int R,
                                int S, int stride, int pad, void* stream) {
  return conv2d_fprop_impl(x, w, y, N, H, W, Cin, Cout, R, S, stride, pad, nullptr, stream);
}

extern "C" int mla_conv2d_fprop_stat_tiles(int N, int H, int W, int R, int S, int stride, int pad) {
  const int OH = out_size(H, R, stride, pad), OW = out_size(W, S, stride, pad);
  if (N < 1 || OH < 1 || OW < 1) return 0;
  return (int)(((long long)N * OH * OW + 127) / 128);
}

// Rows of the [tile][2][Cout] BatchNorm partial-sum buffer mla_conv2d_fprop16 writes for this geometry (the strip kernel of
// the 64-channel layers tiles by image rows, everything else by 128 output pixels).
extern "C" int mla_conv2d_fprop16_stat_tiles(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad) {
  const int OH = out_size(H, R, stride, pad), OW = out_size(W, S, stride, pad);
  if (N < 1 || OH < 1 || OW < 1) return 0;
  if (!force_gather()) {
    const mla::StripPlan sp = mla::strip16_plan(N, H, W, Cin, Cout, R, S, stride, pad);
    if (sp.tiles > 0) return sp.tiles;
  }
  return (int)(((long long)N * OH * OW + 127) / 128);
}

extern "C" int mla_conv2d_fprop_bnstats(const float* x, const float* w, float* y, int N, int H, int W, int Cin, int Cout,
                                        int R, int S, int stride, int pad, float* stat_part, void* stream) {
  if (!stat_part || !mla::aligned16(stat_part)) return MLA_E_BADARG;
  return conv2d_fprop_impl(x, w, y, N, H, W, Cin, Cout, R, S, stride, pad, stat_part, stream);
}

static int conv2d_fprop_impl(const float* x, const float* w, float* y, int N, int H, int W, int Cin, int Cout, int R, int S,
                            int stride, int pad, float* stat_part, void* stream) {
  if (!x || !w || !y || !mla::aligned16(x) || !mla::aligned16(w) || !mla::aligned16(y)) return MLA_E_BADARG;
  if (!conv_shape_ok(N, H, W, Cin, Cout, R, S, stride, pad, 32)) return MLA_E_SHAPE;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  const int OH = out_size(H, R, stride, pad), OW = out_size(W, S, stride, pad);
  if (OH <= 0 || OW <= 0 || (long long)N * OH * OW > 0x7fffffffLL) return MLA_E_SHAPE;
  ConvGemmParams p{};
  p.src = x; p.Hs = H; p.Ws = W; p.Cs = Cin; p.OH = OH; p.OW = OW; p.M = N * OH * OW; p.R = R; p.S = S;
  p.mul = stride; p.sgn = 1; p.off = -pad; p.div = 1; p.kcb = Cin / 32; p.KB = R * S * p.kcb; p.CinW = Cin;
  p.out = y; p.ldo = Cout; p.accumulate = 0; p.Cout = Cout; p.stat_part = stat_part;
  const int BN = (Cout % 128 == 0) ? 128 : 64;
  CUtensorMap map;
  int rc = make_map_2d(&map, w, Cout, (long long)R * S * Cin, BN, false);
  if (rc) return rc;
  dim3 grid((p.M + 127) / 128, Cout / BN);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (force_gather()) return BN == 64 ? launch<0, 64, 4, false>(map, map, p, grid, st) : launch<0, 128, 3, false>(map, map, p, grid, st);
  // gathered operand: x, base pixel (ow*stride - pad, oh*stride - pad), taps as offsets
  p.g_base_w = p.g_base_h = -pad;
  full_taps(p, R, S, false);
  CUtensorMap gmap;
  rc = make_map_im2col(&gmap, x, N, H, W, Cin, -pad, -pad, pad - (S - 1), pad - (R - 1), stride, 128, false);
  if (rc) return rc;
  if (const int BNp = conv_pair_bn(Cout)) {   // CTA pairs: each CTA holds half of the weight tile -> box {32 k, BNp / 2 rows}
    rc = make_map_2d(&map, w, Cout, (long long)R * S * Cin, BNp / 2, false);
    if (rc) return rc;
    dim3 gp((p.M + 127) / 128, Cout / BNp);
    if (BNp == 256) return launch<0, 256, 3, true, 1, 1, true>(map, gmap, p, gp, st);
    if (BNp == 128) return launch<0, 128, 4, true, 1, 1, true>(map, gmap, p, gp, st);
    return launch<0, 64, 4, true, 1, 1, true>(map, gmap, p, gp, st);
  }
  const int cl = conv_cluster();
  if (cl > 1) {   // weight tile split over the cluster: box {32 k, BN / cl rows}
    rc = make_map_2d(&map, w, Cout, (long long)R * S * Cin, BN / cl, false);
    if (rc) return rc;
    if (cl == 2) return BN == 64 ? launch<0, 64, 4, true, 1, 2>(map, gmap, p, grid, st) : launch<0, 128, 3, true, 1, 2>(map, gmap, p, grid, st);
    return BN == 64 ? launch<0, 64, 4, true, 1, 4>(map, gmap, p, grid, st) : launch<0, 128, 3, true, 1, 4>(map, gmap, p, grid, st);
  }
  static const int occ3 = [] { const char* e = getenv("MLA_CONV_OCC3"); return e ? atoi(e) : 0; }();
  if (occ3) return BN == 64 ? launch<0, 64, 3, true>(map, gmap, p, grid, st) : launch<0, 128, 2, true>(map, gmap, p, grid, st);
  return BN == 64 ? launch<0, 64, 4, true>(map, gmap, p, grid, st) : launch<0, 128, 3, true>(map, gmap, p, grid, st);
}

namespace {
// im2col dgrad launch with the configured weight-multicast cluster (64-column tiles have 2 weight panels: <= 2)
int launch_dgrad(int BN, const CUtensorMap& map, const CUtensorMap& gmap, const ConvGemmParams& p, dim3 grid,
                 cudaStream_t st) {
  if (const int BNp = conv_pair_bn(p.CinW)) {
    grid.y = p.CinW / BNp;
    if (BNp == 256) return launch<1, 256, 3, true, 1, 1, true>(map, gmap, p, grid, st);
    if (BNp == 128) return launch<1, 128, 4, true, 1, 1, true>(map, gmap, p, grid, st);
    return launch<1, 64, 4, true, 1, 1, true>(map, gmap, p, grid, st);
  }
  const int cl = conv_cluster();
  if (BN == 64) {
    if (cl >= 2) return launch<1, 64, 4, true, 1, 2>(map, gmap, p, grid, st);
    return launch<1, 64, 4, true>(map, gmap, p, grid, st);
  }
  if (cl == 2) return launch<1, 128, 3, true, 1, 2>(map, gmap, p, grid, st);
  if (cl == 4) return launch<1, 128, 3, true, 1, 4>(map, gmap, p, grid, st);
  return launch<1, 128, 3, true>(map, gmap, p, grid, st);
}
}  // namespace

extern "C" int mla_conv2d_dgrad(const float* dy, const float* w, float* dx, int N, int H, int W, int Cin, int Cout,
                                int R, int S, int stride, int pad, int accumulate, void* stream) {
  if (!dy || !w || !dx || !mla::aligned16(dy) || !mla::aligned16(w) || !mla::aligned16(dx)) return MLA_E_BADARG;
  if (!conv_shape_ok(N, H, W, Cin, Cout, R, S, stride, pad)) return MLA_E_SHAPE;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  const int OH = out_size(H, R, stride, pad), OW = out_size(W, S, stride, pad);
  if (OH <= 0 || OW <= 0 || (long long)N * H * W > 0x7fffffffLL) return MLA_E_SHAPE;
  ConvGemmParams p{};
  p.src = dy; p.Hs = OH; p.Ws = OW; p.Cs = Cout; p.OH = H; p.OW = W; p.M = N * H * W; p.R = R; p.S = S;
  p.mul = 1; p.sgn = -1; p.off = pad; p.div = stride; p.kcb = Cout / 32; p.KB = R * S * p.kcb; p.CinW = Cin;
  p.out = dx; p.ldo = Cin; p.accumulate = accumulate ? 1 : 0; p.Cout = Cout;
  const int BN = (Cin % 128 == 0) ? 128 : 64;
  CUtensorMap map;
  int rc = make_map_2d(&map, w, Cout, (long long)R * S * Cin, 32, true);
  if (rc) return rc;
  dim3 grid((p.M + 127) / 128, Cin / BN);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (force_gather())
    return BN == 64 ? launch<1, 64, 4, false>(map, map, p, grid, st) : launch<1, 128, 3, false>(map, map, p, grid, st);
  if (stride == 1) {
    // dgrad is a convolution of dy with the flipped filter: base pixel (w + pad - (S-1), h + pad - (R-1)),
    // weight tap (r, s) pairs with filter offset (R-1-r, S-1-s)
    p.g_base_w = pad - (S - 1); p.g_base_h = pad - (R - 1);
    full_taps(p, R, S, true);
    CUtensorMap gmap;
    rc = make_map_im2col(&gmap, dy, N, OH, OW, Cout, p.g_base_w, p.g_base_h, p.g_base_w + (W - OW), p.g_base_h + (H - OH),
                         1, 128, false);
    if (rc) return rc;
    return launch_dgrad(BN, map, gmap, p, grid, st);
  }
  // stride s: one dense sub-convolution per output parity class (ph, pw). dx[n, i*s+ph, j*s+pw] sums the taps r
  // with (ph + pad - r) % s == 0, reading dy row i + (ph + pad - r) / s — a stride-1 walk over dy, so the im2col
  // map applies; rows are scattered to their dx pixels by the epilogue.
  for (int ph = 0; ph < stride; ++ph) {
    for (int pw = 0; pw < stride; ++pw) {
      ConvGemmParams q = p;
      const int Hs = (H - ph + stride - 1) / stride, Ws = (W - pw + stride - 1) / stride;
      if (Hs <= 0 || Ws <= 0) continue;
      q.nr = q.ns = 0;
      int qr[8], qs[8], lo_h = 1 << 20, lo_w = 1 << 20;
      for (int r = 0; r < R; ++r) {
        const int t = ph + pad - r;
        if (((t % stride) + stride) % stride != 0) continue;
        qr[q.nr] = (t >= 0 ? t : t - (stride - 1)) / stride;
        q.tap_r[q.nr] = (signed char)r;
        lo_h = min(lo_h, qr[q.nr]);
        ++q.nr;
      }
      for (int c = 0; c < S; ++c) {
        const int t = pw + pad - c;
        if (((t % stride) + stride) % stride != 0) continue;
        qs[q.ns] = (t >= 0 ? t : t - (stride - 1)) / stride;
        q.tap_s[q.ns] = (signed char)c;
        lo_w = min(lo_w, qs[q.ns]);
        ++q.ns;
      }
      q.OH = Hs; q.OW = Ws; q.M = N * Hs * Ws;
      q.o_mul = stride; q.o_ph = ph; q.o_pw = pw; q.o_H = H; q.o_W = W;
      q.KB = q.nr * q.ns * q.kcb;
      dim3 g((q.M + 127) / 128, Cin / BN);
      if (q.KB == 0) {
        if (accumulate) continue;              // nothing to add to these pixels
        lo_h = lo_w = 0;                       // KB == 0: the kernel only writes zeros
      }
      for (int i = 0; i < q.nr; ++i) q.off_r[i] = (signed char)(qr[i] - lo_h);
      for (int j = 0; j < q.ns; ++j) q.off_s[j] = (signed char)(qs[j] - lo_w);
      q.g_base_w = lo_w; q.g_base_h = lo_h;
      CUtensorMap gmap;
      rc = make_map_im2col(&gmap, dy, N, OH, OW, Cout, lo_w, lo_h, lo_w + Ws - OW, lo_h + Hs - OH, 1, 128, false);
      if (rc) return rc;
      rc = launch_dgrad(BN, map, gmap, q, g, st);
      if (rc) return rc;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent 2-byte implicit-GEMM convolution (the fprop-type launches: fprop16 and dgrad16): the same producer / MMA /
// epilogue roles as conv_gemm_kernel<0, BN, STAGES, true, 1, 1, false, ET>, but a CTA walks several 128 x BN output tiles and
// the accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue of tile i (fp32 stores, the accumulating read
// of dgrad, the BatchNorm partial sums of fprop) overlaps the im2col TMA + MMA main loop of tile i + 1. KB > 0 required.
namespace {
template <int BN, int STAGES, int ET>
__global__ void __launch_bounds__(kThreads) conv16_persistent_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                     const __grid_constant__ CUtensorMap tmap_g, ConvGemmParams p,
                                                                     int tiles_m, int tiles) {
  constexpr uint32_t kBBytes = BN * 128, kStageBytes = kABytes + kBBytes;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t acc_full_bar[2];
  __shared__ __align__(8) uint64_t acc_empty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_stat[4 * 2 * BN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_tiles = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
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
    tc::tma_prefetch_desc(&tmap);
    tc::tma_prefetch_desc(&tmap_g);
  }
  if (warp == 5) {
    tc::tmem_alloc(tc::smem_u32(&tmem_slot), 2 * BN);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int KB = p.KB;

  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int m0 = (tile % tiles_m) * 128, n0 = (tile / tiles_m) * BN;
        const int ow = m0 % p.OW, t = m0 / p.OW;
        const int gw = ow * p.mul + p.g_base_w, gh = (t % p.OH) * p.mul + p.g_base_h, gn = t / p.OH;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ((it / STAGES) & 1) ^ 1);
          const uint32_t stage = smem_tiles + s * kStageBytes;
          const uint32_t bar = tc::smem_u32(&full_bar[s]);
          const int tap = kb / p.kcb, cb = kb - tap * p.kcb;
          const int ti = tap / p.ns, tj = tap - ti * p.ns;
          tc::mbar_arrive_expect_tx(bar, kStageBytes);
          tc::tma_load_im2col_4d(stage, &tmap_g, bar, cb * 64, gw, gh, gn, (uint16_t)p.off_s[tj], (uint16_t)p.off_r[ti]);
          tc::tma_load_2d(stage + kABytes, &tmap, bar, (p.tap_r[ti] * p.S + p.tap_s[tj]) * p.CinW + cb * 64, n0);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_f16(128, BN, ET >= 2 ? 1 : 0, ET == 2 ? 1 : 0, 0, 0);
      int it = 0, nt = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++nt) {
        const int buf = nt & 1;
        tc::mbar_wait(tc::smem_u32(&acc_empty_bar[buf]), ((nt >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)buf * BN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          tc::mbar_wait(tc::smem_u32(&full_bar[s]), (it / STAGES) & 1);
          tc::tc_fence_after();
          const uint32_t stage = smem_tiles + s * kStageBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = tc::make_smem_desc(stage + k * 32u, 16u, 1024u, tc::kLayoutSw128);
            const uint64_t bd = tc::make_smem_desc(stage + kABytes + k * 32u, 16u, 1024u, tc::kLayoutSw128);
            tc::umma_f16(acc, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc::umma_commit(tc::smem_u32(&empty_bar[s]));
        }
        tc::umma_commit(tc::smem_u32(&acc_full_bar[buf]));
      }
    }
  } else {
    const float oscale = p.out_scale != nullptr ? __ldg(p.out_scale) : 1.f;   // undoes the power-of-two scale of fp16 gradients
    int nt = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++nt) {
      const int buf = nt & 1;
      const int bx = tile % tiles_m;
      const int m0 = bx * 128, n0 = (tile / tiles_m) * BN;
      tc::mbar_wait(tc::smem_u32(&acc_full_bar[buf]), (nt >> 1) & 1);
      tc::tc_fence_after();
      const int m = m0 + warp * 32 + lane;
      float* orow = nullptr;
      if (m < p.M) {
        long long orow_idx = m;
        if (p.o_mul > 1) {
          const int j = m % p.OW;
          const int t = m / p.OW;
          orow_idx = ((long long)(t / p.OH) * p.o_H + (t % p.OH) * p.o_mul + p.o_ph) * p.o_W + j * p.o_mul + p.o_pw;
        }
        orow = p.out + orow_idx * p.ldo + n0;
      }
      const uint32_t acc = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tc::tmem_ld32(acc + c, v);
        tc::tmem_ld_wait();
        if (orow != nullptr) {
          float4* dst = reinterpret_cast<float4*>(orow + c);
          float4 o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
          if (p.out_scale != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j].x *= oscale; o[j].y *= oscale; o[j].z *= oscale; o[j].w *= oscale; }
          }
          if (p.accumulate) {
            float4 old[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) old[j] = dst[j];
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j].x += old[j].x; o[j].y += old[j].y; o[j].z += old[j].z; o[j].w += old[j].w; }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = o[j];
        }
        if (p.stat_part != nullptr) {      // BatchNorm partial sums of this warp's 32 rows (see conv_gemm_kernel)
          float a[32], b[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { a[j] = __uint_as_float(v[j]); b[j] = a[j] * a[j]; }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
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
      // the accumulator buffer has been read completely: hand it back before the (slower) statistics write-out
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty_bar[buf]));
      if (p.stat_part != nullptr) {
        asm volatile("bar.sync 1, 128;" ::: "memory");   // the 4 epilogue warps only
        for (int t = threadIdx.x; t < BN; t += 128) {
          float sa = 0.f, sb = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) { sa += s_stat[(w * 2 + 0) * BN + t]; sb += s_stat[(w * 2 + 1) * BN + t]; }
          float* dstp = p.stat_part + (size_t)bx * 2 * p.Cout + n0 + t;
          dstp[0] = sa;
          dstp[p.Cout] = sb;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // s_stat is rewritten by the next tile
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tmem_base, 2 * BN);
}

// ------------------------------------------------------------------------------------------------------------------
// The same persistent scheme on CTA PAIRS (tcgen05 cta_group::2): a pair walks 256 x BN output tiles (BN up to 256). Each CTA
// gathers its own 128 im2col rows and stages HALF of the weight tile; the leader issues one 256 x BN x 16 MMA per k-step.
// Why: these convolutions are bound by the L2 -> SM operand stream (~10.5 TB/s measured: 450 TF/s at 44 FLOP/B on the 64-channel
// layers, 683 TF/s at 64 FLOP/B on the 256-channel ones, profiles/r1_conv_variants.md), and a 256 x 256 pair tile moves half the
// operand bytes per FLOP of two 128 x 128 tiles (256 x 128: 3/4; 256 x 64: 5/6).
// Co-residency: a CTA of this kernel takes >= 180 KB of shared memory, i.e. a whole SM (and its cluster the whole TPC), so no other
// TMEM-allocating CTA can sit next to it and tcgen05.alloc.cta_group::2 can never wait on a partner that waits on it — the
// dead-lock the 2-CTA-per-SM pair variants of conv_gemm_kernel could run into under concurrent streams.
template <int BN>
struct PairCfg {
  static constexpr uint32_t kB = (BN / 2) * 128;                 // this CTA's half of the weight tile, bytes per stage
  static constexpr uint32_t kStage = kABytes + kB;
  static constexpr int kStages = BN == 256 ? 6 : (BN == 128 ? 8 : 9);
  static constexpr size_t kSmem = (size_t)kStages * kStage + 1024;
};

template <int BN, int ET>
__global__ void __launch_bounds__(kThreads, 1) conv16_pair_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                  const __grid_constant__ CUtensorMap tmap_g, ConvGemmParams p,
                                                                  int tiles_m, int tiles) {
  using Cfg = PairCfg<BN>;
  constexpr int STAGES = Cfg::kStages;
  constexpr uint32_t kStageBytes = Cfg::kStage;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t acc_full_bar[2];
  __shared__ __align__(8) uint64_t acc_empty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_stat[4 * 2 * BN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_tiles = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t rank = tc::cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(tc::smem_u32(&full_bar[s]), 1);
      tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(tc::smem_u32(&acc_full_bar[b]), 1);
      tc::mbar_init(tc::smem_u32(&acc_empty_bar[b]), 8);     // 4 epilogue warps of each CTA of the pair
    }
    tc::fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tc::tma_prefetch_desc(&tmap);
    tc::tma_prefetch_desc(&tmap_g);
  }
  if (warp == 5) {
    tc::tmem_alloc2(tc::smem_u32(&tmem_slot), 2 * BN);
    tc::tmem_relinquish2();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int KB = p.KB;
  const int cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      for (int tile = cluster; tile < tiles; tile += nclusters) {
        const int m0 = (tile % tiles_m) * 256 + (int)rank * 128, n0 = (tile / tiles_m) * BN + (int)rank * (BN / 2);
        const int ow = m0 % p.OW, t = m0 / p.OW;
        const int gw = ow * p.mul + p.g_base_w, gh = (t % p.OH) * p.mul + p.g_base_h, gn = t / p.OH;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ((it / STAGES) & 1) ^ 1);
          const uint32_t stage = smem_tiles + s * kStageBytes;
          const uint32_t bar = tc::smem_u32(&full_bar[s]);
          const int tap = kb / p.kcb, cb = kb - tap * p.kcb;
          const int ti = tap / p.ns, tj = tap - ti * p.ns;
          if (rank == 0) tc::mbar_arrive_expect_tx(bar, 2u * kStageBytes);          // both CTAs' bytes land on the leader
          tc::tma_load_im2col_4d_2sm(stage, &tmap_g, bar, cb * 64, gw, gh, gn, (uint16_t)p.off_s[tj], (uint16_t)p.off_r[ti]);
          tc::tma_load_2d_2sm(stage + kABytes, &tmap, bar, (p.tap_r[ti] * p.S + p.tap_s[tj]) * p.CinW + cb * 64, n0);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = tc::make_idesc_f16(256, BN, ET >= 2 ? 1 : 0, ET == 2 ? 1 : 0, 0, 0);
      int it = 0, nt = 0;
      for (int tile = cluster; tile < tiles; tile += nclusters, ++nt) {
        const int buf = nt & 1;
        tc::mbar_wait(tc::smem_u32(&acc_empty_bar[buf]), ((nt >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)buf * BN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          tc::mbar_wait(tc::smem_u32(&full_bar[s]), (it / STAGES) & 1);
          tc::tc_fence_after();
          const uint32_t stage = smem_tiles + s * kStageBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = tc::make_smem_desc(stage + k * 32u, 16u, 1024u, tc::kLayoutSw128);
            const uint64_t bd = tc::make_smem_desc(stage + kABytes + k * 32u, 16u, 1024u, tc::kLayoutSw128);
            tc::umma_f16_2sm(acc, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc::umma_commit_2sm(tc::smem_u32(&empty_bar[s]), (uint16_t)3);
        }
        tc::umma_commit_2sm(tc::smem_u32(&acc_full_bar[buf]), (uint16_t)3);
      }
    }
  } else {
    const float oscale = p.out_scale != nullptr ? __ldg(p.out_scale) : 1.f;
    int nt = 0;
    for (int tile = cluster; tile < tiles; tile += nclusters, ++nt) {
      const int buf = nt & 1;
      const int bx = (tile % tiles_m) * 2 + (int)rank;             // 128-row tile index (BatchNorm partials are per 128 rows)
      const int m0 = bx * 128, n0 = (tile / tiles_m) * BN;
      tc::mbar_wait(tc::smem_u32(&acc_full_bar[buf]), (nt >> 1) & 1);
      tc::tc_fence_after();
      const int m = m0 + warp * 32 + lane;
      float* orow = nullptr;
      if (m < p.M) {
        long long orow_idx = m;
        if (p.o_mul > 1) {
          const int j = m % p.OW;
          const int t = m / p.OW;
          orow_idx = ((long long)(t / p.OH) * p.o_H + (t % p.OH) * p.o_mul + p.o_ph) * p.o_W + j * p.o_mul + p.o_pw;
        }
        orow = p.out + orow_idx * p.ldo + n0;
      }
      const uint32_t acc = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tc::tmem_ld32(acc + c, v);
        tc::tmem_ld_wait();
        if (orow != nullptr) {
          float4* dst = reinterpret_cast<float4*>(orow + c);
          float4 o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
          if (p.out_scale != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j].x *= oscale; o[j].y *= oscale; o[j].z *= oscale; o[j].w *= oscale; }
          }
          if (p.accumulate) {
            float4 old[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) old[j] = dst[j];
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j].x += old[j].x; o[j].y += old[j].y; o[j].z += old[j].z; o[j].w += old[j].w; }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = o[j];
        }
        if (p.stat_part != nullptr) {      // BatchNorm partial sums of this warp's 32 rows (see conv_gemm_kernel)
          float a[32], b[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { a[j] = __uint_as_float(v[j]); b[j] = a[j] * a[j]; }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
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
      // the accumulator buffer has been read completely: hand it back to the leader's MMA warp
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_leader(tc::smem_u32(&acc_empty_bar[buf]));
      if (p.stat_part != nullptr) {
        asm volatile("bar.sync 1, 128;" ::: "memory");   // the 4 epilogue warps only
        if (m0 < p.M) {                                   // a 128-row tile entirely past M (odd tile count) has no partial row
          for (int t = threadIdx.x; t < BN; t += 128) {
            float sa = 0.f, sb = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) { sa += s_stat[(w * 2 + 0) * BN + t]; sb += s_stat[(w * 2 + 1) * BN + t]; }
            float* dstp = p.stat_part + (size_t)bx * 2 * p.Cout + n0 + t;
            dstp[0] = sa;
            dstp[p.Cout] = sb;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // s_stat is rewritten by the next tile
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();                       // no CTA exits while its peer can still signal its barriers / read its smem
  if (warp == 5) tc::tmem_dealloc2(tmem_base, 2 * BN);
}

// MLA_CONV_PAIR16=1 routes fprop16 / dgrad16 through this kernel. OFF by default — measured on the ResNet-18 shapes of the
// bench batch (profiles/r2_conv_pair.md): 256 x 256 pair tiles tie with the single-CTA kernel on the 256 / 512-channel layers
// (0.049 vs 0.047 ms: 98 tiles on 74 clusters = two rounds, half of the second one idle) and lose on the 64 / 128-channel
// ones (0.100 vs 0.081 ms, 0.062 vs 0.055 ms), where one CTA per SM cannot hide the accumulator hand-back across the
// pair. The single-CTA kernel already runs at 82-88 % of the chip's L2 -> SM throughput cap on these layers (ncu: 693 MB of
// TMA loads in 66 us on 56x56x64), so the lever the pairs pull (operand bytes) is the right one; what they lack is
// scheduling granularity. Kept selectable, and covered by tests/test_gpu_encoder_kernels.py::test_conv_pair_kernel.
bool conv_pair16() {
  static const bool v = [] {
    const char* e = getenv("MLA_CONV_PAIR16");
    return e != nullptr && e[0] == '1';
  }();
  return v && !force_gather();
}

template <int BN, int ET>
int launch_conv16_pair(const void* w16, bool bf16, long long wrows, long long wcols, const CUtensorMap& gmap,
                       const ConvGemmParams& p, int n_tiles_n, cudaStream_t st) {
  using Cfg = PairCfg<BN>;
  CUtensorMap map;
  int rc = make_map_2d16(&map, w16, bf16, wrows, wcols, BN / 2);          // box {64 k, BN / 2 rows}: one CTA's half
  if (rc) return rc;
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    MLA_CUDA_TRY(cudaFuncSetAttribute(conv16_pair_kernel<BN, ET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
    configured.store(1, std::memory_order_release);
  }
  const int tiles_m = (p.M + 255) / 256, tiles = tiles_m * n_tiles_n;
  const int nclusters = std::max(1, std::min(tiles, mla::device_info().sm_count / 2));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * nclusters); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Cfg::kSmem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  MLA_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv16_pair_kernel<BN, ET>, map, gmap, p, tiles_m, tiles));
  mla::count_launch();
  return 0;
}

// tile width of the pair kernel for `nch` output channels (0 = not applicable)
int pair16_bn(int nch, int M, int KB) {
  if (!conv_pair16() || KB <= 0 || M < 512) return 0;
  return nch % 256 == 0 ? 256 : (nch % 128 == 0 ? 128 : 64);
}

// MLA_CONV_PERSIST=0 goes back to one tile per CTA for fprop16 / dgrad16. Default on: fprop16 301 -> 381 TF/s, dgrad16
// 316 -> 378 TF/s, ResNet step 9.6 -> 8.8 ms (profiles/runs/f8_bench_*.json).
bool conv_persist() {
  static const bool v = [] {
    const char* e = getenv("MLA_CONV_PERSIST");
    return e != nullptr ? e[0] == '1' : true;
  }();
  return v && !force_gather();
}

template <int BN, int STAGES, int ET>
int launch_conv16_persistent(const CUtensorMap& map, const CUtensorMap& gmap, const ConvGemmParams& p, int n_tiles_n,
                             cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (kABytes + BN * 128) + 1024;
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    MLA_CUDA_TRY(cudaFuncSetAttribute(conv16_persistent_kernel<BN, STAGES, ET>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    configured.store(1, std::memory_order_release);
  }
  const int tiles_m = (p.M + 127) / 128, tiles = tiles_m * n_tiles_n;
  static const int per_sm = [] { const char* e = getenv("MLA_CONV_PERSIST_CTAS"); return e ? std::max(1, atoi(e)) : 2; }();
  const int grid = std::min(tiles, per_sm * mla::device_info().sm_count);  // two CTAs per SM fit (96 KB of stages each)
  conv16_persistent_kernel<BN, STAGES, ET><<<grid, kThreads, smem, st>>>(map, gmap, p, tiles_m, tiles);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}
}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// 2-byte operand convolutions (kind::f16, fp32 accumulate). fprop16: x fp16, w fp16 [Cout][R][S][Cin]. dgrad16: dy bf16,
// wt bf16 = the TRANSPOSED filter [Cin][R][S][Cout], which makes dgrad the same K-major GEMM as fprop (flipped taps,
// stride-2 by output parity classes) — no MN-major operand. fp16 has TF32's 10-bit mantissa, so fprop16 multiplies
// the same operand values as the TF32 path; dy in bf16 (8-bit mantissa, fp32 range) sits far below the ~10 % noise TF32
// ReLU-mask flips already put on the encoder gradients.
extern "C" int mla_conv2d_fprop16(const void* x16, const void* w16, float* y, int N, int H, int W, int Cin, int Cout, int R,
                                  int S, int stride, int pad, float* stat_part, void* stream) {
  if (!x16 || !w16 || !y || !mla::aligned16(x16) || !mla::aligned16(w16) || !mla::aligned16(y)) return MLA_E_BADARG;
  if (!conv_shape_ok(N, H, W, Cin, Cout, R, S, stride, pad, 64)) return MLA_E_SHAPE;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  const int OH = out_size(H, R, stride, pad), OW = out_size(W, S, stride, pad);
  if (OH <= 0 || OW <= 0 || (long long)N * OH * OW > 0x7fffffffLL) return MLA_E_SHAPE;
  if (!force_gather()) {   // 64 -> 64, 3x3 / stride 1: halo-resident strips, weights resident in shared memory
    const mla::StripPlan sp = mla::strip16_plan(N, H, W, Cin, Cout, R, S, stride, pad);
    if (sp.tiles > 0) return mla::conv_strip16_run(0, x16, w16, y, N, H, W, 0, stat_part, nullptr, sp, stream);
  }
  ConvGemmParams p{};
  p.OH = OH; p.OW = OW; p.M = N * OH * OW; p.R = R; p.S = S; p.mul = stride;
  p.kcb = Cin / 64; p.KB = R * S * p.kcb; p.CinW = Cin;
  p.out = y; p.ldo = Cout; p.accumulate = 0; p.Cout = Cout; p.stat_part = stat_part;
  p.g_base_w = p.g_base_h = -pad;
  full_taps(p, R, S, false);
  const int BN = (Cout % 128 == 0) ? 128 : 64;
  CUtensorMap map, gmap;
  int rc = make_map_2d16(&map, w16, false, Cout, (long long)R * S * Cin, BN);
  if (rc) return rc;
  rc = make_map_im2col16(&gmap, x16, false, N, H, W, Cin, -pad, -pad, pad - (S - 1), pad - (R - 1), stride, 128);
  if (rc) return rc;
  dim3 grid((p.M + 127) / 128, Cout / BN);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (const int bnp = pair16_bn(Cout, p.M, p.KB)) {
    const long long wc = (long long)R * S * Cin;
    return bnp == 256 ? launch_conv16_pair<256, 1>(w16, false, Cout, wc, gmap, p, Cout / 256, st)
         : bnp == 128 ? launch_conv16_pair<128, 1>(w16, false, Cout, wc, gmap, p, Cout / 128, st)
                      : launch_conv16_pair<64, 1>(w16, false, Cout, wc, gmap, p, Cout / 64, st);
  }
  if (conv_persist() && p.KB > 0)
    return BN == 64 ? launch_conv16_persistent<64, 4, 1>(map, gmap, p, Cout / BN, st)
                    : launch_conv16_persistent<128, 3, 1>(map, gmap, p, Cout / BN, st);
  return BN == 64 ? launch<0, 64, 4, true, 1, 1, false, 1>(map, gmap, p, grid, st)
                  : launch<0, 128, 3, true, 1, 1, false, 1>(map, gmap, p, grid, st);
}

// y [M, N] = x16 [M, K] (fp16) w16 [N, K]^T (fp16) + bias [N] + resid [M, N]: a Linear layer = the 1x1 case of fprop16
// over an N=1 image of M x 1 pixels, with the bias / residual adds in the epilogue. K, N % 64 == 0.
// ------------------------------------------------------------------------------------------------------------------
// Persistent Linear GEMM (fp16 x fp16 -> fp32): y [M, N] = *out_scale * x16 [M, K] w16 [N, K]^T + bias + resid.
// One CTA per SM walks 128 x 256 output tiles (m fastest, so the CTAs of a wave share the weight tile in L2). The
// accumulator is double-buffered in TMEM (2 x 256 of the 512 columns): the four epilogue warps drain tile i (tcgen05.ld ->
// scale / bias / residual -> fp32 stores) while the MMA warp already accumulates tile i + 1 — at K = 768 the epilogue of a
// tile costs as much as its main loop, and in the one-tile-per-CTA kernel above the two never overlap inside a CTA.
// warp 4 = TMA producer (2-D tiled maps, SWIZZLE_128B, 4 stages of 16 + 32 KB), warp 5 = MMA issuer, warps 0-3 = epilogue.
namespace {
constexpr int kLinStages = 4;
constexpr int kLinBN = 256;
constexpr uint32_t kLinABytes = 128 * 128, kLinBBytes = kLinBN * 128, kLinStageBytes = kLinABytes + kLinBBytes;

struct LinearParams {
  float* out;
  const float* bias;
  const float* resid;
  const float* out_scale;
  int M, N, KB, tiles_m, tiles;
};

__global__ void __launch_bounds__(kThreads, 1) linear_gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                             const __grid_constant__ CUtensorMap tmap_b,
                                                                             LinearParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kLinStages];
  __shared__ __align__(8) uint64_t empty_bar[kLinStages];
  __shared__ __align__(8) uint64_t acc_full_bar[2];
  __shared__ __align__(8) uint64_t acc_empty_bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kLinStages; ++s) {
      tc::mbar_init(tc::smem_u32(&full_bar[s]), 1);
      tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(tc::smem_u32(&acc_full_bar[b]), 1);
      tc::mbar_init(tc::smem_u32(&acc_empty_bar[b]), 4);     // one arrive per epilogue warp
    }
    tc::fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_a);
    tc::tma_prefetch_desc(&tmap_b);
  }
  if (warp == 5) {
    tc::tmem_alloc(tc::smem_u32(&tmem_slot), 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int KB = p.KB;

  if (warp == 4) {
    if (lane == 0) {
      int it = 0;                                           // k-blocks issued so far (ring position)
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const int m0 = (tile % p.tiles_m) * 128, n0 = (tile / p.tiles_m) * kLinBN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % kLinStages;
          tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ((it / kLinStages) & 1) ^ 1);
          const uint32_t stage = tiles + s * kLinStageBytes;
          const uint32_t bar = tc::smem_u32(&full_bar[s]);
          tc::mbar_arrive_expect_tx(bar, kLinStageBytes);
          tc::tma_load_2d(stage, &tmap_a, bar, kb * 64, m0);                 // box {64 k, 128 rows}
          tc::tma_load_2d(stage + kLinABytes, &tmap_b, bar, kb * 64, n0);    // box {64 k, 256 rows}
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_f16(128, kLinBN, 0, 0, 0, 0);
      int it = 0, nt = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++nt) {
        const int buf = nt & 1;
        tc::mbar_wait(tc::smem_u32(&acc_empty_bar[buf]), ((nt >> 1) & 1) ^ 1);   // the epilogue has drained this buffer
        tc::tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)buf * kLinBN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % kLinStages;
          tc::mbar_wait(tc::smem_u32(&full_bar[s]), (it / kLinStages) & 1);
          tc::tc_fence_after();
          const uint32_t stage = tiles + s * kLinStageBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = tc::make_smem_desc(stage + k * 32u, 16u, 1024u, tc::kLayoutSw128);
            const uint64_t bd = tc::make_smem_desc(stage + kLinABytes + k * 32u, 16u, 1024u, tc::kLayoutSw128);
            tc::umma_f16(acc, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc::umma_commit(tc::smem_u32(&empty_bar[s]));
        }
        tc::umma_commit(tc::smem_u32(&acc_full_bar[buf]));
      }
    }
  } else {
    // ===================== epilogue warps 0-3: TMEM lanes warp * 32 .. + 31 = rows of the tile =====================
    const float oscale = p.out_scale != nullptr ? __ldg(p.out_scale) : 1.f;
    int nt = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++nt) {
      const int buf = nt & 1;
      const int m0 = (tile % p.tiles_m) * 128, n0 = (tile / p.tiles_m) * kLinBN;
      tc::mbar_wait(tc::smem_u32(&acc_full_bar[buf]), (nt >> 1) & 1);
      tc::tc_fence_after();
      const int row = m0 + warp * 32 + lane;
      float* orow = row < p.M ? p.out + (long long)row * p.N + n0 : nullptr;
      const float* rrow = (p.resid != nullptr && row < p.M) ? p.resid + (long long)row * p.N + n0 : nullptr;
      const uint32_t acc = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * kLinBN;
#pragma unroll 1
      for (int c = 0; c < kLinBN; c += 32) {
        uint32_t v[32];
        tc::tmem_ld32(acc + c, v);
        tc::tmem_ld_wait();
        if (orow != nullptr) {
          float4 o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
          if (p.out_scale != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j].x *= oscale; o[j].y *= oscale; o[j].z *= oscale; o[j].w *= oscale; }
          }
          if (rrow != nullptr) {
            float4 rv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[j] = reinterpret_cast<const float4*>(rrow + c)[j];
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j].x += rv[j].x; o[j].y += rv[j].y; o[j].z += rv[j].z; o[j].w += rv[j].w; }
          }
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c) + j);
              o[j].x += bv.x; o[j].y += bv.y; o[j].z += bv.z; o[j].w += bv.w;
            }
          }
          float4* dst = reinterpret_cast<float4*>(orow + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = o[j];
        }
      }
      // every tcgen05.ld of this buffer has completed (wait::ld above): hand it back to the MMA warp
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty_bar[buf]));
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tmem_base, 512);
}

// The same persistent scheme on CTA PAIRS (cta_group::2): a pair walks 256 x 256 tiles; each CTA stages its own 128 rows
// of x and HALF of the weight tile (32 KB per k-block instead of 48 for the same MMA work: the operand stream from L2 is
// what bounds these GEMMs), the leader issues one 256 x 256 x 16 MMA per k-step, both CTAs drain their 128 accumulator
// rows. Accumulator hand-back: all 8 epilogue warps of the pair arrive on the LEADER's barrier.
constexpr int kPairStages = 6;
constexpr uint32_t kPairStageBytes = 128 * 128 + 128 * 128;   // 128 rows of x + 128 rows of w

__global__ void __launch_bounds__(kThreads, 1) linear_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                       const __grid_constant__ CUtensorMap tmap_b,
                                                                       LinearParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kPairStages];
  __shared__ __align__(8) uint64_t empty_bar[kPairStages];
  __shared__ __align__(8) uint64_t acc_full_bar[2];
  __shared__ __align__(8) uint64_t acc_empty_bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t rank = tc::cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kPairStages; ++s) {
      tc::mbar_init(tc::smem_u32(&full_bar[s]), 1);
      tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(tc::smem_u32(&acc_full_bar[b]), 1);
      tc::mbar_init(tc::smem_u32(&acc_empty_bar[b]), 8);     // 4 epilogue warps of each CTA of the pair
    }
    tc::fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_a);
    tc::tma_prefetch_desc(&tmap_b);
  }
  if (warp == 5) {
    tc::tmem_alloc2(tc::smem_u32(&tmem_slot), 512);
    tc::tmem_relinquish2();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int KB = p.KB;
  const int cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      for (int tile = cluster; tile < p.tiles; tile += nclusters) {
        const int m0 = (tile % p.tiles_m) * 256 + (int)rank * 128, n0 = (tile / p.tiles_m) * kLinBN + (int)rank * 128;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % kPairStages;
          tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ((it / kPairStages) & 1) ^ 1);
          const uint32_t stage = tiles + s * kPairStageBytes;
          const uint32_t bar = tc::smem_u32(&full_bar[s]);
          if (rank == 0) tc::mbar_arrive_expect_tx(bar, 2u * kPairStageBytes);       // both CTAs' bytes land on the leader
          tc::tma_load_2d_2sm(stage, &tmap_a, bar, kb * 64, m0);                     // box {64 k, 128 rows}
          tc::tma_load_2d_2sm(stage + 128 * 128, &tmap_b, bar, kb * 64, n0);         // box {64 k, 128 rows}: half tile
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = tc::make_idesc_f16(256, kLinBN, 0, 0, 0, 0);
      int it = 0, nt = 0;
      for (int tile = cluster; tile < p.tiles; tile += nclusters, ++nt) {
        const int buf = nt & 1;
        tc::mbar_wait(tc::smem_u32(&acc_empty_bar[buf]), ((nt >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)buf * kLinBN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % kPairStages;
          tc::mbar_wait(tc::smem_u32(&full_bar[s]), (it / kPairStages) & 1);
          tc::tc_fence_after();
          const uint32_t stage = tiles + s * kPairStageBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = tc::make_smem_desc(stage + k * 32u, 16u, 1024u, tc::kLayoutSw128);
            const uint64_t bd = tc::make_smem_desc(stage + 128 * 128 + k * 32u, 16u, 1024u, tc::kLayoutSw128);
            tc::umma_f16_2sm(acc, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc::umma_commit_2sm(tc::smem_u32(&empty_bar[s]), (uint16_t)3);
        }
        tc::umma_commit_2sm(tc::smem_u32(&acc_full_bar[buf]), (uint16_t)3);
      }
    }
  } else {
    const float oscale = p.out_scale != nullptr ? __ldg(p.out_scale) : 1.f;
    int nt = 0;
    for (int tile = cluster; tile < p.tiles; tile += nclusters, ++nt) {
      const int buf = nt & 1;
      const int m0 = (tile % p.tiles_m) * 256 + (int)rank * 128, n0 = (tile / p.tiles_m) * kLinBN;
      tc::mbar_wait(tc::smem_u32(&acc_full_bar[buf]), (nt >> 1) & 1);
      tc::tc_fence_after();
      const int row = m0 + warp * 32 + lane;
      float* orow = row < p.M ? p.out + (long long)row * p.N + n0 : nullptr;
      const float* rrow = (p.resid != nullptr && row < p.M) ? p.resid + (long long)row * p.N + n0 : nullptr;
      const uint32_t acc = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * kLinBN;
#pragma unroll 1
      for (int c = 0; c < kLinBN; c += 32) {
        uint32_t v[32];
        tc::tmem_ld32(acc + c, v);
        tc::tmem_ld_wait();
        if (orow != nullptr) {
          float4 o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
          if (p.out_scale != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j].x *= oscale; o[j].y *= oscale; o[j].z *= oscale; o[j].w *= oscale; }
          }
          if (rrow != nullptr) {
            float4 rv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[j] = reinterpret_cast<const float4*>(rrow + c)[j];
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j].x += rv[j].x; o[j].y += rv[j].y; o[j].z += rv[j].z; o[j].w += rv[j].w; }
          }
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c) + j);
              o[j].x += bv.x; o[j].y += bv.y; o[j].z += bv.z; o[j].w += bv.w;
            }
          }
          float4* dst = reinterpret_cast<float4*>(orow + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = o[j];
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_leader(tc::smem_u32(&acc_empty_bar[buf]));
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();                       // no CTA exits while its peer can still signal its barriers / read its smem
  if (warp == 5) tc::tmem_dealloc2(tmem_base, 512);
}

int launch_linear_pair(const void* x16, const void* w16, const float* bias, const float* resid, const float* out_scale,
                       float* y, int M, int K, int N, cudaStream_t st) {
  CUtensorMap ma, mb;
  int rc = make_map_2d16(&ma, x16, false, M, K, 128);
  if (rc) return rc;
  rc = make_map_2d16(&mb, w16, false, N, K, 128);
  if (rc) return rc;
  constexpr size_t smem = (size_t)kPairStages * kPairStageBytes + 1024;
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    MLA_CUDA_TRY(cudaFuncSetAttribute(linear_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.store(1, std::memory_order_release);
  }
  LinearParams p{};
  p.out = y; p.bias = bias; p.resid = resid; p.out_scale = out_scale; p.M = M; p.N = N; p.KB = K / 64;
  p.tiles_m = (M + 255) / 256;
  p.tiles = p.tiles_m * (N / kLinBN);
  const int nclusters = std::max(1, std::min(p.tiles, mla::device_info().sm_count / 2));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * nclusters); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  MLA_CUDA_TRY(cudaLaunchKernelEx(&cfg, linear_gemm_pair_kernel, ma, mb, p));
  mla::count_launch();
  return 0;
}

// MLA_LINEAR_PERSISTENT: 1 (default) = persistent single-CTA kernel (no clusters: safe next to any other stream),
// 2 = persistent CTA pairs (same speed on B200: 62.3 vs 62.2 ms per m3ae step at B = 64), 0 = the one-tile-per-CTA kernels
// (67.6 ms).
int linear_persistent() {
  static const int v = [] {
    const char* e = getenv("MLA_LINEAR_PERSISTENT");
    return e != nullptr ? atoi(e) : 1;
  }();
  return v;
}

int launch_linear_persistent(const void* x16, const void* w16, const float* bias, const float* resid, const float* out_scale,
                             float* y, int M, int K, int N, cudaStream_t st) {
  CUtensorMap ma, mb;
  int rc = make_map_2d16(&ma, x16, false, M, K, 128);
  if (rc) return rc;
  rc = make_map_2d16(&mb, w16, false, N, K, kLinBN);
  if (rc) return rc;
  constexpr size_t smem = (size_t)kLinStages * kLinStageBytes + 1024;
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    MLA_CUDA_TRY(cudaFuncSetAttribute(linear_gemm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.store(1, std::memory_order_release);
  }
  LinearParams p{};
  p.out = y; p.bias = bias; p.resid = resid; p.out_scale = out_scale; p.M = M; p.N = N; p.KB = K / 64;
  p.tiles_m = (M + 127) / 128;
  p.tiles = p.tiles_m * (N / kLinBN);
  const int grid = std::min(p.tiles, mla::device_info().sm_count);
  linear_gemm_persistent_kernel<<<grid, kThreads, smem, st>>>(ma, mb, p);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}
}  // namespace

static int linear_gemm16(const void* x16, const void* w16, const float* bias, const float* resid, const float* out_scale,
                         float* y, int M, int K, int N, void* stream);

extern "C" int mla_linear_forward16(const void* x16, const void* w16, const float* bias, const float* resid, float* y, int M,
                                    int K, int N, void* stream) {
  return linear_gemm16(x16, w16, bias, resid, nullptr, y, M, K, N, stream);
}

// dx [M, K] = *out_scale * dy16 [M, N] (fp16, scaled by a power of two) * wt16 [K, N]^T (fp16: the TRANSPOSED weight,
// mla_filter_transpose16(w, wt16, N, 1, K, 0)): the data gradient of a Linear as the same K-major GEMM as its forward.
extern "C" int mla_linear_dgrad16(const void* dy16, const void* wt16, const float* out_scale, float* dx, int M, int K, int N,
                                  void* stream) {
  if (!out_scale) return MLA_E_BADARG;
  return linear_gemm16(dy16, wt16, nullptr, nullptr, out_scale, dx, M, N, K, stream);
}

static int linear_gemm16(const void* x16, const void* w16, const float* bias, const float* resid, const float* out_scale,
                         float* y, int M, int K, int N, void* stream) {
  if (!x16 || !w16 || !y || !mla::aligned16(x16) || !mla::aligned16(w16) || !mla::aligned16(y) || !mla::aligned16(bias) ||
      !mla::aligned16(resid))
    return MLA_E_BADARG;
  if (!conv_shape_ok(1, M, 1, K, N, 1, 1, 1, 0, 64)) return MLA_E_SHAPE;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  ConvGemmParams p{};
  p.OH = M; p.OW = 1; p.M = M; p.R = 1; p.S = 1; p.mul = 1;
  p.kcb = K / 64; p.KB = p.kcb; p.CinW = K;
  p.out = y; p.ldo = N; p.accumulate = 0; p.Cout = N; p.bias = bias; p.resid = resid; p.out_scale = out_scale;
  full_taps(p, 1, 1, false);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (linear_persistent() == 2 && N % kLinBN == 0 && M > 128 && !force_gather())
    return launch_linear_pair(x16, w16, bias, resid, out_scale, y, M, K, N, st);
  if (linear_persistent() == 1 && N % kLinBN == 0 && M >= 128)
    return launch_linear_persistent(x16, w16, bias, resid, out_scale, y, M, K, N, st);
  const int BN = (N % 128 == 0) ? 128 : 64;
  CUtensorMap map, gmap;
  int rc = make_map_im2col16(&gmap, x16, false, 1, M, 1, K, 0, 0, 0, 0, 1, 128);
  if (rc) return rc;
  // CTA pairs (one 256 x N MMA per k-step over two SMs): the Linear layers run on ONE stream, where pair kernels are
  // safe (see conv_pair_env); 256-column tiles where N allows. MLA_LINEAR_PAIR=0 falls back to single-CTA tiles.
  if (linear_pair() && N % 128 == 0 && M > 128) {
    const int BNp = (N % 256 == 0) ? 256 : 128;
    rc = make_map_2d16(&map, w16, false, N, K, BNp / 2);
    if (rc) return rc;
    dim3 gp((M + 127) / 128, N / BNp);
    return BNp == 256 ? launch<0, 256, 3, true, 1, 1, true, 1>(map, gmap, p, gp, st)
                      : launch<0, 128, 4, true, 1, 1, true, 1>(map, gmap, p, gp, st);
  }
  rc = make_map_2d16(&map, w16, false, N, K, BN);
  if (rc) return rc;
  dim3 grid((M + 127) / 128, N / BN);
  return BN == 64 ? launch<0, 64, 4, true, 1, 1, false, 1>(map, gmap, p, grid, st)
                  : launch<0, 128, 3, true, 1, 1, false, 1>(map, gmap, p, grid, st);
}

// dx [M, K] = dy [M, N] (fp32, TF32-rounded) * w [N, K] (fp32, TF32-rounded): the 1x1 case of mla_conv2d_dgrad, with CTA
// pairs (single-stream Linear layers). K % 64 == 0, N % 64 == 0.
extern "C" int mla_linear_dgrad(const float* dy, const float* w, float* dx, int M, int K, int N, void* stream) {
  if (!dy || !w || !dx || !mla::aligned16(dy) || !mla::aligned16(w) || !mla::aligned16(dx)) return MLA_E_BADARG;
  if (!conv_shape_ok(1, M, 1, K, N, 1, 1, 1, 0)) return MLA_E_SHAPE;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  ConvGemmParams p{};
  p.src = dy; p.Hs = M; p.Ws = 1; p.Cs = N; p.OH = M; p.OW = 1; p.M = M; p.R = 1; p.S = 1;
  p.mul = 1; p.sgn = -1; p.off = 0; p.div = 1; p.kcb = N / 32; p.KB = p.kcb; p.CinW = K;
  p.out = dx; p.ldo = K; p.accumulate = 0; p.Cout = N;
  full_taps(p, 1, 1, true);
  CUtensorMap map, gmap;
  int rc = make_map_2d(&map, w, N, K, 32, true);
  if (rc) return rc;
  rc = make_map_im2col(&gmap, dy, 1, M, 1, N, 0, 0, 0, 0, 1, 128, false);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int BN = (K % 128 == 0) ? 128 : 64;
  dim3 grid((M + 127) / 128, K / BN);
  if (linear_pair() && K % 128 == 0 && M > 128) {
    if (K % 256 == 0) {
      grid.y = K / 256;
      return launch<1, 256, 3, true, 1, 1, true>(map, gmap, p, grid, st);
    }
    return launch<1, 128, 4, true, 1, 1, true>(map, gmap, p, grid, st);
  }
  return BN == 64 ? launch<1, 64, 4, true>(map, gmap, p, grid, st) : launch<1, 128, 3, true>(map, gmap, p, grid, st);
}

static int dgrad16_impl(const void* dy16, const void* wt16, float* dx, int N, int H, int W, int Cin, int Cout, int R, int S,
                        int stride, int pad, int accumulate, void* stream, bool bf16, const float* out_scale);

extern "C" int mla_conv2d_dgrad16(const void* dy16, const void* wt16, float* dx, int N, int H, int W, int Cin, int Cout,
                                  int R, int S, int stride, int pad, int accumulate, void* stream) {
  return dgrad16_impl(dy16, wt16, dx, N, H, W, Cin, Cout, R, S, stride, pad, accumulate, stream, true, nullptr);
}

// fp16 operands: dy16 = fp16(dy * F) with F a power of two (mla_bn_backward_f16), wt16 = the transposed fp16 filter;
// dx (+)= *out_scale * (dy16 (*) wt16), out_scale = 1 / F (device scalar). TF32's 10-bit operand mantissa at the kind::f16 rate.
extern "C" int mla_conv2d_dgrad16_f16(const void* dy16, const void* wt16, const float* out_scale, float* dx, int N, int H, int W,
                                      int Cin, int Cout, int R, int S, int stride, int pad, int accumulate, void* stream) {
  if (!out_scale) return MLA_E_BADARG;
  return dgrad16_impl(dy16, wt16, dx, N, H, W, Cin, Cout, R, S, stride, pad, accumulate, stream, false, out_scale);
}

static int dgrad16_impl(const void* dy16, const void* wt16, float* dx, int N, int H, int W, int Cin, int Cout, int R, int S,
                        int stride, int pad, int accumulate, void* stream, bool bf16, const float* out_scale) {
  if (!dy16 || !wt16 || !dx || !mla::aligned16(dy16) || !mla::aligned16(wt16) || !mla::aligned16(dx)) return MLA_E_BADARG;
  if (!conv_shape_ok(N, H, W, Cin, Cout, R, S, stride, pad)) return MLA_E_SHAPE;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  const int OH = out_size(H, R, stride, pad), OW = out_size(W, S, stride, pad);
  if (OH <= 0 || OW <= 0 || (long long)N * H * W > 0x7fffffffLL) return MLA_E_SHAPE;
  if (!bf16 && !force_gather()) {
    const mla::StripPlan sp = mla::strip16_plan(N, H, W, Cin, Cout, R, S, stride, pad);
    if (sp.tiles > 0)
      return mla::conv_strip16_run(1, dy16, wt16, dx, N, H, W, accumulate ? 1 : 0, nullptr, out_scale, sp, stream);
  }
  ConvGemmParams p{};
  p.OH = H; p.OW = W; p.M = N * H * W; p.R = R; p.S = S; p.mul = 1;
  p.kcb = Cout / 64; p.CinW = Cout;      // GEMM K = dy channels; a filter tap spans Cout columns of the transposed filter
  p.out = dx; p.ldo = Cin; p.accumulate = accumulate ? 1 : 0; p.Cout = Cin; p.out_scale = out_scale;
  const int BN = (Cin % 128 == 0) ? 128 : 64;
  CUtensorMap map, gmap;
  int rc = make_map_2d16(&map, wt16, bf16, Cin, (long long)R * S * Cout, BN);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto go = [&](const ConvGemmParams& q, dim3 g) {
    if (const int bnp = bf16 ? 0 : pair16_bn(Cin, q.M, q.KB)) {
      const long long wc = (long long)R * S * Cout;
      return bnp == 256 ? launch_conv16_pair<256, 1>(wt16, false, Cin, wc, gmap, q, Cin / 256, st)
           : bnp == 128 ? launch_conv16_pair<128, 1>(wt16, false, Cin, wc, gmap, q, Cin / 128, st)
                        : launch_conv16_pair<64, 1>(wt16, false, Cin, wc, gmap, q, Cin / 64, st);
    }
    if (conv_persist() && q.KB > 0) {
      if (bf16)
        return BN == 64 ? launch_conv16_persistent<64, 4, 2>(map, gmap, q, (int)g.y, st)
                        : launch_conv16_persistent<128, 3, 2>(map, gmap, q, (int)g.y, st);
      return BN == 64 ? launch_conv16_persistent<64, 4, 1>(map, gmap, q, (int)g.y, st)
                      : launch_conv16_persistent<128, 3, 1>(map, gmap, q, (int)g.y, st);
    }
    if (bf16)
      return BN == 64 ? launch<0, 64, 4, true, 1, 1, false, 2>(map, gmap, q, g, st)
                      : launch<0, 128, 3, true, 1, 1, false, 2>(map, gmap, q, g, st);
    return BN == 64 ? launch<0, 64, 4, true, 1, 1, false, 1>(map, gmap, q, g, st)
                    : launch<0, 128, 3, true, 1, 1, false, 1>(map, gmap, q, g, st);
  };
  if (stride == 1) {
    p.KB = R * S * p.kcb;
    p.g_base_w = pad - (S - 1); p.g_base_h = pad - (R - 1);
    full_taps(p, R, S, true);
    rc = make_map_im2col16(&gmap, dy16, bf16, N, OH, OW, Cout, p.g_base_w, p.g_base_h, p.g_base_w + (W - OW),
                           p.g_base_h + (H - OH), 1, 128);
    if (rc) return rc;
    return go(p, dim3((p.M + 127) / 128, Cin / BN));
  }
  for (int ph = 0; ph < stride; ++ph) {      // one dense sub-convolution per output parity class (see mla_conv2d_dgrad)
    for (int pw = 0; pw < stride; ++pw) {
      ConvGemmParams q = p;
      const int Hs = (H - ph + stride - 1) / stride, Ws = (W - pw + stride - 1) / stride;
      if (Hs <= 0 || Ws <= 0) continue;
      q.nr = q.ns = 0;
      int qr[8], qs[8], lo_h = 1 << 20, lo_w = 1 << 20;
      for (int r = 0; r < R; ++r) {
        const int t = ph + pad - r;
        if (((t % stride) + stride) % stride != 0) continue;
        qr[q.nr] = (t >= 0 ? t : t - (stride - 1)) / stride;
        q.tap_r[q.nr] = (signed char)r;
        lo_h = min(lo_h, qr[q.nr]);
        ++q.nr;
      }
      for (int c = 0; c < S; ++c) {
        const int t = pw + pad - c;
        if (((t % stride) + stride) % stride != 0) continue;
        qs[q.ns] = (t >= 0 ? t : t - (stride - 1)) / stride;
        q.tap_s[q.ns] = (signed char)c;
        lo_w = min(lo_w, qs[q.ns]);
        ++q.ns;
      }
      q.OH = Hs; q.OW = Ws; q.M = N * Hs * Ws;
      q.o_mul = stride; q.o_ph = ph; q.o_pw = pw; q.o_H = H; q.o_W = W;
      q.KB = q.nr * q.ns * q.kcb;
      if (q.KB == 0) {
        if (accumulate) continue;
        lo_h = lo_w = 0;
      }
      for (int i = 0; i < q.nr; ++i) q.off_r[i] = (signed char)(qr[i] - lo_h);
      for (int j = 0; j < q.ns; ++j) q.off_s[j] = (signed char)(qs[j] - lo_w);
      q.g_base_w = lo_w; q.g_base_h = lo_h;
      rc = make_map_im2col16(&gmap, dy16, bf16, N, OH, OW, Cout, lo_w, lo_h, lo_w + Ws - OW, lo_h + Hs - OH, 1, 128);
      if (rc) return rc;
      rc = go(q, dim3((q.M + 127) / 128, Cin / BN));
      if (rc) return rc;
    }
  }
  return 0;
}

namespace {
struct WgradPlan {
  int BN, NT, splits, kb_per_split, KBtot, OH, OW;
  long long M;
  size_t ws_bytes;
};
// kp = pixels per k-block: 32 (tf32) or 64 (2-byte operands, which also need Cin % 64 == 0)
int wgrad_plan(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, WgradPlan* pl, int kp = 32) {
  if (!conv_shape_ok(N, H, W, Cin, Cout, R, S, stride, pad, kp)) return MLA_E_SHAPE;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  pl->OH = out_size(H, R, stride, pad);
  pl->OW = out_size(W, S, stride, pad);
  pl->M = (long long)N * pl->OH * pl->OW;
  if (pl->OH <= 0 || pl->OW <= 0 || pl->M > 0x7fffffffLL) return MLA_E_SHAPE;
  pl->BN = (Cin % 128 == 0) ? 128 : 64;
  pl->KBtot = (int)((pl->M + kp - 1) / kp);
  // 64-channel inputs with a 3-wide filter: one CTA takes a whole filter row (3 taps, N = 192)
  pl->NT = (Cin == 64 && S == 3 && !force_gather()) ? 3 : 1;
  const int tiles = (R * S / pl->NT) * ((Cin + pl->BN - 1) / pl->BN) * ((Cout + 127) / 128);
  static const int waves3 = [] { const char* e = getenv("MLA_WGRAD_WAVES"); return e ? atoi(e) : 2; }();
  // split-K target: every extra split writes (and the reduction re-reads) a full fp32 copy of dw, so layers whose tiles alone
  // nearly fill the GPU get ONE wave of CTAs; only layers with very few tiles split deeper (measured on the ResNet-18 shapes:
  // 1.22 -> 1.04 ms per visual backward, profiles/r2_wgrad_split_sweep.txt)
  static const int waves_env = [] { const char* e = getenv("MLA_WGRAD_TARGET"); return e ? std::max(1, atoi(e)) : 0; }();
  const int waves1 = waves_env ? waves_env : (tiles < 16 ? 2 : 1);
  const int target = (pl->NT == 3 ? waves3 : waves1) * di.sm_count;   // CTAs in total (2 are resident per SM)
  int splits = (target + tiles - 1) / tiles;
  const int min_kb = kp == 32 ? 8 : 4;                                    // >= 256 pixels per split
  splits = max(1, min(splits, pl->KBtot / min_kb > 0 ? pl->KBtot / min_kb : 1));
  pl->kb_per_split = (pl->KBtot + splits - 1) / splits;
  pl->splits = (pl->KBtot + pl->kb_per_split - 1) / pl->kb_per_split;
  pl->ws_bytes = pl->splits > 1 ? (size_t)pl->splits * Cout * R * S * Cin * sizeof(float) : 0;
  return 0;
}
}  // namespace

extern "C" size_t mla_conv2d_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S, int stride,
                                                   int pad) {
  WgradPlan pl;
  if (wgrad_plan(N, H, W, Cin, Cout, R, S, stride, pad, &pl) != 0) return 0;
  return pl.ws_bytes + 256;
}

extern "C" int mla_conv2d_wgrad(const float* x, const float* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                                int R, int S, int stride, int pad, void* ws, size_t ws_bytes, void* stream) {
  if (!x || !dy || !dw || !mla::aligned16(x) || !mla::aligned16(dy) || !mla::aligned16(dw) || !mla::aligned16(ws))
    return MLA_E_BADARG;
  WgradPlan pl;
  int rc = wgrad_plan(N, H, W, Cin, Cout, R, S, stride, pad, &pl);
  if (rc) return rc;
  if (pl.splits > 1 && (ws == nullptr || ws_bytes < pl.ws_bytes)) return MLA_E_WORKSPACE;
  ConvGemmParams p{};
  p.src = x; p.Hs = H; p.Ws = W; p.Cs = Cin; p.OH = pl.OH; p.OW = pl.OW; p.M = (int)pl.M; p.R = R; p.S = S;
  p.mul = stride; p.sgn = 1; p.off = -pad; p.div = 1; p.kcb = 0; p.KB = 0; p.CinW = Cin;
  p.ldo = (long long)R * S * Cin; p.accumulate = 0; p.Cout = Cout;
  p.kb_per_split = pl.kb_per_split; p.KBtot = pl.KBtot; p.splits = pl.splits;
  p.split_stride = (long long)Cout * R * S * Cin;
  p.out = pl.splits > 1 ? static_cast<float*>(ws) : dw;
  CUtensorMap map;
  rc = make_map_2d(&map, dy, pl.M, Cout, 32, true);
  if (rc) return rc;
  dim3 grid((Cin + pl.BN - 1) / pl.BN, (Cout + 127) / 128, (R * S / pl.NT) * pl.splits);   // last ci tile may be partial
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (force_gather()) {
    if (Cin % pl.BN != 0) return MLA_E_SHAPE;
    rc = pl.BN == 64 ? launch<2, 64, 4, false>(map, map, p, grid, st) : launch<2, 128, 3, false>(map, map, p, grid, st);
  } else {
    p.g_base_w = p.g_base_h = -pad;
    full_taps(p, R, S, false);
    CUtensorMap gmap;
    rc = make_map_im2col(&gmap, x, N, H, W, Cin, -pad, -pad, pad - (S - 1), pad - (R - 1), stride, 32, true);
    if (rc) return rc;
    rc = pl.NT == 3 ? launch<2, 64, 2, true, 3>(map, gmap, p, grid, st)
         : pl.BN == 64 ? launch<2, 64, 4, true>(map, gmap, p, grid, st) : launch<2, 128, 3, true>(map, gmap, p, grid, st);
  }
  if (rc) return rc;
  if (pl.splits > 1) {
    const long long n4 = p.split_stride / 4;
    splitk_reduce_kernel<<<(unsigned)(((n4 + 1) / 2 + 127) / 128), 128, 0, st>>>(static_cast<const float*>(ws), dw, n4, pl.splits, n4);
    MLA_CUDA_TRY(cudaGetLastError());
    mla::count_launch();
  }
  return 0;
}

// wgrad with 2-byte operands: x16 [N,H,W,Cin] bf16, dy16 [N,OH,OW,Cout] bf16 (both MN-major: K = pixels) -> dw fp32
// [Cout,R,S,Cin]. 64-pixel k-blocks, plain SWIZZLE_128B panels; same split-K / multi-tap structure as the TF32 wgrad.
extern "C" size_t mla_conv2d_wgrad16_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S, int stride,
                                                     int pad) {
  WgradPlan pl;
  if (wgrad_plan(N, H, W, Cin, Cout, R, S, stride, pad, &pl, 64) != 0) return 0;
  return pl.ws_bytes + 256;
}

static int wgrad16_impl(const void* x16, const void* dy16, float* dw, int N, int H, int W, int Cin, int Cout, int R, int S,
                        int stride, int pad, void* ws, size_t ws_bytes, void* stream, bool bf16, const float* out_scale);

extern "C" int mla_conv2d_wgrad16(const void* x16, const void* dy16, float* dw, int N, int H, int W, int Cin, int Cout,
                                  int R, int S, int stride, int pad, void* ws, size_t ws_bytes, void* stream) {
  return wgrad16_impl(x16, dy16, dw, N, H, W, Cin, Cout, R, S, stride, pad, ws, ws_bytes, stream, true, nullptr);
}

// fp16 operands: x16 = the fp16 activation the forward convolution already read, dy16 = fp16(dy * F) (mla_bn_backward_f16);
// dw = *out_scale * (dy16^T (*) x16), out_scale = 1 / F (device scalar).
extern "C" int mla_conv2d_wgrad16_f16(const void* x16, const void* dy16, const float* out_scale, float* dw, int N, int H, int W,
                                      int Cin, int Cout, int R, int S, int stride, int pad, void* ws, size_t ws_bytes,
                                      void* stream) {
  if (!out_scale) return MLA_E_BADARG;
  return wgrad16_impl(x16, dy16, dw, N, H, W, Cin, Cout, R, S, stride, pad, ws, ws_bytes, stream, false, out_scale);
}

// dw [N, K] = *out_scale * dy16 [M, N]^T (fp16, scaled by a power of two) * x16 [M, K] (fp16): the weight gradient of a
// Linear; workspace from mla_conv2d_wgrad16_workspace_bytes(1, M, 1, K, N, 1, 1, 1, 0).
extern "C" int mla_linear_wgrad16(const void* x16, const void* dy16, const float* out_scale, float* dw, int M, int K, int N,
                                  void* ws, size_t ws_bytes, void* stream) {
  if (!out_scale) return MLA_E_BADARG;
  return wgrad16_impl(x16, dy16, dw, 1, M, 1, K, N, 1, 1, 1, 0, ws, ws_bytes, stream, false, out_scale);
}

static int wgrad16_impl(const void* x16, const void* dy16, float* dw, int N, int H, int W, int Cin, int Cout, int R, int S,
                        int stride, int pad, void* ws, size_t ws_bytes, void* stream, bool bf16, const float* out_scale) {
  if (!x16 || !dy16 || !dw || !mla::aligned16(x16) || !mla::aligned16(dy16) || !mla::aligned16(dw) || !mla::aligned16(ws))
    return MLA_E_BADARG;
  WgradPlan pl;
  int rc = wgrad_plan(N, H, W, Cin, Cout, R, S, stride, pad, &pl, 64);
  if (rc) return rc;
  if (pl.splits > 1 && (ws == nullptr || ws_bytes < pl.ws_bytes)) return MLA_E_WORKSPACE;
  ConvGemmParams p{};
  p.OH = pl.OH; p.OW = pl.OW; p.M = (int)pl.M; p.R = R; p.S = S; p.mul = stride; p.CinW = Cin;
  p.ldo = (long long)R * S * Cin; p.accumulate = 0; p.Cout = Cout;
  p.kb_per_split = pl.kb_per_split; p.KBtot = pl.KBtot; p.splits = pl.splits;
  p.split_stride = (long long)Cout * R * S * Cin;
  p.out = pl.splits > 1 ? static_cast<float*>(ws) : dw;
  p.g_base_w = p.g_base_h = -pad;
  full_taps(p, R, S, false);
  CUtensorMap map, gmap;
  p.out_scale = out_scale;
  rc = make_map_2d16(&map, dy16, bf16, pl.M, Cout, 64);                       // box {64 co, 64 pixel rows}
  if (rc) return rc;
  rc = make_map_im2col16(&gmap, x16, bf16, N, H, W, Cin, -pad, -pad, pad - (S - 1), pad - (R - 1), stride, 64);
  if (rc) return rc;
  dim3 grid(Cin / pl.BN, (Cout + 127) / 128, (R * S / pl.NT) * pl.splits);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bf16)
    rc = pl.NT == 3 ? launch<2, 64, 2, true, 3, 1, false, 2>(map, gmap, p, grid, st)
         : pl.BN == 64 ? launch<2, 64, 4, true, 1, 1, false, 2>(map, gmap, p, grid, st)
                       : launch<2, 128, 3, true, 1, 1, false, 2>(map, gmap, p, grid, st);
  else
    rc = pl.NT == 3 ? launch<2, 64, 2, true, 3, 1, false, 1>(map, gmap, p, grid, st)
         : pl.BN == 64 ? launch<2, 64, 4, true, 1, 1, false, 1>(map, gmap, p, grid, st)
                       : launch<2, 128, 3, true, 1, 1, false, 1>(map, gmap, p, grid, st);
  if (rc) return rc;
  if (pl.splits > 1) {
    const long long n4 = p.split_stride / 4;
    splitk_reduce_kernel<<<(unsigned)(((n4 + 1) / 2 + 127) / 128), 128, 0, st>>>(static_cast<const float*>(ws), dw, n4, pl.splits, n4);
    MLA_CUDA_TRY(cudaGetLastError());
    mla::count_launch();
  }
  return 0;
}
