// tcgen05 / TMEM forward of the fused attention core of the transformer encoders (reference models/m3ae.py:103-121,
// models/cav_mae.py:93-101 through timm's Attention): softmax(scale * Q K^T, key-padding positions FILLED with -1e7) V for
// head width 64, reading q / k / v in place from the fp16 qkv tensor [B, S, 3, H, 64] through ONE 3-D tensor map and writing
// the head-concatenated fp32 [B, S, H * 64] layout plus the (row max, row sum) pair the backward pass needs. Same contract as
// attn_fwd_kernel<64> (attention.cu), which it replaces for S <= 527.
//
// One CTA per (128-query tile, head, batch row), 192 threads:
//   warp 4    TMA producer: the Q tile and EVERY key / value row of the (batch, head) land in shared memory as 128-byte rows
//             (64 fp16 = one head), 128B-swizzled, rows past S zero-filled, on one mbarrier.
//   warp 5    MMA issuer. S = Q K^T: Q [128 x 64] K-major x K [NK x 64] K-major -> fp32 scores for all NK keys at once in
//             TMEM columns [0, NK) (no online rescaling: S <= 527 keys fit the 512 columns). Then, chunk by chunk as the
//             probabilities arrive, O = P V: P chunk [128 x 64 keys] K-major x V chunk [64 keys x 64] MN-major (rows = keys is
//             exactly how v sits in qkv) -> O in TMEM columns [0, 64), which chunk 0 of S has vacated by then.
//   warps 0-3 softmax, one query row per thread = one TMEM lane: pass 1 row maximum (tcgen05.ld 16 columns at a time), pass 2
//             p = exp(s - max), row sum, fp16 P written into 128B-swizzled shared memory (over the dead K / Q tiles) and
//             published per 64-key chunk through an mbarrier; epilogue O / sum -> global.
// Keys are put on the tensor cores in multiples of 16 (UMMA N granularity): NK = 16 * floor(S / 16). The S - NK < 16 REMAINING
// keys (the class token of the 257-token sequences: 256 + 1) are handled by the softmax threads on the CUDA cores from the
// same shared-memory rows — 128 FMAs per row and key — instead of padding the sequence to 272 columns: S = 257 then needs 256
// TMEM columns and 104 KB of shared memory per CTA, so TWO CTAs are resident per SM and one's softmax overlaps the other's MMAs.
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kDh = 64;
constexpr int kThreads = 320;        // 8 softmax warps, TMA producer, MMA issuer
constexpr int kMaxChunks = 8;          // 64-key chunks on the tensor cores: NK <= 512
constexpr int kMaxS = 512 + 15;
constexpr float kLog2e = 1.4426950408889634f;

struct AttnTcParams {
  const float* mask;   // [B, S] (> 0: padded key) or nullptr
  float* out;          // [B, S, H*64]
  float* stats;        // [B, H, S, 2]
  int B, S, H;
  float scale;
  int nk;              // keys on the tensor cores (multiple of 16)
  int nkb;             // 64-row boxes of K / V in shared memory = ceil(S / 64)
  uint32_t a_bytes;    // region A: [K boxes | Q tile], later overwritten by the P chunks
  uint32_t tmem_cols;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// 8 fp16 of row `row`, 16-byte chunk `c`, of a 128B-swizzled tile of 128-byte rows starting (1024-aligned) at `tile`
__device__ __forceinline__ uint4 lds_chunk(uint32_t tile, int row, int c) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(tile + (uint32_t)row * 128u + tc::swz16(c, row)));
  return v;
}
__device__ __forceinline__ void h8_to_f(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __half22float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}

// ---- softmax building blocks: one 16-column block of this thread's score row
template <bool MASK>
__device__ __forceinline__ void block_max(const uint32_t (&v)[16], uint32_t kbits, float (&m)[4], bool& any_masked) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float t = __uint_as_float(v[i]);
    if (MASK) t = ((kbits >> i) & 1u) ? -INFINITY : t;
    m[i & 3] = fmaxf(m[i & 3], t);
  }
  if (MASK) any_masked = any_masked || (kbits & 0xffffu) != 0u;
}
// p = 2^(s * sc2 - mx) (padded keys: 2^(fill2 - mx)), partial row sums, fp16 P into the swizzled row (2 x 16 bytes)
template <bool MASK>
__device__ __forceinline__ void block_exp(const uint32_t (&v)[16], uint32_t kbits, float sc2, float mx, float fillp,
                                          float (&l)[4], uint32_t prow, int chunk0, int r) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = ex2(fmaf(__uint_as_float(v[8 * c + i]), sc2, -mx));
      if (MASK) t = ((kbits >> (8 * c + i)) & 1u) ? fillp : t;
      e[i] = t;
      l[i & 3] += t;
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + tc::swz16(chunk0 + c, r)), "r"(pack_h2(e[0], e[1])),
                 "r"(pack_h2(e[2], e[3])), "r"(pack_h2(e[4], e[5])), "r"(pack_h2(e[6], e[7]))
                 : "memory");
  }
}
__device__ __forceinline__ uint32_t key_bits16(const uint32_t* kbits, int c0) { return kbits[c0 >> 5] >> (c0 & 31); }

// REMCAP = how many remaining keys the CUDA-core path is compiled for: 1 (S = 256 + 1, 512, ...: one live register) or 15
template <bool MASK, int REMCAP>
__global__ void __launch_bounds__(kThreads, 2) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap, AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t load_bar, s_full, o_full, p_full[kMaxChunks];
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t s_kbits[kMaxS / 32 + 2];   // bit = padded key
  __shared__ float s_x[2][128];                  // row maxima, then row sums, of the two column halves
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int S = p.S, nk = p.nk, rem = S - nk;
  const int npc = (nk + 63) >> 6;                        // P chunks
  const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t k_base = base;                          // nkb boxes of 64 key rows
  const uint32_t q_base = base + (uint32_t)p.nkb * 8192u;   // 128 query rows
  const uint32_t p_base = base;                          // P chunk j at + j * 16 KB (K / Q are dead by then)
  const uint32_t v_base = base + p.a_bytes;              // nkb boxes of 64 value rows

  if (threadIdx.x == 0) {
    tc::mbar_init(tc::smem_u32(&load_bar), 1);
    tc::mbar_init(tc::smem_u32(&s_full), 1);
    tc::mbar_init(tc::smem_u32(&o_full), 1);
    for (int j = 0; j < kMaxChunks; ++j) tc::mbar_init(tc::smem_u32(&p_full[j]), 128);
    tc::fence_mbar_init();
  }
  if (warp == 8 && lane == 0) tc::tma_prefetch_desc(&tmap);
  if (warp == 9) {
    tc::tmem_alloc(tc::smem_u32(&tmem_slot), p.tmem_cols);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      const uint32_t bar = tc::smem_u32(&load_bar);
      tc::mbar_arrive_expect_tx(bar, (uint32_t)(2 + 2 * p.nkb) * 8192u);
      tma_load_3d(q_base, &tmap, bar, h * kDh, q0, b);
      tma_load_3d(q_base + 8192u, &tmap, bar, h * kDh, q0 + 64, b);
      for (int i = 0; i < p.nkb; ++i) tma_load_3d(k_base + (uint32_t)i * 8192u, &tmap, bar, (p.H + h) * kDh, i * 64, b);
      for (int i = 0; i < p.nkb; ++i) tma_load_3d(v_base + (uint32_t)i * 8192u, &tmap, bar, (2 * p.H + h) * kDh, i * 64, b);
    }
  } else if (warp == 9) {
    if (lane == 0 && nk > 0) {
      tc::mbar_wait(tc::smem_u32(&load_bar), 0);
      tc::tc_fence_after();
      for (int n0 = 0; n0 < nk; n0 += 256) {
        const int n = min(256, nk - n0);
        const uint32_t idesc = tc::make_idesc_f16(128, n, 0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = tc::make_smem_desc(q_base + k * 32u, 16u, 1024u, tc::kLayoutSw128);
          const uint64_t bd = tc::make_smem_desc(k_base + (uint32_t)n0 * 128u + k * 32u, 16u, 1024u, tc::kLayoutSw128);
          tc::umma_f16(tmem + (uint32_t)n0, ad, bd, idesc, k != 0 ? 1u : 0u);
        }
      }
      tc::umma_commit(tc::smem_u32(&s_full));
      constexpr uint32_t idesc_pv = tc::make_idesc_f16(128, kDh, 0, 0, 0, 1);     // A = P K-major, B = V MN-major
      for (int j = 0; j < npc; ++j) {
        tc::mbar_wait(tc::smem_u32(&p_full[j]), 0);
        tc::tc_fence_after();
        const int ksteps = min(4, (nk - j * 64) >> 4);
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t ad = tc::make_smem_desc(p_base + (uint32_t)j * 16384u + k * 32u, 16u, 1024u, tc::kLayoutSw128);
          const uint64_t bd = tc::make_smem_desc(v_base + (uint32_t)j * 8192u + k * 2048u, 8192u, 1024u, tc::kLayoutSw128);
          tc::umma_f16(tmem, ad, bd, idesc_pv, (j | k) != 0 ? 1u : 0u);
        }
      }
      tc::umma_commit(tc::smem_u32(&o_full));
    }
  } else {
    // ===================== softmax / epilogue =====================
    // thread = (query row r = TMEM lane, column half): warps 0-3 take the first ceil(npc / 2) 64-key chunks, warps 4-7 the rest
    const int wq = warp & 3, half = warp >> 2;
    const int r = wq * 32 + lane, q = q0 + r;
    const bool wvalid = q0 + wq * 32 < S;                // warp-uniform: this warp owns at least one real query row
    const int jmid = (npc + 1) >> 1;
    const int j_lo = half == 0 ? 0 : jmid, j_hi = half == 0 ? jmid : npc;
    const int c_lo = j_lo * 64, c_hi = min(nk, j_hi * 64);
    if (MASK) {
      for (int w = threadIdx.x; w < (S + 31) / 32; w += 256) {
        uint32_t bits = 0;
        for (int i = 0; i < 32; ++i) {
          const int key = w * 32 + i;
          if (key < S && p.mask[(long long)b * S + key] > 0.f) bits |= 1u << i;
        }
        s_kbits[w] = bits;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    tc::mbar_wait(tc::smem_u32(&load_bar), 0);
    const float sc2 = p.scale * kLog2e;                  // log2 domain: t = s * scale * log2(e)
    const float fill2 = -1e7f * kLog2e;
    const uint32_t trow = tmem + ((uint32_t)(wq * 32) << 16);
    float sx[REMCAP];                                    // the remaining keys' scores (both halves compute them)
    float mx = -INFINITY;
    if (wvalid) {
      bool any_masked = false;
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (rem > 0) {                                     // q . k on the CUDA cores, operands from shared memory
#pragma unroll
        for (int e = 0; e < REMCAP; ++e) sx[e] = 0.f;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          float qf[8];
          h8_to_f(lds_chunk(q_base, r, c), qf);
#pragma unroll
          for (int e = 0; e < REMCAP; ++e) {
            if (e < rem) {
              float kf[8];
              h8_to_f(lds_chunk(k_base, nk + e, c), kf);
#pragma unroll
              for (int d = 0; d < 8; ++d) sx[e] = fmaf(qf[d], kf[d], sx[e]);
            }
          }
        }
#pragma unroll
        for (int e = 0; e < REMCAP; ++e) {
          if (e < rem) {
            const bool pad = MASK && ((s_kbits[(nk + e) >> 5] >> ((nk + e) & 31)) & 1u);
            if (pad) any_masked = true;
            else m4[e & 3] = fmaxf(m4[e & 3], sx[e]);
          }
        }
      }
      if (c_hi > c_lo) {
        tc::mbar_wait(tc::smem_u32(&s_full), 0);
        tc::tc_fence_after();
        // software pipeline over 16-column blocks: the load of block i + 1 is in flight while block i is reduced
        uint32_t va[16], vb[16];
        tmem_ld16(trow + (uint32_t)c_lo, va);
#pragma unroll 1
        for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
          tc::tmem_ld_wait();
          if (c0 + 16 < c_hi) tmem_ld16(trow + (uint32_t)(c0 + 16), vb);
          block_max<MASK>(va, MASK ? key_bits16(s_kbits, c0) : 0u, m4, any_masked);
          if (c0 + 16 < c_hi) {
            tc::tmem_ld_wait();
            if (c0 + 32 < c_hi) tmem_ld16(trow + (uint32_t)(c0 + 32), va);
            block_max<MASK>(vb, MASK ? key_bits16(s_kbits, c0 + 16) : 0u, m4, any_masked);
          }
        }
      }
      mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc2;      // scale > 0: the maximum commutes with the scaling
      if (any_masked) mx = fmaxf(mx, fill2);
      s_x[half][r] = mx;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");       // maxima exchanged; every read of the Q / K tiles is done
    float l = 0.f;
    if (wvalid) {
      mx = fmaxf(s_x[0][r], s_x[1][r]);
      const float fillp = ex2(fill2 - mx);               // probability of a padded key (1 when every key is padded)
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
      if (c_hi > c_lo) {
        uint32_t va[16], vb[16];
        tmem_ld16(trow + (uint32_t)c_lo, va);
#pragma unroll 1
        for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
          const int j = c0 >> 6, ch = (c0 & 63) >> 3;    // P chunk, first 16-byte chunk of the row inside it
          const uint32_t prow = p_base + (uint32_t)j * 16384u + (uint32_t)r * 128u;
          tc::tmem_ld_wait();
          if (c0 + 16 < c_hi) tmem_ld16(trow + (uint32_t)(c0 + 16), vb);
          block_exp<MASK>(va, MASK ? key_bits16(s_kbits, c0) : 0u, sc2, mx, fillp, l4, prow, ch, r);
          if (c0 + 16 < c_hi) {
            tc::tmem_ld_wait();
            if (c0 + 32 < c_hi) tmem_ld16(trow + (uint32_t)(c0 + 32), va);
            block_exp<MASK>(vb, MASK ? key_bits16(s_kbits, c0 + 16) : 0u, sc2, mx, fillp, l4, prow, ch + 2, r);
          }
          if (((c0 + 32) & 63) == 0 || c0 + 32 >= c_hi) {   // the 64-key chunk is complete
            tc::tc_fence_before();
            tc::fence_proxy_async();                     // P chunk j (generic-proxy stores) -> visible to the MMA
            tc::mbar_arrive(tc::smem_u32(&p_full[j]));
          }
        }
      }
      l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      if (rem > 0) {
#pragma unroll
        for (int e = 0; e < REMCAP; ++e) {
          if (e < rem) {
            const bool pad = MASK && ((s_kbits[(nk + e) >> 5] >> ((nk + e) & 31)) & 1u);
            sx[e] = pad ? fillp : ex2(fmaf(sx[e], sc2, -mx));
            if (half == 0) l += sx[e];
          }
        }
      }
      s_x[half][r] = l;
    } else {
      for (int j = j_lo; j < j_hi; ++j) {                // rows past the sequence: garbage P rows, never stored
        tc::fence_proxy_async();
        tc::mbar_arrive(tc::smem_u32(&p_full[j]));
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");       // row sums exchanged
    if (nk > 0) {
      tc::mbar_wait(tc::smem_u32(&o_full), 0);
      tc::tc_fence_after();
    }
    if (wvalid) {
      l = s_x[0][r] + s_x[1][r];
      const float inv = 1.f / l;
      const int d0 = half * 32;                          // this thread's 32 output columns
      float o[32];
      if (nk > 0) {
        uint32_t v[32];
        tc::tmem_ld32(trow + (uint32_t)d0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(v[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0.f;
      }
      if (rem > 0) {
#pragma unroll
        for (int e = 0; e < REMCAP; ++e) {
          if (e < rem) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float vf[8];
              h8_to_f(lds_chunk(v_base, nk + e, (d0 >> 3) + c), vf);
#pragma unroll
              for (int i = 0; i < 8; ++i) o[8 * c + i] = fmaf(sx[e], vf[i], o[8 * c + i]);
            }
          }
        }
      }
      if (q < S) {
        float* orow = p.out + ((long long)b * S + q) * ((long long)p.H * kDh) + h * kDh + d0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<float4*>(orow)[i] = make_float4(o[4 * i] * inv, o[4 * i + 1] * inv, o[4 * i + 2] * inv,
                                                           o[4 * i + 3] * inv);
        // (row max, row sum) in the natural-log domain of the scaled, filled scores — what the backward kernels expect;
        // every key padded: exactly the fill value
        if (half == 0)
          *reinterpret_cast<float2*>(p.stats + 2 * (((long long)b * p.H + h) * S + q)) =
              make_float2(mx == fill2 ? -1e7f : mx / kLog2e, l);
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 9) tc::tmem_dealloc(tmem, p.tmem_cols);
}

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

}  // namespace

namespace mla {

// MLA_ATTN_TC=0 puts the forward back on the mma.sync kernel (A/B comparison)
bool attn_tc_applicable(int S, int Dh) {
  static const bool off = [] { const char* e = getenv("MLA_ATTN_TC"); return e != nullptr && e[0] == '0'; }();
  return !off && Dh == kDh && S >= 1 && S <= kMaxS;
}

int attn_fwd_tc(const void* qkv16, const float* mask, float* out, float* stats, int B, int S, int H, float scale,
                void* stream) {
  const DeviceInfo& di = device_info();
  if (di.ok != 1) return di.ok;
  EncodeTiledFn fn = encode_fn();
  if (!fn) return MLA_E_NODEVICE;
  CUtensorMap tmap;
  cuuint64_t dims[3] = {(cuuint64_t)3 * H * kDh, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)3 * H * kDh * 2, (cuuint64_t)S * 3 * H * kDh * 2};
  cuuint32_t box[3] = {(cuuint32_t)kDh, 64u, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  if (fn(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(qkv16), dims, strides, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return MLA_E_BADARG;
  AttnTcParams p{};
  p.mask = mask; p.out = out; p.stats = stats; p.B = B; p.S = S; p.H = H; p.scale = scale;
  p.nk = (S / 16) * 16;
  p.nkb = (S + 63) / 64;
  const int npc = (p.nk + 63) / 64;
  p.a_bytes = (uint32_t)std::max((p.nkb * 64 + 128) * 128, npc * 16384);
  int cols = 64;
  while (cols < p.nk) cols <<= 1;
  p.tmem_cols = (uint32_t)cols;
  const size_t smem = 1024 + (size_t)p.a_bytes + (size_t)p.nkb * 8192;
  if (smem > (size_t)di.smem_optin - 8 * 1024) return MLA_E_SHAPE;
  static std::atomic<size_t> configured{0};
  if (smem > configured.load(std::memory_order_acquire)) {
    const size_t want = (size_t)di.smem_optin - 8 * 1024;
    MLA_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tc_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
    MLA_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tc_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
    MLA_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tc_kernel<false, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
    MLA_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tc_kernel<true, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
    configured.store(want, std::memory_order_release);
  }
  dim3 grid((S + 127) / 128, H, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool small_rem = S - p.nk <= 1;
  if (mask != nullptr) {
    if (small_rem) attn_fwd_tc_kernel<true, 1><<<grid, kThreads, smem, st>>>(tmap, p);
    else attn_fwd_tc_kernel<true, 15><<<grid, kThreads, smem, st>>>(tmap, p);
  } else {
    if (small_rem) attn_fwd_tc_kernel<false, 1><<<grid, kThreads, smem, st>>>(tmap, p);
    else attn_fwd_tc_kernel<false, 15><<<grid, kThreads, smem, st>>>(tmap, p);
  }
  MLA_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace mla
