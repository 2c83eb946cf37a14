// Memory-bound pieces of the ResNet-18 encoders around the tcgen05 convolutions (NHWC fp32):
//   stem im2col (7x7/2 on the raw NCHW input -> [M][Kp] rows for a tensor-core GEMM)
//   BatchNorm2d training statistics / running-stat update        backbone.py:40,44,150 (nn.BatchNorm2d)
//   BN apply (+ReLU) (+residual with its own BN affine)           backbone.py:41-50
//   BN+ReLU+MaxPool 3x3/2 fused for the stem                      backbone.py:150-152
//   their backward passes, global average pool fwd/bwd            basic_model.py:61-65
// All reductions are two-stage with a fixed order (deterministic); per-channel sums are carried
// in double across threads/blocks so mean/variance match torch's to fp32 rounding.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int kRedThreads = 256;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Round-to-nearest fp32 -> TF32 (10-bit mantissa), kept in an fp32 container. Every tensor that
// feeds a tcgen05 kind::tf32 MMA is rounded by its PRODUCER, so the tensor core's own operand
// truncation (round-toward-zero, a systematic -2^-11 relative bias that BatchNorm hides in training
// but that compounds layer by layer in eval mode) becomes a no-op.
__device__ __forceinline__ float tf32r(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float4 tf32r4(float4 v) { return make_float4(tf32r(v.x), tf32r(v.y), tf32r(v.z), tf32r(v.w)); }

__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    st4(dst + 4 * i, tf32r4(ld4(src + 4 * i)));
}

// ReLU sign bitmask: bit e of the mask belongs to element e of the [M][C] tensor, i.e. float4 number i owns the nibble
// (mask[i / 8] >> 4 * (i % 8)) & 15 (bit q = component q is > 0). Written by the forward BN-apply, read by the two
// BN-backward passes instead of the 32x larger activation itself.
__device__ __forceinline__ unsigned relu_nibble(const unsigned int* __restrict__ mask, long long i) {
  return (__ldg(mask + (i >> 3)) >> (4 * (int)(i & 7))) & 15u;
}

// ------------------------------------------------------------------------------ 2-byte copies
__device__ __forceinline__ uint2 pack_h4(float4 v) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<const unsigned int*>(&a), *reinterpret_cast<const unsigned int*>(&b));
}
// fp16 pack with saturation to +-65504 (an overflow of the power-of-two-scaled gradients becomes the largest finite value,
// never inf)
__device__ __forceinline__ uint2 pack_h4_sat(float4 v) {
  uint2 r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r.x) : "f"(v.y), "f"(v.x));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r.y) : "f"(v.w), "f"(v.z));
  return r;
}
__device__ __forceinline__ uint2 pack_b4(float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<const unsigned int*>(&a), *reinterpret_cast<const unsigned int*>(&b));
}

__global__ void cast16_kernel(const float* __restrict__ src, uint2* __restrict__ dst, long long n4, int bf16) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ld4(src + 4 * i);
    dst[i] = bf16 ? pack_b4(v) : pack_h4(v);
  }
}

// wt[ci][rs][co] = fp16 or bf16 (w[co][rs][ci]); one block per (rs, 32x32 tile), transposed through shared memory
__global__ void filter_transpose16_kernel(const float* __restrict__ w, unsigned short* __restrict__ wt, int Cout, int RS,
                                          int Cin, int bf16) {
  __shared__ float tile[32][33];
  const int rs = blockIdx.z, co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int co = co0 + j, ci = ci0 + tx;
    tile[j][tx] = (co < Cout && ci < Cin) ? w[((long long)co * RS + rs) * Cin + ci] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int ci = ci0 + j, co = co0 + tx;
    if (ci < Cin && co < Cout) {
      const float v = tile[tx][j];
      unsigned short bits;
      if (bf16) { const __nv_bfloat16 h = __float2bfloat16_rn(v); bits = *reinterpret_cast<const unsigned short*>(&h); }
      else { const __half h = __float2half_rn(v); bits = *reinterpret_cast<const unsigned short*>(&h); }
      wt[((long long)ci * RS + rs) * Cout + co] = bits;
    }
  }
}

// All filters of an encoder in ONE launch: seg[i] = {element offset of filter i in the flat parameter buffer (= in the 2-byte
// buffer), Cout, RS, Cin, first tile}; a block finds its filter by a linear scan over the (<= 64) segments.
struct TransposeSeg { long long off; int Cout, RS, Cin, tile0; };
__global__ void filter_transpose16_batch_kernel(const float* __restrict__ flat, unsigned short* __restrict__ flat_t,
                                                const TransposeSeg* __restrict__ seg, int nseg, int bf16) {
  __shared__ float tile[32][33];
  int i = 0;
  while (i + 1 < nseg && (int)blockIdx.x >= seg[i + 1].tile0) ++i;
  const TransposeSeg sg = seg[i];
  const int t = blockIdx.x - sg.tile0;
  const int tci = (sg.Cin + 31) / 32, tco = (sg.Cout + 31) / 32;
  const int rs = t / (tci * tco), rem = t - rs * (tci * tco);
  const int co0 = (rem / tci) * 32, ci0 = (rem % tci) * 32;
  const float* w = flat + sg.off;
  unsigned short* wt = flat_t + sg.off;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int co = co0 + j, ci = ci0 + tx;
    tile[j][tx] = (co < sg.Cout && ci < sg.Cin) ? w[((long long)co * sg.RS + rs) * sg.Cin + ci] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int ci = ci0 + j, co = co0 + tx;
    if (ci < sg.Cin && co < sg.Cout) {
      const float v = tile[tx][j];
      unsigned short bits;
      if (bf16) { const __nv_bfloat16 h = __float2bfloat16_rn(v); bits = *reinterpret_cast<const unsigned short*>(&h); }
      else { const __half h = __float2half_rn(v); bits = *reinterpret_cast<const unsigned short*>(&h); }
      wt[((long long)ci * sg.RS + rs) * sg.Cout + co] = bits;
    }
  }
}

// ------------------------------------------------------------------------------ stem im2col
// col[m][k], m = (n, oh, ow), k = (r*S + s)*Cin + ci for k < R*S*Cin, zero for the pad columns.
// Input element (n, ci, h, w) lives at in[(n / T)*sB + (n % T)*sT + ci*sC + h*W + w] (raw NCHW /
// NCTHW tensors: the frame fold of backbone.py:144-147 is pure index arithmetic here).
// One block per output row (n, oh): the R input rows of every channel are staged in shared memory
// (coalesced, zero-padded left/right/top/bottom), a k -> offset table replaces the per-element
// div/mod, and the row's OW x Kp outputs leave as coalesced float4 stores. HBM-bound by the col write.
// OUT16: the matrix leaves as an fp16 copy (fprop16 operand) and, when col16b != NULL, a bf16 copy (wgrad16 operand)
// instead of TF32-rounded fp32.
template <bool OUT16>
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ in, float* __restrict__ col,
                                                          uint2* __restrict__ col16, uint2* __restrict__ col16b, int T,
                                                          long long sB, long long sT, long long sC, int Cin, int H, int W,
                                                          int OH, int OW, int R, int S, int stride, int pad, int K, int Kp) {
  extern __shared__ __align__(16) float sm[];
  const int Wp = W + 2 * pad + 2;                  // padded row pitch (+2 keeps the last window in range)
  int* koff = reinterpret_cast<int*>(sm + (((size_t)Cin * R * Wp + 3) & ~(size_t)3));
  const int n = blockIdx.x / OH, oh = blockIdx.x - n * OH;
  const float* src = in + (long long)(n / T) * sB + (long long)(n % T) * sT;
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
    int o = -1;
    if (k < K) {
      const int ci = k % Cin, rs = k / Cin;
      o = (ci * R + rs / S) * Wp + rs % S;
    }
    koff[k] = o;
  }
  const int rows = Cin * R;
  for (int i = threadIdx.x; i < rows * Wp; i += blockDim.x) {
    const int row = i / Wp, wp = i - row * Wp;
    const int ci = row / R, r = row - ci * R;
    const int h = oh * stride + r - pad, w = wp - pad;
    float v = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W) v = src[(long long)ci * sC + (long long)h * W + w];
    sm[i] = OUT16 ? v : tf32r(v);
  }
  __syncthreads();
  const int K4 = Kp >> 2;
  float* dst = OUT16 ? nullptr : col + (long long)blockIdx.x * OW * Kp;
  const long long base4 = (long long)blockIdx.x * OW * K4;
  for (int i = threadIdx.x; i < OW * K4; i += blockDim.x) {
    const int ow = i / K4, k = (i - ow * K4) * 4;
    const int base = ow * stride;
    const int4 o = *reinterpret_cast<const int4*>(koff + k);
    float4 v;
    v.x = o.x >= 0 ? sm[o.x + base] : 0.f;
    v.y = o.y >= 0 ? sm[o.y + base] : 0.f;
    v.z = o.z >= 0 ? sm[o.z + base] : 0.f;
    v.w = o.w >= 0 ? sm[o.w + base] : 0.f;
    if (OUT16) {
      col16[base4 + i] = pack_h4(v);
      if (col16b != nullptr) col16b[base4 + i] = pack_b4(v);
    } else {
      st4(dst + 4 * (long long)i, v);
    }
  }
}

// dst[row][0..kp) = src[row][0..k) zero padded (kp >= k), or the inverse copy when unpad.
__global__ void pad_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int k, int kp, int unpad) {
  const int total = rows * (unpad ? k : kp);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (unpad) {
      const int r = i / k, c = i - r * k;
      dst[i] = src[(long long)r * kp + c];
    } else {
      const int r = i / kp, c = i - r * kp;
      dst[i] = c < k ? src[(long long)r * k + c] : 0.f;
    }
  }
}

// ------------------------------------------------------------------- per-channel reductions
// One launch per BatchNorm statistic pass: grid (row blocks, 64-channel chunks), block 16 float4 lanes x 16
// rows. Every block writes its partial sums (double); the LAST block to finish a channel chunk (atomic
// ticket) adds the partials in block order — fixed order, so the result is deterministic — and runs the
// finalisation (statistics / running stats, or dgamma / dbeta) for its 64 channels. No second launch.
// MODE 0: a = x, b = x*x (BN statistics).
// MODE 1: g = dz * (z > 0 or no mask); a = g, b = g * (y - mean) * invstd (BN backward sums).
// MODE 3: MODE 1 for the stem, walking the POOLED pixels (M = their number): g = the pooled gradient where the ReLU is alive,
//         x = the fp16 convolution output at the window's argmax (pool_window); the |g| bound is multiplied by 4 (a pre-pool
//         pixel can be the argmax of four windows).
// MODE 2: x = per-tile partial sums [rows = tiles][2][C] written by the fprop epilogue: a = x[r][0][c],
//         b = x[r][1][c]; finalised like MODE 0 (f.M = the real number of pixels).
__device__ __forceinline__ float4 ldh4(const __half* p) {     // four fp16 -> float4 (8-byte aligned)
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// The stem's fused ReLU + MaxPool(3, 2, 1) in the backward pass (C = 64). idx = argmax window position 0..8 of a pooled
// element, with bit 4 set when the pooled value was not positive (dead ReLU: matches no position).
struct PoolGather {
  const float* dp;              // [N][PH][PW][C] gradient of the pooled activation
  const unsigned char* idx;     // [N][PH][PW][C]
  const __half* y16;            // [N][H][W][C] convolution output (MODE 3 of the reduction reads it at the argmax)
  int H, W, PH, PW;             // pre-pool / pooled extent
};
// Gradient at pre-pool pixel (n, h, w), channels [c, c + 4): the sum over the (at most four) pooling windows that contain the
// pixel of dp where the window's recorded argmax is this pixel.
__device__ __forceinline__ float4 pool_gather(const PoolGather& pg, int n, int h, int w, int c, int C) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int oh0 = h >> 1, oh1 = (h + 1) >> 1, ow0 = w >> 1, ow1 = (w + 1) >> 1;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int oh = a ? oh1 : oh0;
    if ((a && oh1 == oh0) || oh >= pg.PH) continue;
    const int r = h - (2 * oh - 1);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int ow = b ? ow1 : ow0;
      if ((b && ow1 == ow0) || ow >= pg.PW) continue;
      const int pos = r * 3 + (w - (2 * ow - 1));
      const long long o = (((long long)n * pg.PH + oh) * pg.PW + ow) * C + c;
      const uchar4 id = *reinterpret_cast<const uchar4*>(pg.idx + o);
      const float4 d = ld4(pg.dp + o);
      if (id.x == pos) acc.x += d.x;
      if (id.y == pos) acc.y += d.y;
      if (id.z == pos) acc.z += d.z;
      if (id.w == pos) acc.w += d.w;
    }
  }
  return acc;
}
// MODE 3 of the reduction walks the POOLED pixels (4x fewer than the pre-pool ones): sum g = sum over windows of dp, sum g * xhat =
// sum over windows of dp * xhat(argmax). Returns dp masked by the dead flag in `g` and the convolution output at each channel's
// argmax in `x`.
__device__ __forceinline__ void pool_window(const PoolGather& pg, unsigned prow, int c, int C, float4* x, float4* g) {
  const unsigned pw = prow % (unsigned)pg.PW, t = prow / (unsigned)pg.PW;
  const unsigned ph = t % (unsigned)pg.PH, n = t / (unsigned)pg.PH;
  const long long o = (long long)prow * C + c;
  const uchar4 id = *reinterpret_cast<const uchar4*>(pg.idx + o);
  const float4 d = ld4(pg.dp + o);
  const unsigned char ids[4] = {id.x, id.y, id.z, id.w};
  const float ds[4] = {d.x, d.y, d.z, d.w};
  float xs[4], gs[4];
  const __half* ybase = pg.y16 + (((long long)n * pg.H + (2 * (int)ph - 1)) * pg.W + (2 * (int)pw - 1)) * C + c;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int pos = ids[q];
    const bool alive = pos < 9;
    const int r = (pos * 11) >> 5, s2 = pos - 3 * r;          // pos / 3, pos % 3 for 0..8
    xs[q] = alive ? __half2float(ybase[((long long)r * pg.W + s2) * C + q]) : 0.f;
    gs[q] = alive ? ds[q] : 0.f;
  }
  *x = make_float4(xs[0], xs[1], xs[2], xs[3]);
  *g = make_float4(gs[0], gs[1], gs[2], gs[3]);
}

struct BnFinal {
  long long M;
  // MODE 0
  const float* gamma; const float* beta;
  float* running_mean; float* running_var;
  float momentum, eps;
  float* mean_out; float* invstd_out; float* scale_out; float* shift_out;
  // MODE 1
  float* dgamma; float* dbeta; float* sums;   // sums [2][C]: dbeta, dgamma (read by the apply pass)
  // MODE 1, scaled-fp16 gradients: bound_bits[chunk] accumulates (atomicMax on the bit pattern of non-negative floats:
  // order-independent, hence deterministic) max over the chunk of |g| * |gamma| * invstd; the last block of the chunk
  // publishes it to cbound[chunk] and clears the accumulator. NULL = off.
  unsigned int* bound_bits; float* cbound;
  PoolGather pool;     // MODE 3
};

template <int MODE>
__global__ void __launch_bounds__(kRedThreads, (MODE == 1 || MODE == 3) ? 3 : 4) channel_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                                     const float* __restrict__ z,
                                                                     const float* __restrict__ mean,
                                                                     const float* __restrict__ invstd, long long M, int C,
                                                                     long long rows_per_block, double* __restrict__ part,
                                                                     unsigned int* __restrict__ counters, BnFinal f,
                                                                     const unsigned int* __restrict__ mask) {
  __shared__ double s_acc[16][2][64];
  __shared__ double s_fin[2][64];
  __shared__ int s_last;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int chunk = blockIdx.y, nrb = gridDim.x;
  const int c0 = chunk * 64 + 4 * tx;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
  {
    float4 mu = make_float4(0, 0, 0, 0), is = make_float4(1, 1, 1, 1);
    if (MODE == 1 || MODE == 3) { mu = ld4(mean + c0); is = ld4(invstd + c0); }
    // this thread's rows: r0 + ty + 16 * i, i < n. Four rows per step with all their loads issued first (memory-level
    // parallelism); fp32 accumulation over at most 32 rows, then flushed to double.
    const long long n = (r1 - r0 - ty + 15) / 16;
    const long long rs = (MODE == 2) ? 2LL * C : (long long)C;   // row stride in floats
    const float* xp = x + (r0 + ty) * rs + c0;
    const float* dzp = (MODE == 1) ? dz + (r0 + ty) * rs + c0 : nullptr;
    const float* zp = (MODE == 1 && z != nullptr) ? z + (r0 + ty) * rs + c0 : nullptr;
    float fa[4] = {0, 0, 0, 0}, fb[4] = {0, 0, 0, 0};
    float gmx[4] = {0, 0, 0, 0};                                     // MODE 1: max |g| per channel of this thread
    auto flush = [&]() {
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] += (double)fa[q]; b[q] += (double)fb[q]; fa[q] = 0.f; fb[q] = 0.f; }
    };
    auto row = [&](const float4& v, const float4& w, unsigned nb) {   // nb: bit q = component q passes the ReLU mask
      if (MODE == 2) {          // w = the second half of the partial row (sum of squares)
        fa[0] += v.x; fa[1] += v.y; fa[2] += v.z; fa[3] += v.w;
        fb[0] += w.x; fb[1] += w.y; fb[2] += w.z; fb[3] += w.w;
      } else if (MODE == 0) {
        fa[0] += v.x; fa[1] += v.y; fa[2] += v.z; fa[3] += v.w;
        fb[0] = fmaf(v.x, v.x, fb[0]); fb[1] = fmaf(v.y, v.y, fb[1]);
        fb[2] = fmaf(v.z, v.z, fb[2]); fb[3] = fmaf(v.w, v.w, fb[3]);
      } else {                  // w = dz
        const float gx = (nb & 1u) ? w.x : 0.f, gy = (nb & 2u) ? w.y : 0.f;
        const float gz = (nb & 4u) ? w.z : 0.f, gw = (nb & 8u) ? w.w : 0.f;
        fa[0] += gx; fa[1] += gy; fa[2] += gz; fa[3] += gw;
        fb[0] = fmaf(gx, (v.x - mu.x) * is.x, fb[0]); fb[1] = fmaf(gy, (v.y - mu.y) * is.y, fb[1]);
        fb[2] = fmaf(gz, (v.z - mu.z) * is.z, fb[2]); fb[3] = fmaf(gw, (v.w - mu.w) * is.w, fb[3]);
        gmx[0] = fmaxf(gmx[0], fabsf(gx)); gmx[1] = fmaxf(gmx[1], fabsf(gy));
        gmx[2] = fmaxf(gmx[2], fabsf(gz)); gmx[3] = fmaxf(gmx[3], fabsf(gw));
      }
    };
    auto bits = [&](long long o, long long rowidx) -> unsigned {   // ReLU pass bits of the float4 at row offset o
      if (MODE != 1) return 15u;
      if (mask != nullptr) return relu_nibble(mask, (rowidx * C + c0) >> 2);
      if (zp != nullptr) {
        const float4 zz = ld4(zp + o);
        return (zz.x > 0.f ? 1u : 0u) | (zz.y > 0.f ? 2u : 0u) | (zz.z > 0.f ? 4u : 0u) | (zz.w > 0.f ? 8u : 0u);
      }
      return 15u;
    };
    const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
    const int kFlush = (MODE == 2) ? 1 : 8;     // steps of 4 rows between flushes
    long long i = 0;
    int since = 0;
    for (; i + 4 <= n; i += 4) {
      float4 v[4], w[4];
      unsigned nb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long o = (i + u) * 16 * rs;
        if (MODE == 3) {
          pool_window(f.pool, (unsigned)(r0 + ty + (i + u) * 16), c0, C, &v[u], &w[u]);
        } else {
          v[u] = ld4(xp + o);
          w[u] = (MODE == 2) ? ld4(xp + o + C) : (MODE == 1 ? ld4(dzp + o) : one);
        }
        nb[u] = bits(o, r0 + ty + (i + u) * 16);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) row(v[u], w[u], nb[u]);
      if (++since == kFlush) { flush(); since = 0; }
    }
    for (; i < n; ++i) {
      const long long o = i * 16 * rs;
      float4 v, w;
      if (MODE == 3) {
        pool_window(f.pool, (unsigned)(r0 + ty + i * 16), c0, C, &v, &w);
      } else {
        v = ld4(xp + o);
        w = (MODE == 2) ? ld4(xp + o + C) : (MODE == 1 ? ld4(dzp + o) : one);
      }
      row(v, w, bits(o, r0 + ty + i * 16));
    }
    flush();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      s_acc[ty][0][4 * tx + q] = a[q];
      s_acc[ty][1][4 * tx + q] = b[q];
    }
    if ((MODE == 1 || MODE == 3) && f.bound_bits != nullptr) {
      const float4 ga = ld4(f.gamma + c0);
      float m = fmaxf(fmaxf(gmx[0] * fabsf(ga.x) * is.x, gmx[1] * fabsf(ga.y) * is.y),
                      fmaxf(gmx[2] * fabsf(ga.z) * is.z, gmx[3] * fabsf(ga.w) * is.w));
      m = mla::warp_max(m);
      if (MODE == 3) m *= 4.f;
      if (!(m == m)) m = __int_as_float(0x7f800000);                   // NaN gradients: the bound becomes +inf
      if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(&f.bound_bits[chunk], __float_as_uint(m));
    }
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    double sum = 0;
#pragma unroll
    for (int y = 0; y < 16; ++y) sum += s_acc[y][which][c];   // fixed order
    part[((size_t)chunk * nrb + blockIdx.x) * 128 + threadIdx.x] = sum;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&counters[chunk], 1u) == (unsigned)(nrb - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    // column (threadIdx.x & 127) of the partial rows; the two halves of the block take alternate partial rows and are
    // combined in a fixed order (deterministic)
    __shared__ double s_half[128];
    const int col = threadIdx.x & 127, half = threadIdx.x >> 7;
    const double* pp = part + (size_t)chunk * nrb * 128 + col;
    double sum = 0;
    // this serial tail sits on the critical path of every BatchNorm launch (up to 592 partial rows on the 64-channel
    // layers): 12 independent L2 loads in flight per thread, added in row order
    int blk = half;
    for (; blk + 22 < nrb; blk += 24) {
      double v[12];
#pragma unroll
      for (int u = 0; u < 12; ++u) v[u] = __ldcg(pp + (size_t)(blk + 2 * u) * 128);
#pragma unroll
      for (int u = 0; u < 12; ++u) sum += v[u];
    }
    for (; blk < nrb; blk += 2) sum += __ldcg(pp + (size_t)blk * 128);
    if (half == 1) s_half[col] = sum;
    __syncthreads();
    if (half == 0) s_fin[col >> 6][col & 63] = sum + s_half[col];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    counters[chunk] = 0;   // leave the workspace reusable
    if ((MODE == 1 || MODE == 3) && f.bound_bits != nullptr) {
      f.cbound[chunk] = __uint_as_float(atomicExch(&f.bound_bits[chunk], 0u));
    }
  }
  if (threadIdx.x >= 64) return;
  const int c = chunk * 64 + threadIdx.x;
  const double sa = s_fin[0][threadIdx.x], sb = s_fin[1][threadIdx.x];
  if (MODE == 0 || MODE == 2) {
    // mean, biased var -> invstd, scale = gamma*invstd, shift = beta - mean*scale; running stats: momentum
    // update with the UNBIASED variance (torch BatchNorm2d semantics)
    const double mu = sa / (double)f.M;
    double var = sb / (double)f.M - mu * mu;
    if (var < 0) var = 0;
    const float is = (float)(1.0 / sqrt(var + (double)f.eps));
    const float sc = f.gamma[c] * is;
    f.mean_out[c] = (float)mu;
    f.invstd_out[c] = is;
    f.scale_out[c] = sc;
    f.shift_out[c] = f.beta[c] - (float)mu * sc;
    if (f.running_mean != nullptr) {
      const double unbiased = f.M > 1 ? var * (double)f.M / (double)(f.M - 1) : var;
      f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * (float)mu;
      f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)unbiased;
    }
  } else {
    if (f.dbeta) f.dbeta[c] = (float)sa;
    if (f.dgamma) f.dgamma[c] = (float)sb;
    f.sums[c] = (float)sa;
    f.sums[C + c] = (float)sb;
  }
}

__global__ void bn_eval_coeffs_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps, int C,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] / sqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - rm[c] * sc;
}

// out = relu?( y*scale + shift + (res ? res*rscale + rshift : 0) ). Four float4 per thread and step, loads first.
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, const float* __restrict__ res,
                                                       const float* __restrict__ rscale, const float* __restrict__ rshift,
                                                       int relu, float* __restrict__ out, long long n4, int C4,
                                                       unsigned int* __restrict__ mask_out, uint2* __restrict__ out16,
                                                       uint2* __restrict__ out16b) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool pow2 = (C4 & (C4 - 1)) == 0;
  // the loop bound is warp-uniform (i0 - lane), so all 32 lanes stay together for the mask shuffles
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 - (threadIdx.x & 31) < n4; i0 += 4 * stride) {
    float4 v[4], r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
        v[u] = ld4(y + 4 * i);
        if (res != nullptr) r[u] = ld4(res + 4 * i);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i - (threadIdx.x & 31) >= n4) break;     // warp-uniform
      const bool live = i < n4;                    // tail lanes of the last warp only take part in the shuffles
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        const int c = (pow2 ? (int)(i & (C4 - 1)) : (int)(i % C4)) * 4;     // (a 64-bit modulo costs ~100 instructions)
        const float4 sc = ld4(scale + c), sh = ld4(shift + c);
        o = make_float4(fmaf(v[u].x, sc.x, sh.x), fmaf(v[u].y, sc.y, sh.y), fmaf(v[u].z, sc.z, sh.z),
                        fmaf(v[u].w, sc.w, sh.w));
        if (res != nullptr) {
          float4 rr = r[u];
          if (rscale != nullptr) {
            const float4 rs = ld4(rscale + c), rb = ld4(rshift + c);
            rr = make_float4(fmaf(rr.x, rs.x, rb.x), fmaf(rr.y, rs.y, rb.y), fmaf(rr.z, rs.z, rb.z), fmaf(rr.w, rs.w, rb.w));
          }
          o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        // fp32 copy: residual add / pooling, and the operand of the TF32 convolutions — rounded to TF32 only when it is
        // one (no 2-byte copy requested): on the 2-byte path the residual stream stays in full fp32, as under cuDNN
        if (out != nullptr) st4(out + 4 * i, (out16 == nullptr && out16b == nullptr) ? tf32r4(o) : o);
        if (out16 != nullptr) out16[i] = pack_h4(o);       // fp16 copy: the operand of the kind::f16 forward convolutions
        if (out16b != nullptr) out16b[i] = pack_b4(o);     // bf16 copy: the x operand of the kind::f16 wgrad
      }
      if (mask_out != nullptr) {       // 8 lanes = 8 float4 = one 32-bit mask word
        unsigned w = ((o.x > 0.f ? 1u : 0u) | (o.y > 0.f ? 2u : 0u) | (o.z > 0.f ? 4u : 0u) | (o.w > 0.f ? 8u : 0u))
                     << (4 * (threadIdx.x & 7));
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        w |= __shfl_xor_sync(0xffffffffu, w, 4);
        if (live && (threadIdx.x & 7) == 0) mask_out[i >> 3] = w;
      }
    }
  }
}

// dy = gamma*invstd * (g - dbeta/M - xhat*dgamma/M), g = dz*(z>0); optionally also writes g.
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ z, const float* __restrict__ y,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ sums, float inv_m,
                                    float* __restrict__ dy, float* __restrict__ g_out, long long n4, int C4,
                                    const unsigned int* __restrict__ mask, uint2* __restrict__ dy16,
                                    const float* __restrict__ cbound, int chunks, float* __restrict__ gscale) {
  const int C = C4 * 4;
  // scaled-fp16 gradients: F = the power of two that puts the bound of |dy| (from the reduction pass) into [2^8, 2^9):
  // 2^7 of headroom below fp16's largest finite value (the bound ignores the mean terms; conversions saturate), fp16's
  // full 10-bit mantissa down to 2^-22 of the bound. Every thread derives the same F; block 0 publishes F and 1 / F.
  float F = 1.f;
  if (cbound != nullptr) {
    float bound = 0.f;
    for (int k = 0; k < chunks; ++k) bound = fmaxf(bound, cbound[k]);
    if (bound > 0.f && bound < __int_as_float(0x7f800000)) {
      int e;
      frexpf(bound, &e);                        // bound = m * 2^e, m in [0.5, 1)
      e = max(-100, min(100, 9 - e));
      F = ldexpf(1.f, e);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { gscale[0] = F; gscale[1] = 1.f / F; }
  }
  const bool pow2 = (C4 & (C4 - 1)) == 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (pow2 ? (int)(i & (C4 - 1)) : (int)(i % C4)) * 4;
    float4 g = ld4(dz + 4 * i);
    if (z != nullptr) {
      const float4 zz = ld4(z + 4 * i);
      g.x = zz.x > 0.f ? g.x : 0.f; g.y = zz.y > 0.f ? g.y : 0.f;
      g.z = zz.z > 0.f ? g.z : 0.f; g.w = zz.w > 0.f ? g.w : 0.f;
    } else if (mask != nullptr) {
      const unsigned nb = relu_nibble(mask, i);
      g.x = (nb & 1u) ? g.x : 0.f; g.y = (nb & 2u) ? g.y : 0.f;
      g.z = (nb & 4u) ? g.z : 0.f; g.w = (nb & 8u) ? g.w : 0.f;
    }
    if (g_out != nullptr) st4(g_out + 4 * i, g);
    const float4 v = ld4(y + 4 * i), mu = ld4(mean + c), is = ld4(invstd + c), ga = ld4(gamma + c);
    const float4 db = ld4(sums + c), dg = ld4(sums + C + c);
    float4 o;
    o.x = ga.x * is.x * (g.x - db.x * inv_m - (v.x - mu.x) * is.x * dg.x * inv_m);
    o.y = ga.y * is.y * (g.y - db.y * inv_m - (v.y - mu.y) * is.y * dg.y * inv_m);
    o.z = ga.z * is.z * (g.z - db.z * inv_m - (v.z - mu.z) * is.z * dg.z * inv_m);
    o.w = ga.w * is.w * (g.w - db.w * inv_m - (v.w - mu.w) * is.w * dg.w * inv_m);
    if (dy != nullptr) st4(dy + 4 * i, tf32r4(o));    // dy only feeds dgrad / wgrad
    if (dy16 != nullptr) {
      if (cbound != nullptr) dy16[i] = pack_h4_sat(make_float4(o.x * F, o.y * F, o.z * F, o.w * F));   // fp16, scaled
      else dy16[i] = pack_b4(o);                  // bf16 copy: the operand of the bf16 kind::f16 dgrad
    }
  }
}

// ------------------------------------------------------------------ stem BN+ReLU+MaxPool 3x3/2 p1
__global__ void bn_relu_maxpool_kernel(const float* __restrict__ y, const float* __restrict__ scale,
                                       const float* __restrict__ shift, float* __restrict__ out,
                                       unsigned char* __restrict__ idx, int N, int H, int W, int C4, int OH, int OW,
                                       uint2* __restrict__ out16, uint2* __restrict__ out16b) {
  const long long total = (long long)N * OH * OW * C4;
  const int C = C4 * 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    long long t = i / C4;
    const int ow = (int)(t % OW); t /= OW;
    const int oh = (int)(t % OH);
    const int n = (int)(t / OH);
    const float4 sc = ld4(scale + c), sh = ld4(shift + c);
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int bi[4] = {0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = 2 * oh - 1 + r;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = 2 * ow - 1 + s;
        if (w < 0 || w >= W) continue;
        const float4 v = ld4(y + (((long long)n * H + h) * W + w) * C + c);
        const float a[4] = {fmaxf(fmaf(v.x, sc.x, sh.x), 0.f), fmaxf(fmaf(v.y, sc.y, sh.y), 0.f),
                            fmaxf(fmaf(v.z, sc.z, sh.z), 0.f), fmaxf(fmaf(v.w, sc.w, sh.w), 0.f)};
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (a[q] > best[q]) { best[q] = a[q]; bi[q] = r * 3 + s; }
      }
    }
    st4(out + 4 * i, tf32r4(make_float4(best[0], best[1], best[2], best[3])));
    if (out16 != nullptr) out16[i] = pack_h4(make_float4(best[0], best[1], best[2], best[3]));
    if (out16b != nullptr) out16b[i] = pack_b4(make_float4(best[0], best[1], best[2], best[3]));
    reinterpret_cast<uchar4*>(idx)[i] = make_uchar4((unsigned char)bi[0], (unsigned char)bi[1], (unsigned char)bi[2],
                                                    (unsigned char)bi[3]);
  }
}

// g[n,h,w,c] = sum over pooling windows whose argmax is (h,w) of dp * (p > 0)   (ReLU mask folded in)
__global__ void maxpool_relu_bwd_kernel(const float* __restrict__ dp, const float* __restrict__ p,
                                        const unsigned char* __restrict__ idx, float* __restrict__ g, int N, int H, int W,
                                        int C4, int OH, int OW) {
  const long long total = (long long)N * H * W * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long long t = i / C4;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int oh0 = h >> 1, oh1 = (h + 1) >> 1, ow0 = w >> 1, ow1 = (w + 1) >> 1;
    for (int a = 0; a < 2; ++a) {
      const int oh = a ? oh1 : oh0;
      if ((a && oh1 == oh0) || oh >= OH) continue;
      const int r = h - (2 * oh - 1);
      for (int b = 0; b < 2; ++b) {
        const int ow = b ? ow1 : ow0;
        if ((b && ow1 == ow0) || ow >= OW) continue;
        const int s = w - (2 * ow - 1);
        const long long o = (((long long)n * OH + oh) * OW + ow) * C4 + c4;
        const uchar4 id = reinterpret_cast<const uchar4*>(idx)[o];
        const float4 d = ld4(dp + 4 * o), pv = ld4(p + 4 * o);
        const int pos = r * 3 + s;
        if (id.x == pos && pv.x > 0.f) acc[0] += d.x;
        if (id.y == pos && pv.y > 0.f) acc[1] += d.y;
        if (id.z == pos && pv.z > 0.f) acc[2] += d.z;
        if (id.w == pos && pv.w > 0.f) acc[3] += d.w;
      }
    }
    st4(g + 4 * i, make_float4(acc[0], acc[1], acc[2], acc[3]));
  }
}

__device__ __forceinline__ void unpack_h8(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __half22float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}

// The same fused BN + ReLU + MaxPool over the stem convolution's fp16 output (stem_s2d.cu), C = 64: one block per pooled row
// (n, ph), one thread per (pooled pixel, 8 channels) — no integer division, 16-byte loads. idx bit 4 marks a pooled value
// that is not positive (dead ReLU), so the backward gather needs neither the pooled activation nor a separate mask.
__global__ void __launch_bounds__(256) bn_relu_maxpool16_kernel(const __half* __restrict__ y, const float* __restrict__ scale,
                                                                const float* __restrict__ shift, float* __restrict__ out,
                                                                uint4* __restrict__ out16, unsigned char* __restrict__ idx,
                                                                int H, int W, int OH, int OW) {
  constexpr int C = 64;
  const int n = blockIdx.x / OH, oh = blockIdx.x - n * OH;
  for (int j = threadIdx.x; j < OW * 8; j += blockDim.x) {
    const int ow = j >> 3, c = (j & 7) * 8;
    float sc[8], sh[8];
    *reinterpret_cast<float4*>(sc) = ld4(scale + c); *reinterpret_cast<float4*>(sc + 4) = ld4(scale + c + 4);
    *reinterpret_cast<float4*>(sh) = ld4(shift + c); *reinterpret_cast<float4*>(sh + 4) = ld4(shift + c + 4);
    uint4 v[9];
    bool ok[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = 2 * oh - 1 + r;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = 2 * ow - 1 + s;
        ok[r * 3 + s] = h >= 0 && h < H && w >= 0 && w < W;
        if (ok[r * 3 + s]) v[r * 3 + s] = *reinterpret_cast<const uint4*>(y + (((long long)n * H + h) * W + w) * C + c);
      }
    }
    float best[8];
    int bi[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { best[q] = -INFINITY; bi[q] = 0; }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      if (!ok[t]) continue;
      float x[8];
      unpack_h8(v[t], x);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float a = fmaxf(fmaf(x[q], sc[q], sh[q]), 0.f);
        if (a > best[q]) { best[q] = a; bi[q] = t; }
      }
    }
    const long long o = (((long long)n * OH + oh) * OW + ow) * C + c;
    if (out != nullptr) {
      st4(out + o, make_float4(best[0], best[1], best[2], best[3]));
      st4(out + o + 4, make_float4(best[4], best[5], best[6], best[7]));
    }
    if (out16 != nullptr) {
      const uint2 lo = pack_h4(make_float4(best[0], best[1], best[2], best[3]));
      const uint2 hi = pack_h4(make_float4(best[4], best[5], best[6], best[7]));
      out16[o >> 3] = make_uint4(lo.x, lo.y, hi.x, hi.y);
    }
    unsigned b[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) b[q] = (unsigned)bi[q] | (best[q] > 0.f ? 0u : 16u);
    *reinterpret_cast<uint2*>(idx + o) = make_uint2(b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24),
                                                    b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24));
  }
}

// BatchNorm backward of the stem with the MaxPool + ReLU backward folded in: g is gathered from the pooled gradient, xhat
// comes from the fp16 convolution output, dy leaves as fp16 * F (see bn_bwd_apply_kernel). The dense fp32 gradient of the
// pre-pool activation (411 MB per visual batch: written once, read twice) never exists. One block per pre-pool row (n, h), one
// thread per (pixel, 8 channels); C = 64.
__global__ void __launch_bounds__(256) pool_bn_bwd_apply_kernel(PoolGather pg, const float* __restrict__ mean,
                                                                const float* __restrict__ invstd,
                                                                const float* __restrict__ gamma, const float* __restrict__ sums,
                                                                float inv_m, uint4* __restrict__ dy16,
                                                                const float* __restrict__ cbound, int chunks,
                                                                float* __restrict__ gscale) {
  constexpr int C = 64;
  float F = 1.f;
  {
    float bound = 0.f;
    for (int k = 0; k < chunks; ++k) bound = fmaxf(bound, cbound[k]);
    if (bound > 0.f && bound < __int_as_float(0x7f800000)) {
      int e;
      frexpf(bound, &e);
      e = max(-100, min(100, 9 - e));
      F = ldexpf(1.f, e);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { gscale[0] = F; gscale[1] = 1.f / F; }
  }
  const int n = blockIdx.x / pg.H, h = blockIdx.x - n * pg.H;
  for (int j = threadIdx.x; j < pg.W * 8; j += blockDim.x) {
    const int w = j >> 3, c = (j & 7) * 8;
    const long long o = (((long long)n * pg.H + h) * pg.W + w) * C + c;
    const uint4 yv = *reinterpret_cast<const uint4*>(pg.y16 + o);
    const float4 g0 = pool_gather(pg, n, h, w, c, C), g1 = pool_gather(pg, n, h, w, c + 4, C);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float v[8], r[8];
    unpack_h8(yv, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float mu = __ldg(mean + c + q), is = __ldg(invstd + c + q), ga = __ldg(gamma + c + q);
      const float db = __ldg(sums + c + q), dg = __ldg(sums + C + c + q);
      r[q] = ga * is * (g[q] - db * inv_m - (v[q] - mu) * is * dg * inv_m) * F;
    }
    const uint2 lo = pack_h4_sat(make_float4(r[0], r[1], r[2], r[3])), hi = pack_h4_sat(make_float4(r[4], r[5], r[6], r[7]));
    dy16[o >> 3] = make_uint4(lo.x, lo.y, hi.x, hi.y);
  }
}

// --------------------------------------------------------------------------- global average pool
__global__ void avgpool_fwd_kernel(const float* __restrict__ fm, float* __restrict__ feat, int rows, int C) {
  // grid.x = samples; feat[b][c] = mean over `rows` consecutive NHWC rows
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += fm[((long long)b * rows + r) * C + c];
    feat[(long long)b * C + c] = s / (float)rows;
  }
}
__global__ void avgpool_bwd_kernel(const float* __restrict__ dfeat, float* __restrict__ dfm, int B, int rows, int C4) {
  const long long total = (long long)B * rows * C4;
  const float inv = 1.f / (float)rows;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const long long b = i / ((long long)rows * C4);
    const float4 d = ld4(dfeat + (b * C4 + c4) * 4);
    st4(dfm + 4 * i, make_float4(d.x * inv, d.y * inv, d.z * inv, d.w * inv));
  }
}

int ew_grid(long long n, int threads) {
  const mla::DeviceInfo& di = mla::device_info();
  long long g = (n + threads - 1) / threads;
  const long long cap = (long long)di.sm_count * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

struct RedPlan {
  int nrb, chunks;          // row blocks x 64-channel chunks
  long long rows_per_block;
  // workspace: [ticket counters | bound bits | cbound | sums 2C floats | partials]. The first three regions hold state that
  // persists between calls (self-clearing counters / accumulators) and therefore sit at FIXED offsets (64 chunks max),
  // whatever (M, C) the call has: one workspace serves every BatchNorm layer of an encoder.
  size_t off_sums, off_part, off_bound, bytes;
};
// per_sm = resident CTAs per SM of the reduction kernel that will run (its __launch_bounds__ minimum): the grid is ONE full
// wave. (A cap of 4 per SM under a residency of 3 — the backward modes — ran as 1 1/3 waves: 2.9 TB/s on the 56x56x64 layer.)
int red_plan(long long M, int C, RedPlan* pl, int per_sm = 4) {
  if (M < 1 || C < 64 || (C & 63) || C > 4096) return MLA_E_SHAPE;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  pl->chunks = C / 64;
  long long nrb = (M + 127) / 128;                                   // >= 8 rows per thread
  const long long cap = max(1, per_sm * di.sm_count / pl->chunks);
  if (nrb > cap) nrb = cap;
  pl->rows_per_block = (M + nrb - 1) / nrb;
  pl->nrb = (int)((M + pl->rows_per_block - 1) / pl->rows_per_block);
  pl->off_bound = 256;                      // 64 x uint bound bits, then 64 x float published bounds
  pl->off_sums = 768;
  pl->off_part = pl->off_sums + mla::align_up(2 * (size_t)C * sizeof(float), 256);
  pl->bytes = pl->off_part + (size_t)pl->chunks * pl->nrb * 128 * sizeof(double);
  return 0;
}

}  // namespace

#define MLA_LAUNCH_CHECK()            \
  MLA_CUDA_TRY(cudaGetLastError());   \
  mla::count_launch()

static int stem_im2col_impl(const float* in, float* col, void* col16, void* col16b, int N, int T, long long sB, long long sT,
                            long long sC, int Cin, int H, int W, int R, int S, int stride, int pad, int Kp, void* stream) {
  if (!in || N < 1 || T < 1 || Cin < 1 || Kp < R * S * Cin || (Kp & 3)) return MLA_E_BADARG;
  const int OH = (H + 2 * pad - R) / stride + 1, OW = (W + 2 * pad - S) / stride + 1;
  if (OH < 1 || OW < 1) return MLA_E_SHAPE;
  const size_t smem = ((((size_t)Cin * R * (W + 2 * pad + 2) + 3) & ~(size_t)3) + (size_t)Kp) * sizeof(float);
  if (smem > 200 * 1024) return MLA_E_SHAPE;
  static std::atomic<int> cfg{0};
  if (!cfg.load()) {
    MLA_CUDA_TRY(cudaFuncSetAttribute(stem_im2col_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MLA_CUDA_TRY(cudaFuncSetAttribute(stem_im2col_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cfg.store(1);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (col16 != nullptr)
    stem_im2col_kernel<true><<<N * OH, 256, smem, st>>>(in, nullptr, static_cast<uint2*>(col16), static_cast<uint2*>(col16b), T,
                                                       sB, sT, sC, Cin, H, W, OH, OW, R, S, stride, pad, R * S * Cin, Kp);
  else
    stem_im2col_kernel<false><<<N * OH, 256, smem, st>>>(in, col, nullptr, nullptr, T, sB, sT, sC, Cin, H, W, OH, OW, R, S,
                                                        stride, pad, R * S * Cin, Kp);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_stem_im2col(const float* in, float* col, int N, int T, long long sB, long long sT, long long sC, int Cin,
                               int H, int W, int R, int S, int stride, int pad, int Kp, void* stream) {
  if (!col || !mla::aligned16(col)) return MLA_E_BADARG;
  return stem_im2col_impl(in, col, nullptr, nullptr, N, T, sB, sT, sC, Cin, H, W, R, S, stride, pad, Kp, stream);
}

extern "C" int mla_stem_im2col16(const float* in, void* col16, void* col16b, int N, int T, long long sB, long long sT,
                                 long long sC, int Cin, int H, int W, int R, int S, int stride, int pad, int Kp, void* stream) {
  if (!col16 || (reinterpret_cast<uintptr_t>(col16) & 7u) || (reinterpret_cast<uintptr_t>(col16b) & 7u) || (Kp & 63))
    return MLA_E_BADARG;
  return stem_im2col_impl(in, nullptr, col16, col16b, N, T, sB, sT, sC, Cin, H, W, R, S, stride, pad, Kp, stream);
}

extern "C" int mla_round_tf32(const float* src, float* dst, long long n, void* stream) {
  if (!src || !dst || n < 4 || (n & 3) || !mla::aligned16(src) || !mla::aligned16(dst)) return MLA_E_BADARG;
  round_tf32_kernel<<<ew_grid(n / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n / 4);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_cast16(const float* src, void* dst16, long long n, int bf16, void* stream) {
  if (!src || !dst16 || n < 4 || (n & 3) || !mla::aligned16(src) || (reinterpret_cast<uintptr_t>(dst16) & 7u)) return MLA_E_BADARG;
  cast16_kernel<<<ew_grid(n / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<uint2*>(dst16), n / 4, bf16);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_filter_transpose16(const float* w, void* wt16, int Cout, int RS, int Cin, int bf16, void* stream) {
  if (!w || !wt16 || Cout < 1 || RS < 1 || Cin < 1) return MLA_E_BADARG;
  dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, RS);
  filter_transpose16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, static_cast<unsigned short*>(wt16), Cout, RS,
                                                                              Cin, bf16);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_filter_transpose16_batch(const float* flat, void* flat_t16, const void* seg_table, int nseg, int ntiles,
                                            int bf16, void* stream) {
  if (!flat || !flat_t16 || !seg_table || nseg < 1 || ntiles < 1) return MLA_E_BADARG;
  filter_transpose16_batch_kernel<<<ntiles, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      flat, static_cast<unsigned short*>(flat_t16), static_cast<const TransposeSeg*>(seg_table), nseg, bf16);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_pad_rows(const float* src, float* dst, int rows, int k, int kp, int unpad, void* stream) {
  if (!src || !dst || rows < 1 || k < 1 || kp < k) return MLA_E_BADARG;
  pad_rows_kernel<<<ew_grid((long long)rows * kp, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, rows, k, kp,
                                                                                                  unpad);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t mla_bn_workspace_bytes(long long M, int C) {
  RedPlan pl;
  if (red_plan(M, C, &pl) != 0) return 0;
  return pl.bytes;
}

extern "C" int mla_bn_train_stats(const float* y, long long M, int C, const float* gamma, const float* beta,
                                  float* running_mean, float* running_var, float momentum, float eps, float* mean_out,
                                  float* invstd_out, float* scale_out, float* shift_out, void* ws, size_t ws_bytes,
                                  void* stream) {
  if (!y || !gamma || !beta || !mean_out || !invstd_out || !scale_out || !shift_out) return MLA_E_BADARG;
  RedPlan pl;
  int rc = red_plan(M, C, &pl);
  if (rc) return rc;
  if (!ws || ws_bytes < pl.bytes) return MLA_E_WORKSPACE;
  char* base = static_cast<char*>(ws);
  BnFinal f{};
  f.M = M; f.gamma = gamma; f.beta = beta; f.running_mean = running_mean; f.running_var = running_var;
  f.momentum = momentum; f.eps = eps; f.mean_out = mean_out; f.invstd_out = invstd_out; f.scale_out = scale_out;
  f.shift_out = shift_out;
  channel_reduce_kernel<0><<<dim3(pl.nrb, pl.chunks), kRedThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      y, nullptr, nullptr, nullptr, nullptr, M, C, pl.rows_per_block, reinterpret_cast<double*>(base + pl.off_part),
      reinterpret_cast<unsigned int*>(base), f, nullptr);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_bn_stats_from_partials(const float* part, int ntiles, long long M, int C, const float* gamma,
                                          const float* beta, float* running_mean, float* running_var, float momentum,
                                          float eps, float* mean_out, float* invstd_out, float* scale_out, float* shift_out,
                                          void* ws, size_t ws_bytes, void* stream) {
  if (!part || !gamma || !beta || !mean_out || !invstd_out || !scale_out || !shift_out || ntiles < 1 || M < 1)
    return MLA_E_BADARG;
  RedPlan pl;
  int rc = red_plan(ntiles, C, &pl);
  if (rc) return rc;
  if (!ws || ws_bytes < pl.bytes) return MLA_E_WORKSPACE;
  char* base = static_cast<char*>(ws);
  BnFinal f{};
  f.M = M; f.gamma = gamma; f.beta = beta; f.running_mean = running_mean; f.running_var = running_var;
  f.momentum = momentum; f.eps = eps; f.mean_out = mean_out; f.invstd_out = invstd_out; f.scale_out = scale_out;
  f.shift_out = shift_out;
  channel_reduce_kernel<2><<<dim3(pl.nrb, pl.chunks), kRedThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      part, nullptr, nullptr, nullptr, nullptr, ntiles, C, pl.rows_per_block, reinterpret_cast<double*>(base + pl.off_part),
      reinterpret_cast<unsigned int*>(base), f, nullptr);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean,
                                  const float* running_var, float eps, int C, float* scale_out, float* shift_out,
                                  void* stream) {
  if (!gamma || !beta || !running_mean || !running_var || !scale_out || !shift_out || C < 1) return MLA_E_BADARG;
  bn_eval_coeffs_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(gamma, beta, running_mean,
                                                                                       running_var, eps, C, scale_out,
                                                                                       shift_out);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_bn_apply_ex(const float* y, const float* scale, const float* shift, const float* res,
                               const float* res_scale, const float* res_shift, int relu, float* out,
                               unsigned int* relu_mask, void* out16, void* out16b, long long M, int C, void* stream) {
  if (!y || !scale || !shift || (!out && !out16 && !out16b) || M < 1 || C < 4 || (C & 3)) return MLA_E_BADARG;
  if (relu_mask != nullptr && (C & 31)) return MLA_E_SHAPE;    // whole mask words per row group
  const long long n4 = M * (C / 4);
  bn_apply_kernel<<<ew_grid((n4 + 3) / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y, scale, shift, res, res_scale, res_shift, relu, out, n4, C / 4, relu_mask, static_cast<uint2*>(out16),
      static_cast<uint2*>(out16b));
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_bn_apply_mask(const float* y, const float* scale, const float* shift, const float* res,
                                 const float* res_scale, const float* res_shift, int relu, float* out,
                                 unsigned int* relu_mask, long long M, int C, void* stream) {
  if (!out) return MLA_E_BADARG;
  return mla_bn_apply_ex(y, scale, shift, res, res_scale, res_shift, relu, out, relu_mask, nullptr, nullptr, M, C, stream);
}

extern "C" int mla_bn_apply(const float* y, const float* scale, const float* shift, const float* res, const float* res_scale,
                            const float* res_shift, int relu, float* out, long long M, int C, void* stream) {
  return mla_bn_apply_mask(y, scale, shift, res, res_scale, res_shift, relu, out, nullptr, M, C, stream);
}

static int bn_backward_impl(const float* dz, const float* z, const unsigned int* mask, const float* y, const float* mean,
                            const float* invstd, const float* gamma, long long M, int C, float* dgamma, float* dbeta,
                            float* dy, float* g_out, void* ws, size_t ws_bytes, void* stream, void* dy16 = nullptr,
                            float* gscale = nullptr) {
  if (!dz || !y || !mean || !invstd || !gamma || (!dy && !dy16)) return MLA_E_BADARG;
  if (gscale != nullptr && !dy16) return MLA_E_BADARG;
  RedPlan pl;
  int rc = red_plan(M, C, &pl, 3);
  if (rc) return rc;
  if (!ws || ws_bytes < pl.bytes) return MLA_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(ws);
  float* sums = reinterpret_cast<float*>(base + pl.off_sums);
  BnFinal f{};
  f.M = M; f.dgamma = dgamma; f.dbeta = dbeta; f.sums = sums;
  float* cbound = nullptr;
  if (gscale != nullptr) {           // the bound accumulators live in the (zero-initialised, self-clearing) workspace
    f.gamma = gamma;
    f.bound_bits = reinterpret_cast<unsigned int*>(base + pl.off_bound);
    cbound = reinterpret_cast<float*>(base + pl.off_bound) + 64;
    f.cbound = cbound;
  }
  channel_reduce_kernel<1><<<dim3(pl.nrb, pl.chunks), kRedThreads, 0, st>>>(
      y, dz, z, mean, invstd, M, C, pl.rows_per_block, reinterpret_cast<double*>(base + pl.off_part),
      reinterpret_cast<unsigned int*>(base), f, mask);
  MLA_LAUNCH_CHECK();
  const long long n4 = M * (C / 4);
  bn_bwd_apply_kernel<<<ew_grid(n4, 256), 256, 0, st>>>(dz, z, y, mean, invstd, gamma, sums, 1.f / (float)M, dy, g_out, n4,
                                                       C / 4, mask, static_cast<uint2*>(dy16), cbound, pl.chunks, gscale);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_bn_backward(const float* dz, const float* z, const float* y, const float* mean, const float* invstd,
                               const float* gamma, long long M, int C, float* dgamma, float* dbeta, float* dy, float* g_out,
                               void* ws, size_t ws_bytes, void* stream) {
  return bn_backward_impl(dz, z, nullptr, y, mean, invstd, gamma, M, C, dgamma, dbeta, dy, g_out, ws, ws_bytes, stream);
}

extern "C" int mla_bn_backward_ex(const float* dz, const float* z, const unsigned int* relu_mask, const float* y,
                                  const float* mean, const float* invstd, const float* gamma, long long M, int C,
                                  float* dgamma, float* dbeta, float* dy, void* dy16, float* g_out, void* ws,
                                  size_t ws_bytes, void* stream) {
  return bn_backward_impl(dz, z, relu_mask, y, mean, invstd, gamma, M, C, dgamma, dbeta, dy, g_out, ws, ws_bytes, stream,
                          dy16);
}

// BN backward whose dy leaves as fp16 multiplied by a power of two F chosen from the data (gscale[0] = F, gscale[1] = 1 / F,
// device scalars the consuming dgrad / wgrad multiply their accumulators by): the operand keeps TF32's 10-bit mantissa.
extern "C" int mla_bn_backward_f16(const float* dz, const unsigned int* relu_mask, const float* y, const float* mean,
                                   const float* invstd, const float* gamma, long long M, int C, float* dgamma, float* dbeta,
                                   void* dy16, float* g_out, float* gscale, void* ws, size_t ws_bytes, void* stream) {
  if (!dy16 || !gscale) return MLA_E_BADARG;
  return bn_backward_impl(dz, nullptr, relu_mask, y, mean, invstd, gamma, M, C, dgamma, dbeta, nullptr, g_out, ws, ws_bytes,
                          stream, dy16, gscale);
}

extern "C" int mla_bn_backward_mask(const float* dz, const unsigned int* relu_mask, const float* y, const float* mean,
                                    const float* invstd, const float* gamma, long long M, int C, float* dgamma,
                                    float* dbeta, float* dy, float* g_out, void* ws, size_t ws_bytes, void* stream) {
  if (!relu_mask) return MLA_E_BADARG;
  return bn_backward_impl(dz, nullptr, relu_mask, y, mean, invstd, gamma, M, C, dgamma, dbeta, dy, g_out, ws, ws_bytes,
                          stream);
}

extern "C" int mla_bn_relu_maxpool_ex(const float* y, const float* scale, const float* shift, float* out, void* out16,
                                      void* out16b, unsigned char* idx, int N, int H, int W, int C, void* stream) {
  if (!y || !scale || !shift || !out || !idx || (C & 3)) return MLA_E_BADARG;
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)N * OH * OW * (C / 4);
  bn_relu_maxpool_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y, scale, shift, out, idx, N, H, W, C / 4, OH, OW, static_cast<uint2*>(out16), static_cast<uint2*>(out16b));
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_bn_relu_maxpool(const float* y, const float* scale, const float* shift, float* out, unsigned char* idx,
                                   int N, int H, int W, int C, void* stream) {
  return mla_bn_relu_maxpool_ex(y, scale, shift, out, nullptr, nullptr, idx, N, H, W, C, stream);
}

extern "C" int mla_maxpool_relu_backward(const float* dp, const float* p, const unsigned char* idx, float* g, int N, int H,
                                         int W, int C, void* stream) {
  if (!dp || !p || !idx || !g || (C & 3)) return MLA_E_BADARG;
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)N * H * W * (C / 4);
  maxpool_relu_bwd_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dp, p, idx, g, N, H, W, C / 4,
                                                                                             OH, OW);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_bn_relu_maxpool16(const void* y16, const float* scale, const float* shift, float* out, void* out16,
                                     unsigned char* idx, int N, int H, int W, int C, void* stream) {
  if (!y16 || !scale || !shift || (!out && !out16) || !idx || N < 1 || H < 1 || W < 1) return MLA_E_BADARG;
  if (C != 64) return MLA_E_SHAPE;
  if (!mla::aligned16(y16) || !mla::aligned16(out) || !mla::aligned16(out16) || (reinterpret_cast<uintptr_t>(idx) & 7u))
    return MLA_E_BADARG;
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  bn_relu_maxpool16_kernel<<<N * OH, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(y16), scale, shift, out, static_cast<uint4*>(out16), idx, H, W, OH, OW);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_pool_bn_backward_f16(const float* dp, const unsigned char* idx, const void* y16, const float* mean,
                                        const float* invstd, const float* gamma, int N, int H, int W, int C, float* dgamma,
                                        float* dbeta, void* dy16, float* gscale, void* ws, size_t ws_bytes, void* stream) {
  if (!dp || !idx || !y16 || !mean || !invstd || !gamma || !dy16 || !gscale || N < 1 || H < 1 || W < 1) return MLA_E_BADARG;
  if (C != 64) return MLA_E_SHAPE;
  if (!mla::aligned16(dp) || !mla::aligned16(y16) || !mla::aligned16(dy16) || (reinterpret_cast<uintptr_t>(idx) & 7u))
    return MLA_E_BADARG;
  const long long M = (long long)N * H * W;
  const int PH = (H + 2 - 3) / 2 + 1, PW = (W + 2 - 3) / 2 + 1;
  const long long MP = (long long)N * PH * PW;           // the reduction walks the pooled pixels
  if (MP >= (1LL << 31)) return MLA_E_SHAPE;
  RedPlan pl;
  int rc = red_plan(MP, C, &pl, 3);
  if (rc) return rc;
  if (!ws || ws_bytes < pl.bytes) return MLA_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(ws);
  float* sums = reinterpret_cast<float*>(base + pl.off_sums);
  BnFinal f{};
  f.M = M; f.dgamma = dgamma; f.dbeta = dbeta; f.sums = sums; f.gamma = gamma;
  f.bound_bits = reinterpret_cast<unsigned int*>(base + pl.off_bound);
  float* cbound = reinterpret_cast<float*>(base + pl.off_bound) + 64;
  f.cbound = cbound;
  f.pool.dp = dp; f.pool.idx = idx; f.pool.y16 = static_cast<const __half*>(y16);
  f.pool.H = H; f.pool.W = W; f.pool.PH = PH; f.pool.PW = PW;
  channel_reduce_kernel<3><<<dim3(pl.nrb, pl.chunks), kRedThreads, 0, st>>>(
      nullptr, nullptr, nullptr, mean, invstd, MP, C, pl.rows_per_block, reinterpret_cast<double*>(base + pl.off_part),
      reinterpret_cast<unsigned int*>(base), f, nullptr);
  MLA_LAUNCH_CHECK();
  pool_bn_bwd_apply_kernel<<<N * H, 256, 0, st>>>(f.pool, mean, invstd, gamma, sums, 1.f / (float)M,
                                                 static_cast<uint4*>(dy16), cbound, pl.chunks, gscale);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_avgpool_forward(const float* fm, float* feat, int B, int rows, int C, void* stream) {
  if (!fm || !feat || B < 1 || rows < 1 || C < 1) return MLA_E_BADARG;
  avgpool_fwd_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(fm, feat, rows, C);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_avgpool_backward(const float* dfeat, float* dfm, int B, int rows, int C, void* stream) {
  if (!dfeat || !dfm || B < 1 || rows < 1 || (C & 3)) return MLA_E_BADARG;
  const long long total = (long long)B * rows * (C / 4);
  avgpool_bwd_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dfeat, dfm, B, rows, C / 4);
  MLA_LAUNCH_CHECK();
  return 0;
}
