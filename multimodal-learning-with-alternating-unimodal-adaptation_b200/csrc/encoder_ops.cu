// Memory-bound pieces of the ResNet-18 encoders around the tcgen05 convolutions (NHWC fp32):
//   stem im2col (7x7/2 on the raw NCHW input -> [M][Kp] rows for a tensor-core GEMM)
//   BatchNorm2d training statistics / running-stat update        backbone.py:40,44,150 (nn.BatchNorm2d)
//   BN apply (+ReLU) (+residual with its own BN affine)           backbone.py:41-50
//   BN+ReLU+MaxPool 3x3/2 fused for the stem                      backbone.py:150-152
//   their backward passes, global average pool fwd/bwd            basic_model.py:61-65
// All reductions are two-stage with a fixed order (deterministic); per-channel sums are carried
// in double across threads/blocks so mean/variance match torch's to fp32 rounding.
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

// ------------------------------------------------------------------------------ stem im2col
// col[m][k], m = (n, oh, ow), k = (r*S + s)*Cin + ci for k < R*S*Cin, zero for the pad columns.
// Input element (n, ci, h, w) lives at in[(n / T)*sB + (n % T)*sT + ci*sC + h*W + w] (raw NCHW /
// NCTHW tensors: the frame fold of backbone.py:144-147 is pure index arithmetic here).
// One block per output row (n, oh): the R input rows of every channel are staged in shared memory
// (coalesced, zero-padded left/right/top/bottom), a k -> offset table replaces the per-element
// div/mod, and the row's OW x Kp outputs leave as coalesced float4 stores. HBM-bound by the col write.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ in, float* __restrict__ col, int T,
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
    sm[i] = tf32r(v);
  }
  __syncthreads();
  const int K4 = Kp >> 2;
  float* dst = col + (long long)blockIdx.x * OW * Kp;
  for (int i = threadIdx.x; i < OW * K4; i += blockDim.x) {
    const int ow = i / K4, k = (i - ow * K4) * 4;
    const int base = ow * stride;
    const int4 o = *reinterpret_cast<const int4*>(koff + k);
    float4 v;
    v.x = o.x >= 0 ? sm[o.x + base] : 0.f;
    v.y = o.y >= 0 ? sm[o.y + base] : 0.f;
    v.z = o.z >= 0 ? sm[o.z + base] : 0.f;
    v.w = o.w >= 0 ? sm[o.w + base] : 0.f;
    st4(dst + 4 * (long long)i, v);
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
// MODE 2: x = per-tile partial sums [rows = tiles][2][C] written by the fprop epilogue: a = x[r][0][c],
//         b = x[r][1][c]; finalised like MODE 0 (f.M = the real number of pixels).
struct BnFinal {
  long long M;
  // MODE 0
  const float* gamma; const float* beta;
  float* running_mean; float* running_var;
  float momentum, eps;
  float* mean_out; float* invstd_out; float* scale_out; float* shift_out;
  // MODE 1
  float* dgamma; float* dbeta; float* sums;   // sums [2][C]: dbeta, dgamma (read by the apply pass)
};

template <int MODE>
__global__ void __launch_bounds__(kRedThreads) channel_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                                     const float* __restrict__ z,
                                                                     const float* __restrict__ mean,
                                                                     const float* __restrict__ invstd, long long M, int C,
                                                                     long long rows_per_block, double* __restrict__ part,
                                                                     unsigned int* __restrict__ counters, BnFinal f) {
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
    if (MODE == 1) { mu = ld4(mean + c0); is = ld4(invstd + c0); }
    long long r = r0 + ty;
    while (r < r1) {
      float fa[4] = {0, 0, 0, 0}, fb[4] = {0, 0, 0, 0};
      for (int it = 0; it < (MODE == 2 ? 2 : 32) && r < r1; ++it, r += 16) {   // a few rows in fp32, then flush to double
        const float4 v = ld4(x + r * (MODE == 2 ? 2 * C : C) + c0);
        if (MODE == 2) {
          const float4 q = ld4(x + r * 2 * C + C + c0);
          fa[0] += v.x; fa[1] += v.y; fa[2] += v.z; fa[3] += v.w;
          fb[0] += q.x; fb[1] += q.y; fb[2] += q.z; fb[3] += q.w;
        } else if (MODE == 0) {
          fa[0] += v.x; fa[1] += v.y; fa[2] += v.z; fa[3] += v.w;
          fb[0] = fmaf(v.x, v.x, fb[0]); fb[1] = fmaf(v.y, v.y, fb[1]);
          fb[2] = fmaf(v.z, v.z, fb[2]); fb[3] = fmaf(v.w, v.w, fb[3]);
        } else {
          float4 g = ld4(dz + r * C + c0);
          if (z != nullptr) {
            const float4 zz = ld4(z + r * C + c0);
            g.x = zz.x > 0.f ? g.x : 0.f; g.y = zz.y > 0.f ? g.y : 0.f;
            g.z = zz.z > 0.f ? g.z : 0.f; g.w = zz.w > 0.f ? g.w : 0.f;
          }
          fa[0] += g.x; fa[1] += g.y; fa[2] += g.z; fa[3] += g.w;
          fb[0] = fmaf(g.x, (v.x - mu.x) * is.x, fb[0]); fb[1] = fmaf(g.y, (v.y - mu.y) * is.y, fb[1]);
          fb[2] = fmaf(g.z, (v.z - mu.z) * is.z, fb[2]); fb[3] = fmaf(g.w, (v.w - mu.w) * is.w, fb[3]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] += (double)fa[q]; b[q] += (double)fb[q]; }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      s_acc[ty][0][4 * tx + q] = a[q];
      s_acc[ty][1][4 * tx + q] = b[q];
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
  if (threadIdx.x < 128) {
    const double* pp = part + (size_t)chunk * nrb * 128 + threadIdx.x;
    double sum = 0;
#pragma unroll 8
    for (int blk = 0; blk < nrb; ++blk) sum += __ldcg(pp + (size_t)blk * 128);   // block order: deterministic
    s_fin[threadIdx.x >> 6][threadIdx.x & 63] = sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) counters[chunk] = 0;   // leave the workspace reusable
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

// out = relu?( y*scale + shift + (res ? res*rscale + rshift : 0) )
__global__ void bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                                const float* __restrict__ res, const float* __restrict__ rscale,
                                const float* __restrict__ rshift, int relu, float* __restrict__ out, long long n4, int C4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 v = ld4(y + 4 * i), sc = ld4(scale + c), sh = ld4(shift + c);
    float4 o = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
    if (res != nullptr) {
      float4 r = ld4(res + 4 * i);
      if (rscale != nullptr) {
        const float4 rs = ld4(rscale + c), rb = ld4(rshift + c);
        r = make_float4(fmaf(r.x, rs.x, rb.x), fmaf(r.y, rs.y, rb.y), fmaf(r.z, rs.z, rb.z), fmaf(r.w, rs.w, rb.w));
      }
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    st4(out + 4 * i, tf32r4(o));   // consumers are tcgen05 convolutions (and the residual add / pooling)
  }
}

// dy = gamma*invstd * (g - dbeta/M - xhat*dgamma/M), g = dz*(z>0); optionally also writes g.
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ z, const float* __restrict__ y,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ sums, float inv_m,
                                    float* __restrict__ dy, float* __restrict__ g_out, long long n4, int C4) {
  const int C = C4 * 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    float4 g = ld4(dz + 4 * i);
    if (z != nullptr) {
      const float4 zz = ld4(z + 4 * i);
      g.x = zz.x > 0.f ? g.x : 0.f; g.y = zz.y > 0.f ? g.y : 0.f;
      g.z = zz.z > 0.f ? g.z : 0.f; g.w = zz.w > 0.f ? g.w : 0.f;
    }
    if (g_out != nullptr) st4(g_out + 4 * i, g);
    const float4 v = ld4(y + 4 * i), mu = ld4(mean + c), is = ld4(invstd + c), ga = ld4(gamma + c);
    const float4 db = ld4(sums + c), dg = ld4(sums + C + c);
    float4 o;
    o.x = ga.x * is.x * (g.x - db.x * inv_m - (v.x - mu.x) * is.x * dg.x * inv_m);
    o.y = ga.y * is.y * (g.y - db.y * inv_m - (v.y - mu.y) * is.y * dg.y * inv_m);
    o.z = ga.z * is.z * (g.z - db.z * inv_m - (v.z - mu.z) * is.z * dg.z * inv_m);
    o.w = ga.w * is.w * (g.w - db.w * inv_m - (v.w - mu.w) * is.w * dg.w * inv_m);
    st4(dy + 4 * i, tf32r4(o));    // dy only feeds dgrad / wgrad
  }
}

// ------------------------------------------------------------------ stem BN+ReLU+MaxPool 3x3/2 p1
__global__ void bn_relu_maxpool_kernel(const float* __restrict__ y, const float* __restrict__ scale,
                                       const float* __restrict__ shift, float* __restrict__ out,
                                       unsigned char* __restrict__ idx, int N, int H, int W, int C4, int OH, int OW) {
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
  size_t off_sums, off_part, bytes;   // workspace layout: [counters | sums 2C floats | partials]
};
int red_plan(long long M, int C, RedPlan* pl) {
  if (M < 1 || C < 64 || (C & 63) || C > 4096) return MLA_E_SHAPE;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  pl->chunks = C / 64;
  long long nrb = (M + 127) / 128;                                   // >= 8 rows per thread
  const long long cap = max(1, 2 * di.sm_count / pl->chunks);
  if (nrb > cap) nrb = cap;
  pl->rows_per_block = (M + nrb - 1) / nrb;
  pl->nrb = (int)((M + pl->rows_per_block - 1) / pl->rows_per_block);
  pl->off_sums = mla::align_up((size_t)pl->chunks * sizeof(unsigned int), 256);
  pl->off_part = pl->off_sums + mla::align_up(2 * (size_t)C * sizeof(float), 256);
  pl->bytes = pl->off_part + (size_t)pl->chunks * pl->nrb * 128 * sizeof(double);
  return 0;
}

}  // namespace

#define MLA_LAUNCH_CHECK()            \
  MLA_CUDA_TRY(cudaGetLastError());   \
  mla::count_launch()

extern "C" int mla_stem_im2col(const float* in, float* col, int N, int T, long long sB, long long sT, long long sC, int Cin,
                               int H, int W, int R, int S, int stride, int pad, int Kp, void* stream) {
  if (!in || !col || N < 1 || T < 1 || Cin < 1 || Kp < R * S * Cin || (Kp & 3) || !mla::aligned16(col)) return MLA_E_BADARG;
  const int OH = (H + 2 * pad - R) / stride + 1, OW = (W + 2 * pad - S) / stride + 1;
  if (OH < 1 || OW < 1) return MLA_E_SHAPE;
  const size_t smem = ((((size_t)Cin * R * (W + 2 * pad + 2) + 3) & ~(size_t)3) + (size_t)Kp) * sizeof(float);
  if (smem > 200 * 1024) return MLA_E_SHAPE;
  static std::atomic<int> cfg{0};
  if (!cfg.load()) {
    MLA_CUDA_TRY(cudaFuncSetAttribute(stem_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cfg.store(1);
  }
  stem_im2col_kernel<<<N * OH, 256, smem, static_cast<cudaStream_t>(stream)>>>(in, col, T, sB, sT, sC, Cin, H, W, OH, OW, R, S,
                                                                              stride, pad, R * S * Cin, Kp);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_round_tf32(const float* src, float* dst, long long n, void* stream) {
  if (!src || !dst || n < 4 || (n & 3) || !mla::aligned16(src) || !mla::aligned16(dst)) return MLA_E_BADARG;
  round_tf32_kernel<<<ew_grid(n / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n / 4);
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
      reinterpret_cast<unsigned int*>(base), f);
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
      reinterpret_cast<unsigned int*>(base), f);
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

extern "C" int mla_bn_apply(const float* y, const float* scale, const float* shift, const float* res, const float* res_scale,
                            const float* res_shift, int relu, float* out, long long M, int C, void* stream) {
  if (!y || !scale || !shift || !out || M < 1 || C < 4 || (C & 3)) return MLA_E_BADARG;
  const long long n4 = M * (C / 4);
  bn_apply_kernel<<<ew_grid(n4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, scale, shift, res, res_scale,
                                                                                  res_shift, relu, out, n4, C / 4);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_bn_backward(const float* dz, const float* z, const float* y, const float* mean, const float* invstd,
                               const float* gamma, long long M, int C, float* dgamma, float* dbeta, float* dy, float* g_out,
                               void* ws, size_t ws_bytes, void* stream) {
  if (!dz || !y || !mean || !invstd || !gamma || !dy) return MLA_E_BADARG;
  RedPlan pl;
  int rc = red_plan(M, C, &pl);
  if (rc) return rc;
  if (!ws || ws_bytes < pl.bytes) return MLA_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(ws);
  float* sums = reinterpret_cast<float*>(base + pl.off_sums);
  BnFinal f{};
  f.M = M; f.dgamma = dgamma; f.dbeta = dbeta; f.sums = sums;
  channel_reduce_kernel<1><<<dim3(pl.nrb, pl.chunks), kRedThreads, 0, st>>>(
      y, dz, z, mean, invstd, M, C, pl.rows_per_block, reinterpret_cast<double*>(base + pl.off_part),
      reinterpret_cast<unsigned int*>(base), f);
  MLA_LAUNCH_CHECK();
  const long long n4 = M * (C / 4);
  bn_bwd_apply_kernel<<<ew_grid(n4, 256), 256, 0, st>>>(dz, z, y, mean, invstd, gamma, sums, 1.f / (float)M, dy, g_out, n4,
                                                       C / 4);
  MLA_LAUNCH_CHECK();
  return 0;
}

extern "C" int mla_bn_relu_maxpool(const float* y, const float* scale, const float* shift, float* out, unsigned char* idx,
                                   int N, int H, int W, int C, void* stream) {
  if (!y || !scale || !shift || !out || !idx || (C & 3)) return MLA_E_BADARG;
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)N * OH * OW * (C / 4);
  bn_relu_maxpool_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, scale, shift, out, idx, N, H,
                                                                                            W, C / 4, OH, OW);
  MLA_LAUNCH_CHECK();
  return 0;
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
