// Dataset tuple producer for the visual modality — replaces, per batch, what the reference does per
// sample on the CPU in dataset/dataset.py:123-161 (AVDataset.__getitem__; the same Compose in :448-480 / :753-803):
//     PIL image -> RandomResizedCrop(224) [+ RandomHorizontalFlip]  |  Resize((224, 224))      (torchvision on PIL)
//               -> ToTensor() -> Normalize(mean, std) -> frames stacked on a new dim 1: [3, T, 224, 224]
// and the deterministic transform of the other datasets (CAVDataset :251-256, M3AEDataset test mode :414-421):
//     Resize(s, BICUBIC) [shorter side -> s] -> CenterCrop(s) -> ToTensor() -> Normalize(mean, std)
// Input: the decoded uint8 HWC RGB frames of a batch packed into one device buffer + one descriptor per frame (geometry,
// crop box, flip, destination slot). Output: the model's input tensor [B, 3, T, OH, OW] fp32, BIT-IDENTICAL to the
// torchvision / Pillow path:
//   * Pillow's Image.resize(BILINEAR) (what torchvision's Resize / resized_crop call on PIL images) is a two-pass separable
//     convolution with an antialiasing support that grows with the down-scaling factor (libImaging/Resample.c): per output
//     coordinate, a window [xmin, xmin + n) of normalised filter weights (BILINEAR: triangle, support 1; BICUBIC: Keys
//     a = -0.5, support 2) computed in double precision, converted to 22-bit
//     fixed point ((int)(0.5 + k * 2^22)), accumulated in int32 from 2^21, shifted and clipped to 8 bits — horizontally into
//     an 8-bit intermediate image first, then vertically. frame_coeffs_kernel restates precompute_coeffs() +
//     normalize_coeffs_8bpc() operation by operation in IEEE double (explicit _rn intrinsics: no FMA contraction);
//     frame_resample_kernel restates ImagingResampleHorizontal_8bpc / Vertical_8bpc (8-bit intermediate values included).
//   * ToTensor = float(u8) / 255 (IEEE division), Normalize = (x - mean) / std in fp32, in that order.
// The crop is a window of the source frame (PIL crop copies pixels, resize then works on the copy: same values).
#include <math.h>
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int kMaxK = 32;          // taps per output coordinate: ceil(scale) * 2 + 1 <= 32  ->  crop / out < 15.5
constexpr int kPrec = 32 - 8 - 2;  // PRECISION_BITS of Resample.c
constexpr int kDescInts = 14;      // {src offset lo, hi, H, W, top, left, crop_h, crop_w, flip, out slot, RH, RW, oy, ox}

// (RH, RW) = size the crop window is RESAMPLED to; the [OH, OW] output is the window of that image starting at (oy, ox):
// Resize((OH, OW)) / RandomResizedCrop: RH = OH, RW = OW, oy = ox = 0;  Resize(s) + CenterCrop(s): the shorter side -> s,
// (oy, ox) = the centre-crop origin (only the output window is ever computed; its pixels do not depend on the rest).
struct Desc {
  long long off;
  int H, W, top, left, ch, cw, flip, slot, RH, RW, oy, ox;
};
__device__ __forceinline__ Desc load_desc(const int* __restrict__ d, int f) {
  const int* p = d + (size_t)f * kDescInts;
  Desc r;
  r.off = (long long)(unsigned int)p[0] | ((long long)p[1] << 32);
  r.H = p[2]; r.W = p[3]; r.top = p[4]; r.left = p[5]; r.ch = p[6]; r.cw = p[7]; r.flip = p[8]; r.slot = p[9];
  r.RH = p[10]; r.RW = p[11]; r.oy = p[12]; r.ox = p[13];
  return r;
}

struct Lim {
  int OH, OW, max_crop_h, nslots, bicubic;
  long long src_bytes;
};
// taps per output coordinate: (int)ceil(support) * 2 + 1, support = filter support (1 bilinear, 2 bicubic) * max(in / out, 1)
__host__ __device__ __forceinline__ int taps(int in_size, int out_size, int bicubic) {
  const int fs = bicubic ? 2 : 1;
  const int up = in_size > out_size ? (fs * in_size + out_size - 1) / out_size : fs;       // ceil(fs * in / out)
  return up * 2 + 1;
}
// A descriptor every kernel may act on: the crop window inside the frame, the frame inside the source buffer, the taps
// within kMaxK (ceil(scale) * 2 + 1), the destination slot inside the batch. Invalid frames are skipped by every kernel and
// reported through `status`.
__device__ __forceinline__ bool desc_ok(const Desc& d, const Lim& l) {
  return d.H > 0 && d.W > 0 && d.ch > 0 && d.cw > 0 && d.top >= 0 && d.left >= 0 && d.top + d.ch <= d.H &&
         d.left + d.cw <= d.W && d.ch <= l.max_crop_h && d.slot >= 0 && d.slot < l.nslots && d.off >= 0 &&
         d.off + (long long)d.H * d.W * 3 <= l.src_bytes && d.oy >= 0 && d.ox >= 0 && d.oy + l.OH <= d.RH &&
         d.ox + l.OW <= d.RW && taps(d.cw, d.RW, l.bicubic) <= kMaxK && taps(d.ch, d.RH, l.bicubic) <= kMaxK;
}

// tables per frame: [axis 0 = horizontal: OW entries | axis 1 = vertical: OH entries], each entry {xmin, n, k[kMaxK]}
constexpr int kEntry = 2 + kMaxK;

__global__ void __launch_bounds__(128) frame_coeffs_kernel(const int* __restrict__ desc, Lim lim, int* __restrict__ tab) {
  const int OH = lim.OH, OW = lim.OW;
  const int f = blockIdx.y, axis = blockIdx.z;
  const int out_size = axis == 0 ? OW : OH;
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= out_size) return;
  const Desc d = load_desc(desc, f);
  if (!desc_ok(d, lim)) return;
  const int in_size = axis == 0 ? d.cw : d.ch;
  const int full = axis == 0 ? d.RW : d.RH;                       // size of the resampled image along this axis
  const int xf = xx + (axis == 0 ? d.ox : d.oy);                  // this output coordinate inside it
  int* e = tab + ((size_t)f * (OW + OH) + (axis == 0 ? 0 : OW) + xx) * kEntry;
  // precompute_coeffs(inSize, in0 = 0, in1 = inSize, outSize, filter): BILINEAR support 1.0, BICUBIC support 2.0 (a = -0.5)
  const double scale = __ddiv_rn((double)in_size, (double)full);
  const double fscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(lim.bicubic ? 2.0 : 1.0, fscale);
  const double ss = __ddiv_rn(1.0, fscale);
  const double center = __dmul_rn((double)xf + 0.5, scale);       // in0 + (xx + 0.5) * scale, in0 = 0
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double k[kMaxK];
  double ww = 0.0;
#pragma unroll 1
  for (int x = 0; x < xmax; ++x) {
    double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
    if (a < 0.0) a = -a;
    double w;
    if (lim.bicubic) {
      // bicubic_filter, a = -0.5:  |x| < 1: ((a + 2) x - (a + 3)) x x + 1;  |x| < 2: (((x - 5) x + 8) x - 4) a
      if (a < 1.0) w = __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, a), 2.5), a), a), 1.0);
      else if (a < 2.0) w = __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(a, 5.0), a), 8.0), a), 4.0), -0.5);
      else w = 0.0;
    } else {
      w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
    }
    k[x] = w;
    ww = __dadd_rn(ww, w);
  }
  e[0] = xmin;
  e[1] = xmax;
#pragma unroll 1
  for (int x = 0; x < kMaxK; ++x) {
    int v = 0;
    if (x < xmax) {
      double c = k[x];
      if (ww != 0.0) c = __ddiv_rn(c, ww);
      // normalize_coeffs_8bpc: (int)(+-0.5 + k * (1 << PRECISION_BITS)), truncation toward zero
      const double s = __dmul_rn(c, (double)(1 << kPrec));
      v = c < 0.0 ? (int)__dadd_rn(-0.5, s) : (int)__dadd_rn(0.5, s);
    }
    e[2 + x] = v;
  }
}

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrec;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// Both passes + ToTensor + Normalize + flip + scatter into [B, 3, T, OH, OW], one thread per output pixel: for every
// vertical tap the thread evaluates the horizontal pass of that input row at its own column — rounded and clipped to 8 bits
// exactly as Pillow stores it in its intermediate image — and feeds it to the vertical sum. The intermediate image is never
// materialised (it would cost a write and a read of crop_h x OW x 3 bytes per frame and a second launch); the horizontal
// sums are re-evaluated once per vertical tap instead (~5 x 5 taps at the usual 1.6x down-scaling), from source bytes that
// neighbouring threads share in L1. NH8: every lane's horizontal window fits 8 taps (coefficients held in registers).
template <bool NH8>
__device__ __forceinline__ void frame_pixel(const unsigned char* __restrict__ src, const Desc& d, const int* __restrict__ he,
                                            const int* __restrict__ ve, int& v0, int& v1, int& v2) {
  const int xmin = he[0], nh = he[1], ymin = ve[0], nv = ve[1];
  int kh[8];
  if (NH8) {
#pragma unroll
    for (int s_ = 0; s_ < 8; ++s_) kh[s_] = s_ < nh ? he[2 + s_] : 0;
  }
  v0 = v1 = v2 = 1 << (kPrec - 1);
  const unsigned char* p = src + d.off + ((long long)(d.top + ymin) * d.W + d.left + xmin) * 3;
  for (int t = 0; t < nv; ++t, p += (long long)d.W * 3) {
    int s0 = 1 << (kPrec - 1), s1 = s0, s2 = s0;
    if (NH8) {
#pragma unroll
      for (int s_ = 0; s_ < 8; ++s_) {
        if (s_ < nh) { s0 += (int)p[3 * s_] * kh[s_]; s1 += (int)p[3 * s_ + 1] * kh[s_]; s2 += (int)p[3 * s_ + 2] * kh[s_]; }
      }
    } else {
      for (int s_ = 0; s_ < nh; ++s_) {
        const int k = he[2 + s_];
        s0 += (int)p[3 * s_] * k; s1 += (int)p[3 * s_ + 1] * k; s2 += (int)p[3 * s_ + 2] * k;
      }
    }
    const int kv = ve[2 + t];
    v0 += clip8(s0) * kv; v1 += clip8(s1) * kv; v2 += clip8(s2) * kv;
  }
}

__global__ void __launch_bounds__(256) frame_resample_kernel(const unsigned char* __restrict__ src, const int* __restrict__ desc,
                                                             const int* __restrict__ tab, Lim lim, int T, float3 mean,
                                                             float3 std, float* __restrict__ out) {
  const int OH = lim.OH, OW = lim.OW;
  const int f = blockIdx.y;
  const Desc d = load_desc(desc, f);
  if (!desc_ok(d, lim)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= OH * OW) return;
  const int yo = i / OW, xo = i - yo * OW;
  const int* he = tab + ((size_t)f * (OW + OH) + xo) * kEntry;
  const int* ve = tab + ((size_t)f * (OW + OH) + OW + yo) * kEntry;
  int v0, v1, v2;
  // Pillow skips a pass whose size does not change; with identity coefficients (1, 0) the pass reproduces its input, so
  // evaluating it is the same thing
  if (taps(d.cw, d.RW, lim.bicubic) <= 8) frame_pixel<true>(src, d, he, ve, v0, v1, v2);
  else frame_pixel<false>(src, d, he, ve, v0, v1, v2);
  const int b = d.slot / T, tt = d.slot - b * T;
  const int xf = d.flip ? OW - 1 - xo : xo;
  const size_t plane = (size_t)OH * OW;
  float* o = out + (((size_t)b * 3) * T + tt) * plane + (size_t)yo * OW + xf;
  // ToTensor: u8 -> float / 255; Normalize: (x - mean) / std — IEEE fp32, no contraction
  o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8(v0), 255.f), mean.x), std.x);
  o[(size_t)T * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8(v1), 255.f), mean.y), std.y);
  o[2 * (size_t)T * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8(v2), 255.f), mean.z), std.z);
}

// the descriptors live in device memory: the first invalid one is reported as status = index + 1 (0 = all valid)
__global__ void frame_check_kernel(const int* __restrict__ desc, int nframes, Lim lim, int* __restrict__ bad) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nframes) return;
  if (!desc_ok(load_desc(desc, f), lim)) atomicMax(bad, f + 1);
}

size_t tab_bytes(int nframes, int OH, int OW) { return mla::align_up((size_t)nframes * (OW + OH) * kEntry * sizeof(int), 256); }

}  // namespace

extern "C" size_t mla_frames_to_batch_workspace_bytes(int nframes, int OH, int OW, int max_crop_h) {
  if (nframes < 1 || OH < 1 || OW < 1 || max_crop_h < 1) return 0;
  return 256 + tab_bytes(nframes, OH, OW);      // the coefficient tables (max_crop_h only bounds the descriptors)
}

extern "C" int mla_frames_to_batch(const unsigned char* src, long long src_bytes, const int* desc, int nframes, int B, int T,
                                   int OH, int OW, int max_crop_h, int filter, const float* mean3, const float* std3,
                                   float* out, int* status, void* ws, size_t ws_bytes, void* stream) {
  if (!src || !desc || !mean3 || !std3 || !out || nframes < 1 || B < 1 || T < 1 || OH < 1 || OW < 1 || max_crop_h < 1 ||
      src_bytes < 1 || (filter != 0 && filter != 1))
    return MLA_E_BADARG;
  if ((long long)OH * OW >= (1LL << 30)) return MLA_E_SHAPE;
  const size_t need = mla_frames_to_batch_workspace_bytes(nframes, OH, OW, max_crop_h);
  if (!ws || ws_bytes < need) return MLA_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(ws);
  int* tab = reinterpret_cast<int*>(base + 256);
  Lim lim;
  lim.OH = OH; lim.OW = OW; lim.max_crop_h = max_crop_h; lim.nslots = B * T; lim.bicubic = filter; lim.src_bytes = src_bytes;
  if (status != nullptr) {
    MLA_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), st));
    frame_check_kernel<<<(nframes + 127) / 128, 128, 0, st>>>(desc, nframes, lim, status);
    MLA_CUDA_TRY(cudaGetLastError());
    mla::count_launch();
  }
  const int omax = OH > OW ? OH : OW;
  frame_coeffs_kernel<<<dim3((omax + 127) / 128, nframes, 2), 128, 0, st>>>(desc, lim, tab);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  frame_resample_kernel<<<dim3((OH * OW + 255) / 256, nframes), 256, 0, st>>>(src, desc, tab, lim, T,
                                                                              make_float3(mean3[0], mean3[1], mean3[2]),
                                                                              make_float3(std3[0], std3[1], std3[2]), out);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}


// ------------------------------------------------------------------------------------------------------------------
// Audio member of the CAV-MAE-style tuples (dataset/dataset.py:281-294 fbank_aug, :301-321): the pre-computed filterbank
// array [T, F] of every sample goes through SpecAugment masks (rows f0..f1 and frames t0..t1 set to 0 BEFORE the
// normalisation, as the reference does), (x - norm_mean) / norm_std, optional additive noise noise * amp / 10 and a roll
// along time — one pass, the parameters drawn on the host in the reference's order (mla_b200/dataset.py).
//   params [B][6] int32 = {f0, f1, t0, t1, roll shift, add noise}; amp [B]; noise [B][T][F] or NULL.
__global__ void __launch_bounds__(256) spec_augment_kernel(const float* __restrict__ x, const int* __restrict__ params,
                                                           const float* __restrict__ amp, const float* __restrict__ noise,
                                                           float mean, float std, int skip_norm, int T, int F,
                                                           float* __restrict__ out) {
  const int b = blockIdx.y;
  const int* pr = params + (size_t)b * 6;
  const int f0 = pr[0], f1 = pr[1], t0 = pr[2], t1 = pr[3], shift = pr[4], has_noise = pr[5];
  const size_t n = (size_t)T * F;
  // a block walks whole frames t (rows of F bins), threads over the bins: no index division, coalesced rows in and out
  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    const bool tmask = t >= t0 && t < t1;
    int tt = (t + shift) % T;                                                       // torch.roll(fbank, shift, 0)
    if (tt < 0) tt += T;
    const size_t src = (size_t)b * n + (size_t)t * F, dst = (size_t)b * n + (size_t)tt * F;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      float v = x[src + f];
      if (tmask || (f >= f0 && f < f1)) v = 0.f;                                    // masked_fill(mask, 0.)
      if (!skip_norm) v = __fdiv_rn(__fsub_rn(v, mean), std);
      if (has_noise && noise != nullptr)                                            // fbank + rand * amp / 10
        v = __fadd_rn(v, __fdiv_rn(__fmul_rn(noise[src + f], amp[b]), 10.f));
      out[dst + f] = v;
    }
  }
}

extern "C" int mla_spec_to_batch(const float* fbank, const int* params, const float* amp, const float* noise, float mean,
                                 float std, int skip_norm, int B, int T, int F, float* out, void* stream) {
  if (!fbank || !params || !amp || !out || B < 1 || T < 1 || F < 1 || fbank == out) return MLA_E_BADARG;
  const unsigned gx = (unsigned)std::min(T, 256);
  const unsigned threads = F >= 256 ? 256 : (F > 64 ? 128 : 64);
  spec_augment_kernel<<<dim3(gx, B), threads, 0, static_cast<cudaStream_t>(stream)>>>(fbank, params, amp, noise, mean, std, skip_norm,
                                                                                   T, F, out);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}
