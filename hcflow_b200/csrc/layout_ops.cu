// Layout plumbing: NCHW <-> NHWC at the module boundary (with the dequantisation noise /
// clamp / 8-bit rounding of the arch wrappers folded in), squeeze2d / unsqueeze2d and the
// Haar analysis / synthesis pair.  All memory-bound, one read + one write per element.
#include <cuda_fp16.h>

#include "common.cuh"

namespace hcf {

constexpr int LT = 256;

// thread = (b, y, x); loops channels.  NCHW side is coalesced per channel plane.
__global__ void __launch_bounds__(LT) nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                          int B, int C, int HW, int ld,
                                                          const float* __restrict__ noise, float noise_scale) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= B * HW) return;
  const int b = pix / HW, r = pix % HW;
  float* d = dst + (size_t)pix * ld;
  for (int c = 0; c < C; ++c) {
    const size_t i = ((size_t)b * C + c) * HW + r;
    float v = src[i];
    if (noise) v = v + noise[i] * noise_scale;
    d[c] = v;
  }
}

__device__ __forceinline__ float post_op(float v, int post) {
  if (post >= 1) v = fminf(fmaxf(v, 0.f), 1.f);
  if (post == 2) v = rintf(v * 255.f) / 255.f;  // torch.round = round-half-even = rintf
  return v;
}

__global__ void __launch_bounds__(LT) nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                          int B, int C, int HW, int ld, int post) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= B * HW) return;
  const int b = pix / HW, r = pix % HW;
  const float* s = src + (size_t)pix * ld;
  for (int c = 0; c < C; ++c) dst[((size_t)b * C + c) * HW + r] = post_op(s[c], post);
}

// low-res thread (b, y, x): 4C outputs  dst[c*4 + i*2 + j] = src[(2y+i, 2x+j), c]
__global__ void __launch_bounds__(LT) squeeze_kernel(const float* __restrict__ src, float* __restrict__ dst, int B,
                                                     int C, int H, int W, int src_ld, int dst_ld, int inverse) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= B * H * W) return;
  const int x = pix % W, y = (pix / W) % H, b = pix / (W * H);
  const int W2 = 2 * W, H2 = 2 * H;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) {
      const size_t hi = ((size_t)(b * H2 + 2 * y + i) * W2 + 2 * x + j);
      for (int c = 0; c < C; ++c) {
        if (!inverse) dst[(size_t)pix * dst_ld + c * 4 + i * 2 + j] = src[hi * src_ld + c];
        else dst[hi * dst_ld + c] = src[(size_t)pix * src_ld + c * 4 + i * 2 + j];
      }
    }
}

// channel-slice copy between two NHWC views of the same spatial size (Split / cat plumbing that
// can not be a pure view because of TMA's 16-byte alignment rule)
__global__ void __launch_bounds__(LT) copy_view_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                       int npix, int C, int src_ld, int dst_ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * C) return;
  const int pix = i / C, c = i % C;
  dst[(size_t)pix * dst_ld + c] = src[(size_t)pix * src_ld + c];
}

// Haar: band k of channel c at low-res channel k*C + c.
//   k0 = (a+b+c+d)/4, k1 = (a-b+c-d)/4, k2 = (a+b-c-d)/4, k3 = (a-b-c+d)/4
//   with a=(2y,2x) b=(2y,2x+1) c=(2y+1,2x) d=(2y+1,2x+1)            (Basic.py:455-466)
__global__ void __launch_bounds__(LT) haar_kernel(const float* __restrict__ src, float* __restrict__ dst, int B,
                                                  int C, int H, int W, int src_ld, int dst_ld, int inverse) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= B * H * W) return;
  const int x = pix % W, y = (pix / W) % H, b = pix / (W * H);
  const int W2 = 2 * W, H2 = 2 * H;
  const size_t p00 = ((size_t)(b * H2 + 2 * y) * W2 + 2 * x);
  const size_t p01 = p00 + 1, p10 = p00 + W2, p11 = p10 + 1;
  for (int c = 0; c < C; ++c) {
    if (!inverse) {
      const float a = src[p00 * src_ld + c], bb = src[p01 * src_ld + c];
      const float cc = src[p10 * src_ld + c], d = src[p11 * src_ld + c];
      float* o = dst + (size_t)pix * dst_ld;
      o[0 * C + c] = (a + bb + cc + d) * 0.25f;
      o[1 * C + c] = (a - bb + cc - d) * 0.25f;
      o[2 * C + c] = (a + bb - cc - d) * 0.25f;
      o[3 * C + c] = (a - bb - cc + d) * 0.25f;
    } else {
      const float* s = src + (size_t)pix * src_ld;
      const float k0 = s[0 * C + c], k1 = s[1 * C + c], k2 = s[2 * C + c], k3 = s[3 * C + c];
      dst[p00 * dst_ld + c] = k0 + k1 + k2 + k3;
      dst[p01 * dst_ld + c] = k0 - k1 + k2 - k3;
      dst[p10 * dst_ld + c] = k0 + k1 - k2 - k3;
      dst[p11 * dst_ld + c] = k0 - k1 - k2 + k3;
    }
  }
}

static int check_layout(const hcf_layout_args* a) {
  HCF_REQUIRE(a && a->src && a->dst, "layout: null args");
  HCF_REQUIRE(a->B > 0 && a->C > 0 && a->H > 0 && a->W > 0 && a->ld >= a->C, "layout: shape");
  return 0;
}

static int check_squeeze(const hcf_squeeze_args* a, bool inverse) {
  HCF_REQUIRE(a && a->src && a->dst, "squeeze: null args");
  HCF_REQUIRE(a->B > 0 && a->C > 0 && a->H > 0 && a->W > 0, "squeeze: shape");
  const int lo = 4 * a->C, hi = a->C;
  HCF_REQUIRE(a->src_ld >= (inverse ? lo : hi) && a->dst_ld >= (inverse ? hi : lo), "squeeze: ld");
  return 0;
}

}  // namespace hcf

extern "C" int hcf_nchw_to_nhwc(const hcf_layout_args* a, void* stream) {
  using namespace hcf;
  int rc = check_layout(a);
  if (rc) return rc;
  const int n = a->B * a->H * a->W;
  nchw_to_nhwc_kernel<<<ceil_div(n, LT), LT, 0, (cudaStream_t)stream>>>(a->src, a->dst, a->B, a->C, a->H * a->W,
                                                                        a->ld, a->noise, a->noise_scale);
  return finish_launch("hcf_nchw_to_nhwc");
}

extern "C" int hcf_nhwc_to_nchw(const hcf_layout_args* a, void* stream) {
  using namespace hcf;
  int rc = check_layout(a);
  if (rc) return rc;
  HCF_REQUIRE(a->post >= 0 && a->post <= 2, "nhwc_to_nchw: post %d", a->post);
  const int n = a->B * a->H * a->W;
  nhwc_to_nchw_kernel<<<ceil_div(n, LT), LT, 0, (cudaStream_t)stream>>>(a->src, a->dst, a->B, a->C, a->H * a->W,
                                                                        a->ld, a->post);
  return finish_launch("hcf_nhwc_to_nchw");
}

static int squeeze_like(const hcf_squeeze_args* a, void* stream, int inverse, int haar, const char* what) {
  using namespace hcf;
  int rc = check_squeeze(a, inverse != 0);
  if (rc) return rc;
  const int n = a->B * a->H * a->W;
  if (haar)
    haar_kernel<<<ceil_div(n, LT), LT, 0, (cudaStream_t)stream>>>(a->src, a->dst, a->B, a->C, a->H, a->W,
                                                                  a->src_ld, a->dst_ld, inverse);
  else
    squeeze_kernel<<<ceil_div(n, LT), LT, 0, (cudaStream_t)stream>>>(a->src, a->dst, a->B, a->C, a->H, a->W,
                                                                     a->src_ld, a->dst_ld, inverse);
  return finish_launch(what);
}

extern "C" int hcf_squeeze2d(const hcf_squeeze_args* a, void* s) { return squeeze_like(a, s, 0, 0, "hcf_squeeze2d"); }
extern "C" int hcf_unsqueeze2d(const hcf_squeeze_args* a, void* s) { return squeeze_like(a, s, 1, 0, "hcf_unsqueeze2d"); }
extern "C" int hcf_haar_forward(const hcf_squeeze_args* a, void* s) { return squeeze_like(a, s, 0, 1, "hcf_haar_forward"); }
extern "C" int hcf_haar_inverse(const hcf_squeeze_args* a, void* s) { return squeeze_like(a, s, 1, 1, "hcf_haar_inverse"); }

// fp32 view -> fp16 hi / lo planes of the same geometry (operand format of the fp16 conv chains:
// a = hi + lo / 2048); used for chain inputs that were not produced by a chain epilogue
__global__ void __launch_bounds__(hcf::LT) split16_kernel(const float* __restrict__ src, __half* __restrict__ hi,
                                                          __half* __restrict__ lo, long long n, int C, int ld,
                                                          int dst_ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long pix = i / C;
  const int c = (int)(i - pix * C);
  const float v = src[pix * ld + c];
  const __half h = __float2half_rn(v);
  hi[pix * dst_ld + c] = h;
  if (lo) lo[pix * dst_ld + c] = __float2half_rn((v - __half2float(h)) * 2048.0f);
}

extern "C" int hcf_split16(const float* src, int32_t ld, int32_t C, int64_t npix, void* hi, void* lo, int32_t dst_ld,
                           void* stream) {
  using namespace hcf;
  HCF_REQUIRE(src && hi && ld >= C && dst_ld >= C && C > 0 && npix > 0, "split16: bad args");
  const long long n = (long long)npix * C;
  split16_kernel<<<(unsigned)((n + LT - 1) / LT), LT, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<__half*>(hi),
                                                                                 reinterpret_cast<__half*>(lo), n, C, ld, dst_ld);
  return finish_launch("hcf_split16");
}

// ---- uint8 image edges: images enter and leave the engine as the 8-bit HWC arrays cv2 produces.
// in:  codes/data/util.py:72-86 (astype(float32) / 255.), GTLQ_dataset.py:109-115 (BGR -> RGB, HWC -> CHW)
// out: codes/utils/util.py:790-816 tensor2img (clamp [0,1], RGB -> BGR, (x * 255.0).round() -> uint8, CHW -> HWC)
__global__ void __launch_bounds__(hcf::LT) u8_to_nhwc_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                             long long npix, int ld, int swap_rb) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const uint8_t* sp = src + pix * 3;
  float* d = dst + pix * ld;
  const float c0 = (float)sp[0] / 255.f, c1 = (float)sp[1] / 255.f, c2 = (float)sp[2] / 255.f;
  d[0] = swap_rb ? c2 : c0;
  d[1] = c1;
  d[2] = swap_rb ? c0 : c2;
}

__global__ void __launch_bounds__(hcf::LT) nhwc_to_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst,
                                                             long long npix, int ld, int swap_rb) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const float* sp = src + pix * ld;
  uint8_t* d = dst + pix * 3;
  float v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = rintf(fminf(fmaxf(sp[c], 0.f), 1.f) * 255.f);   // numpy round = half to even
  d[0] = (uint8_t)(swap_rb ? v[2] : v[0]);
  d[1] = (uint8_t)v[1];
  d[2] = (uint8_t)(swap_rb ? v[0] : v[2]);
}

extern "C" int hcf_u8_hwc_to_nhwc(const uint8_t* src, float* dst, int32_t ld, int64_t npix, int32_t swap_rb, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(src && dst && ld >= 3 && npix > 0, "u8_hwc_to_nhwc: bad args");
  u8_to_nhwc_kernel<<<(unsigned)((npix + LT - 1) / LT), LT, 0, (cudaStream_t)stream>>>(src, dst, npix, ld, swap_rb);
  return finish_launch("hcf_u8_hwc_to_nhwc");
}

extern "C" int hcf_nhwc_to_u8_hwc(const float* src, int32_t ld, uint8_t* dst, int64_t npix, int32_t swap_rb, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(src && dst && ld >= 3 && npix > 0, "nhwc_to_u8_hwc: bad args");
  nhwc_to_u8_kernel<<<(unsigned)((npix + LT - 1) / LT), LT, 0, (cudaStream_t)stream>>>(src, dst, npix, ld, swap_rb);
  return finish_launch("hcf_nhwc_to_u8_hwc");
}

// nearest-neighbour upsampling by 2^shift between NHWC views (F.interpolate(..., mode='nearest'),
// FlowNet_SR_x4.py:98,117): materialises a conv input segment so that the conv can run on the tensor cores
__global__ void __launch_bounds__(hcf::LT) upsample_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                                           int B, int C4, int H, int W, int src_ld4, int dst_ld4,
                                                           int shift) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * H * W * C4) return;
  const int c = (int)(i % C4);
  const long long pix = i / C4;
  const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
  const long long sp = ((long long)b * (H >> shift) + (y >> shift)) * (W >> shift) + (x >> shift);
  dst[pix * dst_ld4 + c] = __ldg(src + sp * src_ld4 + c);
}

// a->H, a->W: the DESTINATION (high-res) size; a->C channels, multiples of 4, 16-byte aligned views
extern "C" int hcf_upsample_nearest(const hcf_squeeze_args* a, int32_t shift, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(a && a->src && a->dst && shift >= 1 && shift <= 3, "upsample: bad args");
  HCF_REQUIRE(a->B > 0 && a->C > 0 && a->C % 4 == 0 && a->src_ld % 4 == 0 && a->dst_ld % 4 == 0 && a->src_ld >= a->C &&
                  a->dst_ld >= a->C && a->H % (1 << shift) == 0 && a->W % (1 << shift) == 0 && aligned16(a->src) &&
                  aligned16(a->dst), "upsample: shape / alignment");
  const long long n = (long long)a->B * a->H * a->W * (a->C / 4);
  upsample_kernel<<<(unsigned)((n + LT - 1) / LT), LT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(a->src), reinterpret_cast<float4*>(a->dst), a->B, a->C / 4, a->H, a->W,
      a->src_ld / 4, a->dst_ld / 4, shift);
  return finish_launch("hcf_upsample_nearest");
}

extern "C" int hcf_copy_view(const hcf_squeeze_args* a, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(a && a->src && a->dst, "copy_view: null args");
  HCF_REQUIRE(a->B > 0 && a->C > 0 && a->H > 0 && a->W > 0 && a->src_ld >= a->C && a->dst_ld >= a->C, "copy_view: shape");
  const int n = a->B * a->H * a->W * a->C;
  copy_view_kernel<<<ceil_div(n, LT), LT, 0, (cudaStream_t)stream>>>(a->src, a->dst, a->B * a->H * a->W, a->C,
                                                                     a->src_ld, a->dst_ld);
  return finish_launch("hcf_copy_view");
}

// ================================================================================================ SURVEY 8f-4
// Tiled inference (codes/data/util.py:489-514 test_patchwise): E += patch, W += 1 over the patch's window; E /= W.
// Evaluation metrics on the device (codes/utils/util.py:902-982 calculate_psnr / ssim / calculate_psnr_ssim,
// codes/data/util.py:209-230 bgr2ycbcr): fp64 like the reference's numpy code.
namespace hcf {

// patches [n, C, ph, pw] (NCHW, the module's output layout) accumulated into E [C, H, W] at (y0[i], x0[i]); cnt [H, W]
__global__ void tile_accumulate_kernel(const float* __restrict__ patches, int n, int C, int ph, int pw, const int* __restrict__ y0,
                                       const int* __restrict__ x0, float* __restrict__ E, float* __restrict__ cnt, int H, int W) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * C * ph * pw;
  if (i >= total) return;
  const int x = (int)(i % pw);
  long long r = i / pw;
  const int y = (int)(r % ph); r /= ph;
  const int c = (int)(r % C);
  const int k = (int)(r / C);
  const int gy = y0[k] + y, gx = x0[k] + x;
  if (gy < 0 || gy >= H || gx < 0 || gx >= W) return;
  atomicAdd(E + ((long long)c * H + gy) * W + gx, patches[i]);
  if (c == 0) atomicAdd(cnt + (long long)gy * W + gx, 1.0f);
}
__global__ void tile_normalize_kernel(float* __restrict__ E, const float* __restrict__ cnt, int C, long long hw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)C * hw) return;
  E[i] = E[i] / cnt[i % hw];
}

// pixel value on the reference's [0, 255] scale: images are HWC, uint8 or float in [0, 1]
__device__ __forceinline__ double px255(const void* img, int f32, long long idx) {
  return f32 ? (double)reinterpret_cast<const float*>(img)[idx] * 255.0 : (double)reinterpret_cast<const unsigned char*>(img)[idx];
}
// channel `ch` of pixel (y, x); ch == -1: the Y channel of a BGR image, (24.966 B + 128.553 G + 65.481 R) / 255 + 16
__device__ __forceinline__ double chan255(const void* img, int f32, int W, int C, int y, int x, int ch) {
  const long long base = ((long long)y * W + x) * C;
  if (ch >= 0) return px255(img, f32, base + ch);
  return (24.966 * px255(img, f32, base) + 128.553 * px255(img, f32, base + 1) + 65.481 * px255(img, f32, base + 2)) / 255.0 + 16.0;
}
__device__ __forceinline__ void block_add(double v, double* out) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ double part[8];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) part[w] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    atomicAdd(out, t);
  }
  __syncthreads();
}
// sum of squared differences of channel ch over the cropped image
__global__ void __launch_bounds__(256) sqdiff_kernel(const void* a, const void* b, int f32, int H, int W, int C, int crop, int ch,
                                                     double* out) {
  const int h = H - 2 * crop, w = W - 2 * crop;
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < (long long)h * w; i += (long long)gridDim.x * 256) {
    const int y = crop + (int)(i / w), x = crop + (int)(i % w);
    const double d = chan255(a, f32, W, C, y, x, ch) - chan255(b, f32, W, C, y, x, ch);
    s += d * d;
  }
  block_add(s, out);
}
// sum of the SSIM map (11x11 Gaussian window, sigma 1.5, 'valid' region) of channel ch over the cropped image
__global__ void __launch_bounds__(256) ssim_kernel(const void* a, const void* b, int f32, int H, int W, int C, int crop, int ch,
                                                   const double* __restrict__ win, double* out) {
  const int h = H - 2 * crop - 10, w = W - 2 * crop - 10;
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < (long long)h * w; i += (long long)gridDim.x * 256) {
    const int y = crop + (int)(i / w), x = crop + (int)(i % w);
    double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
    for (int dy = 0; dy < 11; ++dy)
      for (int dx = 0; dx < 11; ++dx) {
        const double wgt = win[dy * 11 + dx];
        const double p = chan255(a, f32, W, C, y + dy, x + dx, ch), q = chan255(b, f32, W, C, y + dy, x + dx, ch);
        m1 += wgt * p; m2 += wgt * q;
        s11 += wgt * p * p; s22 += wgt * q * q; s12 += wgt * p * q;
      }
    const double C1 = (0.01 * 255) * (0.01 * 255), C2 = (0.03 * 255) * (0.03 * 255);
    const double v1 = s11 - m1 * m1, v2 = s22 - m2 * m2, cov = s12 - m1 * m2;
    s += ((2 * m1 * m2 + C1) * (2 * cov + C2)) / ((m1 * m1 + m2 * m2 + C1) * (v1 + v2 + C2));
  }
  block_add(s, out);
}

}  // namespace hcf

extern "C" int hcf_tile_accumulate(const float* patches, int32_t n, int32_t C, int32_t ph, int32_t pw, const int32_t* y0,
                                   const int32_t* x0, float* E, float* cnt, int32_t H, int32_t W, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(patches && y0 && x0 && E && cnt && n > 0 && C > 0 && ph > 0 && pw > 0 && H > 0 && W > 0, "tile_accumulate: bad args");
  const long long total = (long long)n * C * ph * pw;
  tile_accumulate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(patches, n, C, ph, pw, y0, x0, E, cnt, H, W);
  return finish_launch("hcf_tile_accumulate");
}

extern "C" int hcf_tile_normalize(float* E, const float* cnt, int32_t C, int32_t H, int32_t W, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(E && cnt && C > 0 && H > 0 && W > 0, "tile_normalize: bad args");
  const long long hw = (long long)H * W;
  tile_normalize_kernel<<<(unsigned)((C * hw + 255) / 256), 256, 0, (cudaStream_t)stream>>>(E, cnt, C, hw);
  return finish_launch("hcf_tile_normalize");
}

// out[0..C-1]: sum of squared differences per channel, out[C..2C-1]: SSIM-map sums per channel, out[2C]: squared
// differences of the Y channel, out[2C+1]: SSIM-map sum of Y (C == 3 only; BGR order), all on the [0, 255] scale.
// a, b: HWC images, uint8 (is_f32 = 0) or float in [0, 1] (is_f32 = 1); win: 121 doubles (the 11x11 Gaussian window).
extern "C" int hcf_image_metrics(const void* a, const void* b, int32_t is_f32, int32_t H, int32_t W, int32_t C, int32_t crop,
                                 const double* win, double* out, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(a && b && win && out && H - 2 * crop > 10 && W - 2 * crop > 10 && C >= 1 && C <= 4 && crop >= 0,
              "image_metrics: bad args (the cropped image must be larger than the 11x11 SSIM window)");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double) * (2 * C + 2), st);
  if (e != cudaSuccess) { set_error("image_metrics: %s", cudaGetErrorString(e)); return (int)e; }
  const int h = H - 2 * crop, w = W - 2 * crop;
  int g1 = (int)(((long long)h * w + 255) / 256), g2 = (int)(((long long)(h - 10) * (w - 10) + 255) / 256);
  g1 = g1 > 1024 ? 1024 : g1; g2 = g2 > 4096 ? 4096 : g2;
  for (int c = 0; c < C; ++c) {
    sqdiff_kernel<<<g1, 256, 0, st>>>(a, b, is_f32, H, W, C, crop, c, out + c);
    ssim_kernel<<<g2, 256, 0, st>>>(a, b, is_f32, H, W, C, crop, c, win, out + C + c);
  }
  if (C == 3) {
    sqdiff_kernel<<<g1, 256, 0, st>>>(a, b, is_f32, H, W, C, crop, -1, out + 2 * C);
    ssim_kernel<<<g2, 256, 0, st>>>(a, b, is_f32, H, W, C, crop, -1, win, out + 2 * C + 1);
  }
  return finish_launch("hcf_image_metrics");
}
