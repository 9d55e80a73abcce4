// Training-path kernels: the backward halves of the hot-path ops (HCFlow_SR_model.py:184-218
// optimize_parameters) plus the few elementwise forwards the fused inference engine folds into conv epilogues.
// They are surfaced to Python as torch.autograd.Function extensions (hcflow_b200/autograd.py).  Everything is exact
// fp32 on CUDA cores over NHWC tensors: the training path is about gradients that match the reference's autograd, the
// tensor-core engine stays the inference path.
//   conv:      dx = conv(dy, flipped / transposed w) reuses hcf_conv_fp32; dw = hcf_conv_wgrad; db = hcf_channel_sum
//   ActNorm / bias+activation:  y = act((x + b) * s)                      hcf_affine_act_fwd / _bwd
//   coupling:  z2' = (z2 + shift) * exp(ls), ls = 0.318 atan(2 scale)      hcf_coupling_fwd / _bwd  (+ inverse form)
//   Gaussian:  logp = sum -0.5 (2 logs + (x - mu)^2 / e^(2 logs) + ln 2pi)  hcf_gauss_logp_fwd / _bwd
//   y = alpha a + beta b, nearest-upsample adjoint, 8-bit quantisation (straight-through backward, Basic.py:186-198)
#include "common.cuh"

namespace hcf {

// ------------------------------------------------------------------------------------------------ conv weight gradient
// dw[co][ci][ky][kx] = sum_{b,y,x} dy[b,y,x,co] * x[b, y+ky-p, x+kx-p, ci]      (zero padding p = ks / 2)
// grid (ks*ks, ceil(Cin/16), ceil(Cout/16)), block 16x16: thread (ci, co) of the tile; pixels stream through shared
// memory in chunks of 64, the grid's z-split over pixel ranges adds partial sums with atomics.
constexpr int WG_PIX = 64;
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ x, int x_ld, const float* __restrict__ dy,
                                                         int dy_ld, int B, int H, int W, int Cin, int Cout, int ks,
                                                         float* __restrict__ dw, int pix_splits) {
  __shared__ float sx[WG_PIX][17];
  __shared__ float sd[WG_PIX][17];
  const int tap = blockIdx.x % (ks * ks), split = blockIdx.x / (ks * ks);
  const int ky = tap / ks, kx = tap % ks, p = ks / 2;
  const int ci0 = blockIdx.y * 16, co0 = blockIdx.z * 16;
  const int tci = threadIdx.x & 15, tco = threadIdx.x >> 4;
  const long long npix = (long long)B * H * W;
  const long long per = (npix + pix_splits - 1) / pix_splits;
  const long long lo = split * per, hi = lo + per < npix ? lo + per : npix;
  float acc = 0.f;
  for (long long base = lo; base < hi; base += WG_PIX) {
    for (int i = threadIdx.x; i < WG_PIX * 16; i += 256) {
      const int pp = i >> 4, c = i & 15;
      const long long pix = base + pp;
      float vx = 0.f, vd = 0.f;
      if (pix < hi) {
        const int b = (int)(pix / ((long long)H * W));
        const int r = (int)(pix % ((long long)H * W));
        const int yy = r / W, xx = r % W;
        const int sy = yy + ky - p, sxx = xx + kx - p;
        if (sy >= 0 && sy < H && sxx >= 0 && sxx < W && ci0 + c < Cin)
          vx = x[((long long)(b * H + sy) * W + sxx) * x_ld + ci0 + c];
        if (co0 + c < Cout) vd = dy[pix * dy_ld + co0 + c];
      }
      sx[pp][c] = vx;
      sd[pp][c] = vd;
    }
    __syncthreads();
#pragma unroll 8
    for (int pp = 0; pp < WG_PIX; ++pp) acc = fmaf(sd[pp][tco], sx[pp][tci], acc);
    __syncthreads();
  }
  if (ci0 + tci < Cin && co0 + tco < Cout)
    atomicAdd(dw + (((long long)(co0 + tco) * Cin + ci0 + tci) * ks + ky) * ks + kx, acc);
}

// out[c] += sum over pixels of y[pix, c]
__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ y, int ld, int C, long long npix,
                                                          float* __restrict__ out) {
  const int c = blockIdx.x;
  float s = 0.f;
  for (long long i = (long long)blockIdx.y * 256 + threadIdx.x; i < npix; i += (long long)gridDim.y * 256) s += y[i * ld + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(out + c, t);
  }
}

// ------------------------------------------------------------------------------------------------ y = act((x + b) * s)
__device__ __forceinline__ float act_apply(float v, int act) {
  return act == HCF_ACT_RELU ? fmaxf(v, 0.f) : (act == HCF_ACT_LRELU ? (v > 0.f ? v : 0.2f * v) : v);
}
__device__ __forceinline__ float act_grad(float pre, int act) {
  return act == HCF_ACT_RELU ? (pre > 0.f ? 1.f : 0.f) : (act == HCF_ACT_LRELU ? (pre > 0.f ? 1.f : 0.2f) : 1.f);
}
__global__ void affine_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ b, const float* __restrict__ s,
                                      int act, float* __restrict__ y, long long n, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  const float pre = (x[i] + (b ? b[c] : 0.f)) * (s ? s[c] : 1.f);
  y[i] = act_apply(pre, act);
}
// dx = dy * act'(pre) * s;  db[c] += sum dx;  ds[c] += sum dy * act'(pre) * (x + b)
__global__ void __launch_bounds__(256) affine_act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                             const float* __restrict__ b, const float* __restrict__ s, int act,
                                                             float* __restrict__ dx, float* __restrict__ db,
                                                             float* __restrict__ ds, long long npix, int C) {
  // block = 256 threads = 8 pixel lanes x 32 channel lanes; grid.x over channel groups of 32, grid.y strides pixels
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int pl = threadIdx.x >> 5;
  float sb = 0.f, ss = 0.f;
  if (c < C) {
    const float bb = b ? b[c] : 0.f, sc = s ? s[c] : 1.f;
    for (long long pix = (long long)blockIdx.y * 8 + pl; pix < npix; pix += (long long)gridDim.y * 8) {
      const long long i = pix * C + c;
      const float xb = x[i] + bb;
      const float g = dy[i] * act_grad(xb * sc, act);
      dx[i] = g * sc;
      sb += g * sc;
      ss += g * xb;
    }
  }
  __shared__ float pb[8][32], ps[8][32];
  pb[pl][threadIdx.x & 31] = sb;
  ps[pl][threadIdx.x & 31] = ss;
  __syncthreads();
  if (pl == 0 && c < C) {
    float tb = 0.f, ts = 0.f;
    for (int i = 0; i < 8; ++i) { tb += pb[i][threadIdx.x]; ts += ps[i][threadIdx.x]; }
    if (db) atomicAdd(db + c, tb);
    if (ds) atomicAdd(ds + c, ts);
  }
}

// ------------------------------------------------------------------------------------------------ affine coupling
// forward (AffineCouplings.py:52-61):  out = (z2 + shift) * exp(ls),  lsum[img] += sum ls     (h = [shift0, scale0, shift1, ...])
// inverse (AffineCouplings.py:78-85):  out = z2 * exp(-ls) - shift
__global__ void __launch_bounds__(128) coupling_fwd_kernel(const float* __restrict__ z2, const float* __restrict__ h, int nc,
                                                           int inverse, float* __restrict__ out, double* __restrict__ lsum,
                                                           long long npix, int pix_per_img) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = pix < npix;
  float ls_sum = 0.f;
  if (active) {
    for (int j = 0; j < nc; ++j) {
      const float shift = h[pix * 2 * nc + 2 * j], scale = h[pix * 2 * nc + 2 * j + 1];
      const float ls = coupling_logscale(scale);
      const float v = z2[pix * nc + j];
      out[pix * nc + j] = inverse ? v * expf(-ls) - shift : (v + shift) * expf(ls);
      ls_sum += ls;
    }
  }
  if (lsum && !inverse) {
    // one atomic per lane is fine here (training path); images may straddle warps
    if (active) atomicAdd(lsum + pix / pix_per_img, (double)ls_sum);
  }
}
// given d(out) and (forward only) d(lsum[img]):  dz2, dh
__global__ void __launch_bounds__(128) coupling_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ dlsum,
                                                           const float* __restrict__ z2, const float* __restrict__ h, int nc,
                                                           int inverse, float* __restrict__ dz2, float* __restrict__ dh,
                                                           long long npix, int pix_per_img) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const float gl = dlsum ? dlsum[pix / pix_per_img] : 0.f;
  for (int j = 0; j < nc; ++j) {
    const float shift = h[pix * 2 * nc + 2 * j], scale = h[pix * 2 * nc + 2 * j + 1];
    const float ls = coupling_logscale(scale);
    const float dls_dscale = 0.318f * 2.0f / (1.0f + 4.0f * scale * scale);
    const float v = z2[pix * nc + j], g = dout[pix * nc + j];
    float dshift, dls;
    if (!inverse) {
      const float e = expf(ls);
      dz2[pix * nc + j] = g * e;
      dshift = g * e;
      dls = g * (v + shift) * e + gl;
    } else {
      const float e = expf(-ls);
      dz2[pix * nc + j] = g * e;
      dshift = -g;
      dls = -g * v * e;
    }
    dh[pix * 2 * nc + 2 * j] = dshift;
    dh[pix * 2 * nc + 2 * j + 1] = dls * dls_dscale;
  }
}

// ------------------------------------------------------------------------------------------------ diagonal Gaussian
// out[img] += sum over the image's elements of -0.5 (2 logs + (x - mu)^2 / exp(2 logs) + ln 2 pi)      (Basic.py:79-93)
__global__ void __launch_bounds__(256) gauss_logp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mu,
                                                             const float* __restrict__ logs, float logs_const,
                                                             long long per_img, double* __restrict__ out) {
  const int b = blockIdx.y;
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < per_img; i += (long long)gridDim.x * 256) {
    const long long k = (long long)b * per_img + i;
    const float lg = logs ? logs[k] : logs_const;
    const float d = x[k] - mu[k];
    s += (double)(-0.5f * (lg * 2.f + (d * d) / expf(lg * 2.f) + 1.8378770664093453f));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(out + b, t);
  }
}
__global__ void gauss_logp_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ mu,
                                      const float* __restrict__ logs, float logs_const, long long per_img, long long n,
                                      float* __restrict__ dx, float* __restrict__ dmu, float* __restrict__ dlogs) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const float gb = g[k / per_img];
  const float lg = logs ? logs[k] : logs_const;
  const float d = x[k] - mu[k], iv = expf(-2.f * lg);
  const float t = -d * iv * gb;
  if (dx) dx[k] = t;
  if (dmu) dmu[k] = -t;
  if (dlogs) dlogs[k] = (d * d * iv - 1.f) * gb;
}

// prior sample (Basic.py:96-100, ConditionalFlow.py:62-64): out = mean + exp(logs) * eps;  dmean = g, dlogs = g * exp(logs) * eps
__global__ void gauss_sample_kernel(const float* __restrict__ mean, const float* __restrict__ logs, const float* __restrict__ eps,
                                    const float* __restrict__ g, float* __restrict__ out, float* __restrict__ dlogs, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float e = expf(logs[i]) * eps[i];
  if (g) dlogs[i] = g[i] * e;          // backward (dmean = g needs no kernel)
  else out[i] = mean[i] + e;           // forward
}

// ------------------------------------------------------------------------------------------------ small elementwise ops
__global__ void axpby_kernel(const float* __restrict__ a, float alpha, const float* __restrict__ b, float beta,
                             float* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = alpha * a[i] + (b ? beta * b[i] : 0.f);
}
__global__ void quantize_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = rintf(fminf(fmaxf(x[i], 0.f), 1.f) * 255.f) / 255.f;   // torch.round = round-half-even = rintf
}
// adjoint of nearest up-sampling by 2^shift: dst[b,y,x,c] = sum over the (2^shift)^2 block of src
__global__ void downsample_sum_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int H, int W, int C, int shift) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)B * H * W * C;
  if (i >= n) return;
  const int c = (int)(i % C);
  long long r = i / C;
  const int x = (int)(r % W); r /= W;
  const int y = (int)(r % H);
  const int b = (int)(r / H);
  const int f = 1 << shift, HH = H << shift, WW = W << shift;
  float s = 0.f;
  for (int dy = 0; dy < f; ++dy)
    for (int dx = 0; dx < f; ++dx) s += src[(((long long)b * HH + (y * f + dy)) * WW + (x * f + dx)) * C + c];
  dst[i] = s;
}

}  // namespace hcf

// ================================================================================================ C ABI
#define HCF_GRID1(n, bs) (unsigned)(((n) + (bs)-1) / (bs))

extern "C" int hcf_conv_wgrad(const float* x, int32_t x_ld, const float* dy, int32_t dy_ld, int32_t B, int32_t H, int32_t W,
                              int32_t Cin, int32_t Cout, int32_t ks, float* dw, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(x && dy && dw && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (ks == 1 || ks == 3) && x_ld >= Cin && dy_ld >= Cout,
              "conv_wgrad: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * Cin * ks * ks, st);
  if (e != cudaSuccess) { set_error("conv_wgrad: %s", cudaGetErrorString(e)); return (int)e; }
  const long long npix = (long long)B * H * W;
  int splits = (int)((npix + 4095) / 4096);
  if (splits > 64) splits = 64;
  dim3 grid((unsigned)(ks * ks * splits), (unsigned)ceil_div(Cin, 16), (unsigned)ceil_div(Cout, 16));
  conv_wgrad_kernel<<<grid, 256, 0, st>>>(x, x_ld, dy, dy_ld, B, H, W, Cin, Cout, ks, dw, splits);
  return finish_launch("hcf_conv_wgrad");
}

extern "C" int hcf_channel_sum(const float* y, int32_t ld, int32_t C, int64_t npix, float* out, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(y && out && C > 0 && npix > 0 && ld >= C, "channel_sum: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * C, st);
  if (e != cudaSuccess) { set_error("channel_sum: %s", cudaGetErrorString(e)); return (int)e; }
  int gy = (int)((npix + 2047) / 2048);
  if (gy > 64) gy = 64;
  channel_sum_kernel<<<dim3((unsigned)C, (unsigned)gy), 256, 0, st>>>(y, ld, C, npix, out);
  return finish_launch("hcf_channel_sum");
}

extern "C" int hcf_affine_act_fwd(const float* x, const float* bias, const float* scale, int32_t act, float* y, int64_t npix,
                                  int32_t C, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(x && y && npix > 0 && C > 0, "affine_act_fwd: bad args");
  const long long n = npix * C;
  affine_act_fwd_kernel<<<HCF_GRID1(n, 256), 256, 0, (cudaStream_t)stream>>>(x, bias, scale, act, y, n, C);
  return finish_launch("hcf_affine_act_fwd");
}

extern "C" int hcf_affine_act_bwd(const float* dy, const float* x, const float* bias, const float* scale, int32_t act, float* dx,
                                  float* dbias, float* dscale, int64_t npix, int32_t C, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(dy && x && dx && npix > 0 && C > 0, "affine_act_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * C, st);
  if (dscale) cudaMemsetAsync(dscale, 0, sizeof(float) * C, st);
  int gy = (int)((npix + 511) / 512);
  if (gy > 256) gy = 256;
  affine_act_bwd_kernel<<<dim3((unsigned)ceil_div(C, 32), (unsigned)gy), 256, 0, st>>>(dy, x, bias, scale, act, dx, dbias, dscale,
                                                                                        npix, C);
  return finish_launch("hcf_affine_act_bwd");
}

extern "C" int hcf_coupling_fwd(const float* z2, const float* h, int32_t nc, int32_t inverse, float* out, double* lsum, int64_t npix,
                                int32_t pix_per_img, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(z2 && h && out && nc > 0 && npix > 0 && pix_per_img > 0, "coupling_fwd: bad args");
  coupling_fwd_kernel<<<HCF_GRID1(npix, 128), 128, 0, (cudaStream_t)stream>>>(z2, h, nc, inverse, out, lsum, npix, pix_per_img);
  return finish_launch("hcf_coupling_fwd");
}

extern "C" int hcf_coupling_bwd(const float* dout, const float* dlsum, const float* z2, const float* h, int32_t nc, int32_t inverse,
                                float* dz2, float* dh, int64_t npix, int32_t pix_per_img, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(dout && z2 && h && dz2 && dh && nc > 0 && npix > 0 && pix_per_img > 0, "coupling_bwd: bad args");
  coupling_bwd_kernel<<<HCF_GRID1(npix, 128), 128, 0, (cudaStream_t)stream>>>(dout, dlsum, z2, h, nc, inverse, dz2, dh, npix,
                                                                              pix_per_img);
  return finish_launch("hcf_coupling_bwd");
}

extern "C" int hcf_gauss_logp_fwd(const float* x, const float* mean, const float* logs, float logs_const, int32_t B,
                                  int64_t per_img, double* out, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(x && mean && out && B > 0 && per_img > 0, "gauss_logp_fwd: bad args");
  int gx = (int)((per_img + 2047) / 2048);
  if (gx > 64) gx = 64;
  gauss_logp_fwd_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(x, mean, logs, logs_const, per_img, out);
  return finish_launch("hcf_gauss_logp_fwd");
}

extern "C" int hcf_gauss_logp_bwd(const float* g, const float* x, const float* mean, const float* logs, float logs_const, int32_t B,
                                  int64_t per_img, float* dx, float* dmean, float* dlogs, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(g && x && mean && B > 0 && per_img > 0, "gauss_logp_bwd: bad args");
  const long long n = (long long)B * per_img;
  gauss_logp_bwd_kernel<<<HCF_GRID1(n, 256), 256, 0, (cudaStream_t)stream>>>(g, x, mean, logs, logs_const, per_img, n, dx, dmean,
                                                                             dlogs);
  return finish_launch("hcf_gauss_logp_bwd");
}

extern "C" int hcf_gauss_sample(const float* mean, const float* logs, const float* eps, const float* g, float* out, float* dlogs,
                                int64_t n, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(logs && eps && n > 0 && ((g && dlogs) || (!g && mean && out)), "gauss_sample: bad args");
  gauss_sample_kernel<<<HCF_GRID1(n, 256), 256, 0, (cudaStream_t)stream>>>(mean, logs, eps, g, out, dlogs, n);
  return finish_launch("hcf_gauss_sample");
}

extern "C" int hcf_axpby(const float* a, float alpha, const float* b, float beta, float* y, int64_t n, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(a && y && n > 0, "axpby: bad args");
  axpby_kernel<<<HCF_GRID1(n, 256), 256, 0, (cudaStream_t)stream>>>(a, alpha, b, beta, y, n);
  return finish_launch("hcf_axpby");
}

extern "C" int hcf_quantize8(const float* x, float* y, int64_t n, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(x && y && n > 0, "quantize8: bad args");
  quantize_kernel<<<HCF_GRID1(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n);
  return finish_launch("hcf_quantize8");
}

extern "C" int hcf_downsample_sum(const float* src, float* dst, int32_t B, int32_t H, int32_t W, int32_t C, int32_t shift,
                                  void* stream) {
  using namespace hcf;
  HCF_REQUIRE(src && dst && B > 0 && H > 0 && W > 0 && C > 0 && shift >= 1 && shift <= 3, "downsample_sum: bad args");
  const long long n = (long long)B * H * W * C;
  downsample_sum_kernel<<<HCF_GRID1(n, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, B, H, W, C, shift);
  return finish_launch("hcf_downsample_sum");
}
