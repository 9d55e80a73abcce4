// Per-pixel FlowStep arithmetic: coupling affine, invertible 1x1 mix, ActNorm, log-det,
// and the conditional Gaussian prior.  One thread = one pixel; the channel vector lives in
// registers (C is a template parameter), the C x C matrix and the ActNorm vectors in
// shared memory (warp-broadcast reads).  These kernels are HBM/L2-bound: every pixel is
// read once and written once per FlowStep.
#include "common.cuh"

namespace hcf {

constexpr int STEP_THREADS = 128;

struct StepParams {
  int npix, pix_per_img;
  float* z;
  int z_ld;
  const float* h;
  int h_ld, mode, n_pass;
  const float* w;
  const float* an_scale;
  const float* an_bias;
  double* logdet;
};

template <int C>
__device__ __forceinline__ void load_tables(const StepParams& p, float* s_w, float* s_sc, float* s_b) {
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_w[i] = p.w ? p.w[i] : 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_sc[i] = p.an_scale ? p.an_scale[i] : 1.f;
    s_b[i] = p.an_bias ? p.an_bias[i] : 0.f;
  }
  __syncthreads();
}

template <int C>
__device__ __forceinline__ void matvec(const float* s_w, float (&z)[C]) {
  float y[C];
#pragma unroll
  for (int i = 0; i < C; ++i) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < C; ++j) acc = fmaf(s_w[i * C + j], z[j], acc);
    y[i] = acc;
  }
#pragma unroll
  for (int i = 0; i < C; ++i) z[i] = y[i];
}

// affine^-1 -> W^-1 -> actnorm^-1
template <int C>
__global__ void __launch_bounds__(STEP_THREADS) step_inverse_kernel(const StepParams p) {
  __shared__ float s_w[C * C];
  __shared__ float s_sc[C], s_b[C];
  load_tables<C>(p, s_w, s_sc, s_b);
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= p.npix) return;
  float* zp = p.z + (size_t)pix * p.z_ld;
  float z[C];
#pragma unroll
  for (int i = 0; i < C; ++i) z[i] = zp[i];
  if (p.h) {
    const float* hp = p.h + (size_t)pix * p.h_ld;
    if (p.mode == HCF_COUPLING_AFFINE) {
#pragma unroll
      for (int i = 0; i < C; ++i) {
        if (i >= p.n_pass) {
          const int j = i - p.n_pass;
          const float shift = __ldg(hp + 2 * j), scale = __ldg(hp + 2 * j + 1);
          z[i] = z[i] * expf(-coupling_logscale(scale)) - shift;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) z[i] -= __ldg(hp + i);
    }
  }
  if (p.w) matvec<C>(s_w, z);
#pragma unroll
  for (int i = 0; i < C; ++i) zp[i] = z[i] * s_sc[i] - s_b[i];
}

// actnorm -> W
template <int C>
__global__ void __launch_bounds__(STEP_THREADS) step_forward_head_kernel(const StepParams p) {
  __shared__ float s_w[C * C];
  __shared__ float s_sc[C], s_b[C];
  load_tables<C>(p, s_w, s_sc, s_b);
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= p.npix) return;
  float* zp = p.z + (size_t)pix * p.z_ld;
  float z[C];
#pragma unroll
  for (int i = 0; i < C; ++i) z[i] = (zp[i] + s_b[i]) * s_sc[i];
  if (p.w) matvec<C>(s_w, z);
#pragma unroll
  for (int i = 0; i < C; ++i) zp[i] = z[i];
}

__device__ __forceinline__ void add_logdet(double* logdet, int img, double v, bool active) {
  // warp-aggregate when every active lane of the warp sits in one image, else one atomic per lane.
  // Must be called by all 32 lanes of the warp.
  const unsigned full = 0xffffffffu;
  const unsigned mask = __ballot_sync(full, active);
  if (mask == 0u) return;
  const int leader = __ffs(mask) - 1;
  const int img0 = __shfl_sync(full, img, leader);
  const bool same = __all_sync(full, (!active) || img == img0);
  if (same) {
    double s = active ? v : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(full, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(logdet + img0, s);
  } else if (active) {
    atomicAdd(logdet + img, v);
  }
}

// z2 = (z2 + shift) * exp(ls); logdet += sum ls      (generic in C: only touches C - n_pass channels)
__global__ void __launch_bounds__(STEP_THREADS) step_forward_coupling_kernel(const StepParams p, int C) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = pix < p.npix;
  float lsum = 0.f;
  if (active) {
    float* zp = p.z + (size_t)pix * p.z_ld;
    const float* hp = p.h + (size_t)pix * p.h_ld;
    if (p.mode == HCF_COUPLING_AFFINE) {
      for (int i = p.n_pass; i < C; ++i) {
        const int j = i - p.n_pass;
        const float shift = __ldg(hp + 2 * j), scale = __ldg(hp + 2 * j + 1);
        const float ls = coupling_logscale(scale);
        zp[i] = (zp[i] + shift) * expf(ls);
        lsum += ls;
      }
    } else {
      for (int i = 0; i < 3; ++i) zp[i] += __ldg(hp + i);
    }
  }
  if (p.logdet && p.mode == HCF_COUPLING_AFFINE)
    add_logdet(p.logdet, active ? pix / p.pix_per_img : 0, (double)lsum, active);
}

// ---------------------------------------------------------------- prior
struct PriorParams {
  int B, H, W, Cz;
  const float* h;
  int h_ld, atan_logscale;
  const float* eps;
  float* z;
  int z_ld;
  double* logdet;
  float* out_nchw;
};

enum { PRIOR_SAMPLE = 0, PRIOR_LOGP = 1, PRIOR_STD = 2 };

template <int MODE>
__global__ void __launch_bounds__(STEP_THREADS) prior_kernel(const PriorParams p) {
  const int hw = p.H * p.W;
  const int npix = p.B * hw;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = pix < npix;
  float lp = 0.f;
  if (active) {
    const int b = pix / hw, r = pix % hw;
    const float* hp = p.h + (size_t)pix * p.h_ld;
    float* zp = p.z + (size_t)pix * p.z_ld;
    for (int c = 0; c < p.Cz; ++c) {
      const float mean = __ldg(hp + 2 * c);
      float logs = __ldg(hp + 2 * c + 1);
      if (p.atan_logscale) logs = coupling_logscale(logs);
      const size_t nchw = ((size_t)b * p.Cz + c) * hw + r;
      if (MODE == PRIOR_SAMPLE) {
        const float e = p.eps ? __ldg(p.eps + nchw) : 0.f;
        zp[c] = mean + expf(logs) * e;
      } else if (MODE == PRIOR_LOGP) {
        const float d = zp[c] - mean;
        lp += -0.5f * (logs * 2.f + (d * d) / expf(logs * 2.f) + 1.8378770664093453f);
      } else {
        p.out_nchw[nchw] = (zp[c] - mean) * expf(-logs);
      }
    }
  }
  if (MODE == PRIOR_LOGP) add_logdet(p.logdet, active ? pix / hw : 0, (double)lp, active);
}

__global__ void __launch_bounds__(256) gauss_logp_const_kernel(const float* __restrict__ x,
                                                               const float* __restrict__ mean, float logs,
                                                               int n, double* logdet) {
  // grid.y = image, grid.x strides over the n = C*H*W elements of that image
  const int b = blockIdx.y;
  const float var = expf(2.f * logs);
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float d = x[(size_t)b * n + i] - mean[(size_t)b * n + i];
    s += (double)(-0.5f * (logs * 2.f + (d * d) / var + 1.8378770664093453f));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(logdet + b, t);
  }
}

static int to_params(const hcf_step_args* a, StepParams& p, bool need_h) {
  HCF_REQUIRE(a != nullptr && a->z != nullptr, "step: null args");
  HCF_REQUIRE(a->npix > 0 && a->pix_per_img > 0 && a->npix % a->pix_per_img == 0, "step: npix");
  HCF_REQUIRE(a->C > 0 && a->z_ld >= a->C, "step: C %d ld %d", a->C, a->z_ld);
  HCF_REQUIRE(!need_h || a->h != nullptr, "step: missing coupling input h");
  HCF_REQUIRE(a->mode == HCF_COUPLING_AFFINE || a->mode == HCF_COUPLING_SHIFT_FIRST3, "step: mode");
  HCF_REQUIRE(a->n_pass >= 0 && a->n_pass <= a->C, "step: n_pass");
  p.npix = a->npix; p.pix_per_img = a->pix_per_img; p.z = a->z; p.z_ld = a->z_ld;
  p.h = a->h; p.h_ld = a->h_ld; p.mode = a->mode; p.n_pass = a->n_pass;
  p.w = a->w; p.an_scale = a->an_scale; p.an_bias = a->an_bias; p.logdet = a->logdet;
  return 0;
}

#define HCF_DISPATCH_C(C, KERNEL, ...)                                       \
  switch (C) {                                                               \
    case 3: KERNEL<3> __VA_ARGS__; break;                                    \
    case 6: KERNEL<6> __VA_ARGS__; break;                                    \
    case 9: KERNEL<9> __VA_ARGS__; break;                                    \
    case 12: KERNEL<12> __VA_ARGS__; break;                                  \
    case 21: KERNEL<21> __VA_ARGS__; break;                                  \
    case 24: KERNEL<24> __VA_ARGS__; break;                                  \
    case 45: KERNEL<45> __VA_ARGS__; break;                                  \
    case 48: KERNEL<48> __VA_ARGS__; break;                                  \
    default:                                                                 \
      ::hcf::set_error("step: channel count %d not instantiated", C);        \
      return HCF_ENOTSUP;                                                    \
  }

}  // namespace hcf

extern "C" int hcf_step_inverse(const hcf_step_args* a, void* stream) {
  using namespace hcf;
  StepParams p;
  int rc = to_params(a, p, false);
  if (rc) return rc;
  const dim3 grid(ceil_div(p.npix, STEP_THREADS));
  cudaStream_t st = (cudaStream_t)stream;
  HCF_DISPATCH_C(a->C, step_inverse_kernel, <<<grid, STEP_THREADS, 0, st>>>(p));
  return finish_launch("hcf_step_inverse");
}

extern "C" int hcf_step_forward_head(const hcf_step_args* a, void* stream) {
  using namespace hcf;
  StepParams p;
  int rc = to_params(a, p, false);
  if (rc) return rc;
  const dim3 grid(ceil_div(p.npix, STEP_THREADS));
  cudaStream_t st = (cudaStream_t)stream;
  HCF_DISPATCH_C(a->C, step_forward_head_kernel, <<<grid, STEP_THREADS, 0, st>>>(p));
  return finish_launch("hcf_step_forward_head");
}

extern "C" int hcf_step_forward_coupling(const hcf_step_args* a, void* stream) {
  using namespace hcf;
  StepParams p;
  int rc = to_params(a, p, true);
  if (rc) return rc;
  const dim3 grid(ceil_div(p.npix, STEP_THREADS));
  step_forward_coupling_kernel<<<grid, STEP_THREADS, 0, (cudaStream_t)stream>>>(p, a->C);
  return finish_launch("hcf_step_forward_coupling");
}

static int prior_params(const hcf_prior_args* a, hcf::PriorParams& p) {
  using namespace hcf;
  HCF_REQUIRE(a != nullptr && a->h != nullptr && a->z != nullptr, "prior: null args");
  HCF_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->Cz > 0, "prior: shape");
  HCF_REQUIRE(a->h_ld >= 2 * a->Cz && a->z_ld >= a->Cz, "prior: ld");
  p.B = a->B; p.H = a->H; p.W = a->W; p.Cz = a->Cz; p.h = a->h; p.h_ld = a->h_ld;
  p.atan_logscale = a->atan_logscale; p.eps = a->eps_nchw; p.z = a->z; p.z_ld = a->z_ld;
  p.logdet = a->logdet; p.out_nchw = a->out_nchw;
  return 0;
}

extern "C" int hcf_prior_sample(const hcf_prior_args* a, void* stream) {
  using namespace hcf;
  PriorParams p;
  int rc = prior_params(a, p);
  if (rc) return rc;
  prior_kernel<PRIOR_SAMPLE><<<ceil_div(p.B * p.H * p.W, STEP_THREADS), STEP_THREADS, 0, (cudaStream_t)stream>>>(p);
  return finish_launch("hcf_prior_sample");
}

extern "C" int hcf_prior_logp(const hcf_prior_args* a, void* stream) {
  using namespace hcf;
  PriorParams p;
  int rc = prior_params(a, p);
  if (rc) return rc;
  HCF_REQUIRE(p.logdet != nullptr, "prior_logp: logdet is NULL");
  prior_kernel<PRIOR_LOGP><<<ceil_div(p.B * p.H * p.W, STEP_THREADS), STEP_THREADS, 0, (cudaStream_t)stream>>>(p);
  return finish_launch("hcf_prior_logp");
}

extern "C" int hcf_prior_standardize(const hcf_prior_args* a, void* stream) {
  using namespace hcf;
  PriorParams p;
  int rc = prior_params(a, p);
  if (rc) return rc;
  HCF_REQUIRE(p.out_nchw != nullptr, "prior_standardize: out is NULL");
  prior_kernel<PRIOR_STD><<<ceil_div(p.B * p.H * p.W, STEP_THREADS), STEP_THREADS, 0, (cudaStream_t)stream>>>(p);
  return finish_launch("hcf_prior_standardize");
}

extern "C" int hcf_gauss_logp_const(const float* x, const float* mean, float logs, int32_t B, int32_t n,
                                    double* logdet, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(x && mean && logdet && B > 0 && n > 0, "gauss_logp_const: bad args");
  dim3 grid((unsigned)min(ceil_div(n, 256), 64), (unsigned)B);
  gauss_logp_const_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, mean, logs, n, logdet);
  return finish_launch("hcf_gauss_logp_const");
}
