// fp32 CUDA-core convolution (exact-fp32 parity mode, and every shape the tensor-core
// kernel does not take: Cin in {3,6,10,...}, Cout in {6,12,22,...}, nearest-upsampled or
// concatenated inputs).  NHWC, stride 1, "same" padding, ks in {1,3}.
//
// One CTA = 256 threads = an 8x16 pixel tile x BN output channels.  K is walked in chunks
// of 8 input channels: the halo tile (10x18 pixels x 8 ch) and the 9x8xBN weight slab are
// staged in shared memory; each thread owns TM consecutive pixels of one tile row x 4
// output channels and slides the three horizontal taps over TM+2 float4 loads, so the
// inner loop is ~17 FFMA per LDS.128.  The concat / nearest-upsample / channel-slice
// plumbing of the reference (torch.cat, F.interpolate, Split) is address arithmetic in
// the loader, and ActNorm / exp(3*logs) / (Leaky)ReLU / residual scaling live in the
// epilogue (see include/hcflow_b200.h for the reference lines each replaces).
#include "common.cuh"

namespace hcf {

struct SegDev {
  const float* ptr;
  int ld, C, up, kbase, vec;
};

struct ConvParams {
  int B, H, W, nseg;
  SegDev seg[3];
  int kpad, cout, npad;
  const float* w;
  const float* bias;
  const float* scale;
  int act;
  float* out;
  int out_ld;
  float* out2;
  int out2_ld;
  const float* res1;
  int res1_ld;
  float alpha1;
  const float* res2;
  int res2_ld;
  float alpha2;
  int tiles_x, tiles_y;
  int out_vec;
};

constexpr int TH = 8, TW = 16, KC = 8, NT = 256;

__device__ __forceinline__ float pick(const float4& v, int k) {
  return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

template <int KS, int BN>
__global__ void __launch_bounds__(NT, 2) conv_fp32_kernel(const ConvParams p) {
  constexpr int HALO = KS / 2;
  constexpr int IW = TW + 2 * HALO, IH = TH + 2 * HALO;
  constexpr int NG = BN / 4;             // channel groups (float4 each)
  constexpr int PG = NT / NG;            // pixel groups
  constexpr int TM = (TH * TW) / PG;     // consecutive pixels per thread
  constexpr int TAPS = KS * KS;
  static_assert(TM >= 1 && TW % TM == 0, "tile split");
  __shared__ __align__(16) float s_in[IH * IW * KC];
  __shared__ __align__(16) float s_w[TAPS * KC * BN];

  const int tid = threadIdx.x;
  int tile = blockIdx.x;
  const int tile_x = tile % p.tiles_x;
  tile /= p.tiles_x;
  const int tile_y = tile % p.tiles_y;
  const int b = tile / p.tiles_y;
  const int y0 = tile_y * TH, x0 = tile_x * TW;
  const int n0 = blockIdx.y * BN;

  const int tx = tid % NG, ty = tid / NG;
  const int prow = (ty * TM) / TW, pcol = (ty * TM) % TW;

  float acc[TM][4];
#pragma unroll
  for (int m = 0; m < TM; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;

  for (int kc = 0; kc < p.kpad; kc += KC) {
    int s = 0;
    if (p.nseg > 1 && kc >= p.seg[1].kbase) s = 1;
    if (p.nseg > 2 && kc >= p.seg[2].kbase) s = 2;
    const float* sptr = p.seg[s].ptr;
    const int sld = p.seg[s].ld, sC = p.seg[s].C, sup = p.seg[s].up, svec = p.seg[s].vec;
    const int c0 = kc - p.seg[s].kbase;
    const int Hs = p.H >> sup, Ws = p.W >> sup;

    for (int idx = tid; idx < IH * IW * 2; idx += NT) {
      const int pix = idx >> 1, half = idx & 1;
      const int iy = pix / IW, ix = pix % IW;
      const int gy = y0 + iy - HALO, gx = x0 + ix - HALO;
      const int c = c0 + half * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W && c < sC) {
        const float* src = sptr + ((size_t)(b * Hs + (gy >> sup)) * Ws + (gx >> sup)) * sld + c;
        if (svec && c + 3 < sC) {
          v = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          v.x = __ldg(src);
          if (c + 1 < sC) v.y = __ldg(src + 1);
          if (c + 2 < sC) v.z = __ldg(src + 2);
          if (c + 3 < sC) v.w = __ldg(src + 3);
        }
      }
      *reinterpret_cast<float4*>(&s_in[pix * KC + half * 4]) = v;
    }
    for (int idx = tid; idx < TAPS * KC * NG; idx += NT) {
      const int n4 = idx % NG, rk = idx / NG;
      const int tap = rk / KC, k = rk % KC;
      const float4 v = __ldg(reinterpret_cast<const float4*>(
          p.w + ((size_t)(tap * p.kpad + kc + k)) * p.npad + n0 + n4 * 4));
      *reinterpret_cast<float4*>(&s_w[rk * BN + n4 * 4]) = v;
    }
    __syncthreads();

#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
#pragma unroll
      for (int k4 = 0; k4 < 2; ++k4) {
        float4 a[TM + KS - 1];
#pragma unroll
        for (int j = 0; j < TM + KS - 1; ++j)
          a[j] = *reinterpret_cast<const float4*>(&s_in[((prow + ky) * IW + pcol + j) * KC + k4 * 4]);
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float4 wv =
                *reinterpret_cast<const float4*>(&s_w[((ky * KS + kx) * KC + k4 * 4 + kk) * BN + tx * 4]);
#pragma unroll
            for (int m = 0; m < TM; ++m) {
              const float av = pick(a[m + kx], kk);
              acc[m][0] = fmaf(av, wv.x, acc[m][0]);
              acc[m][1] = fmaf(av, wv.y, acc[m][1]);
              acc[m][2] = fmaf(av, wv.z, acc[m][2]);
              acc[m][3] = fmaf(av, wv.w, acc[m][3]);
            }
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue
  const int c = n0 + tx * 4;
  if (c >= p.cout) return;
  float bs[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {1.f, 1.f, 1.f, 1.f};
  if (p.bias) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + c));
    bs[0] = t.x; bs[1] = t.y; bs[2] = t.z; bs[3] = t.w;
  }
  if (p.scale) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p.scale + c));
    sc[0] = t.x; sc[1] = t.y; sc[2] = t.z; sc[3] = t.w;
  }
  const int gy = y0 + prow;
  if (gy >= p.H) return;
  const bool full = p.out_vec && (c + 3 < p.cout);
#pragma unroll
  for (int m = 0; m < TM; ++m) {
    const int gx = x0 + pcol + m;
    if (gx >= p.W) continue;
    const size_t pix = ((size_t)b * p.H + gy) * p.W + gx;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = (acc[m][j] + bs[j]) * sc[j];
      if (p.act == HCF_ACT_RELU) t = fmaxf(t, 0.f);
      else if (p.act == HCF_ACT_LRELU) t = t > 0.f ? t : 0.2f * t;
      v[j] = t;
    }
    if (full) {
      if (p.res1) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(p.res1 + pix * p.res1_ld + c));
        v[0] = v[0] * p.alpha1 + r.x; v[1] = v[1] * p.alpha1 + r.y;
        v[2] = v[2] * p.alpha1 + r.z; v[3] = v[3] * p.alpha1 + r.w;
      }
      if (p.res2) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(p.res2 + pix * p.res2_ld + c));
        v[0] = v[0] * p.alpha2 + r.x; v[1] = v[1] * p.alpha2 + r.y;
        v[2] = v[2] * p.alpha2 + r.z; v[3] = v[3] * p.alpha2 + r.w;
      }
      const float4 o = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(p.out + pix * p.out_ld + c) = o;
      if (p.out2) *reinterpret_cast<float4*>(p.out2 + pix * p.out2_ld + c) = o;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (c + j < p.cout) {
          float t = v[j];
          if (p.res1) t = t * p.alpha1 + __ldg(p.res1 + pix * p.res1_ld + c + j);
          if (p.res2) t = t * p.alpha2 + __ldg(p.res2 + pix * p.res2_ld + c + j);
          p.out[pix * p.out_ld + c + j] = t;
          if (p.out2) p.out2[pix * p.out2_ld + c + j] = t;
        }
      }
    }
  }
}

template <int KS, int BN>
static int launch(const ConvParams& p, cudaStream_t st) {
  dim3 grid((unsigned)(p.tiles_x * p.tiles_y * p.B), (unsigned)(p.npad / BN));
  conv_fp32_kernel<KS, BN><<<grid, NT, 0, st>>>(p);
  return finish_launch("hcf_conv_fp32");
}

int validate_conv_args(const hcf_conv_args* a) {
  HCF_REQUIRE(a != nullptr, "conv: null args");
  HCF_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0, "conv: bad shape %d %d %d", a->B, a->H, a->W);
  HCF_REQUIRE(a->nseg >= 1 && a->nseg <= 3, "conv: nseg %d", a->nseg);
  HCF_REQUIRE(a->ks == 1 || a->ks == 3, "conv: ks %d", a->ks);
  HCF_REQUIRE(a->cout > 0 && a->npad >= a->cout, "conv: cout %d npad %d", a->cout, a->npad);
  HCF_REQUIRE(a->npad == 16 || a->npad == 32 || a->npad % 64 == 0, "conv: npad %d", a->npad);
  int k = 0;
  for (int i = 0; i < a->nseg; ++i) {
    HCF_REQUIRE(a->seg[i].ptr && a->seg[i].C > 0 && a->seg[i].ld >= a->seg[i].C, "conv: seg %d", i);
    HCF_REQUIRE(a->seg[i].up_shift >= 0 && a->seg[i].up_shift <= 3, "conv: up_shift");
    HCF_REQUIRE((a->H % (1 << a->seg[i].up_shift)) == 0 && (a->W % (1 << a->seg[i].up_shift)) == 0,
                "conv: H,W not divisible by the upsampling factor");
    k += (a->seg[i].C + 7) / 8 * 8;
  }
  HCF_REQUIRE(k == a->kpad, "conv: kpad %d != %d", a->kpad, k);
  HCF_REQUIRE(a->w && aligned16(a->w), "conv: weights must be 16B aligned");
  HCF_REQUIRE(a->out && a->out_ld >= a->cout, "conv: out");
  HCF_REQUIRE(!a->bias || aligned16(a->bias), "conv: bias alignment");
  HCF_REQUIRE(!a->scale || aligned16(a->scale), "conv: scale alignment");
  HCF_REQUIRE(!a->pre || (!a->res1 && !a->res2 && a->pre_ld >= a->cout), "conv: pre excludes res1 / res2");
  return 0;
}

}  // namespace hcf

extern "C" int hcf_conv_fp32(const hcf_conv_args* a, void* stream) {
  if (a && (a->pre || a->step || a->raw2)) {
    hcf::set_error("conv_fp32: the pre-activation addend and the fused FlowStep are tensor-core kernel features");
    return HCF_ENOTSUP;
  }
  using namespace hcf;
  int rc = validate_conv_args(a);
  if (rc) return rc;
  ConvParams p;
  p.B = a->B; p.H = a->H; p.W = a->W; p.nseg = a->nseg;
  int k = 0;
  for (int i = 0; i < 3; ++i) {
    if (i < a->nseg) {
      p.seg[i].ptr = a->seg[i].ptr; p.seg[i].ld = a->seg[i].ld; p.seg[i].C = a->seg[i].C;
      p.seg[i].up = a->seg[i].up_shift; p.seg[i].kbase = k;
      p.seg[i].vec = (aligned16(a->seg[i].ptr) && (a->seg[i].ld % 4 == 0)) ? 1 : 0;
      k += (a->seg[i].C + 7) / 8 * 8;
    } else {
      p.seg[i].ptr = nullptr; p.seg[i].ld = 0; p.seg[i].C = 0; p.seg[i].up = 0;
      p.seg[i].kbase = 1 << 30; p.seg[i].vec = 0;
    }
  }
  p.kpad = a->kpad; p.cout = a->cout; p.npad = a->npad;
  p.w = a->w; p.bias = a->bias; p.scale = a->scale; p.act = a->act;
  p.out = a->out; p.out_ld = a->out_ld; p.out2 = a->out2; p.out2_ld = a->out2_ld;
  p.res1 = a->res1; p.res1_ld = a->res1_ld; p.alpha1 = a->alpha1;
  p.res2 = a->res2; p.res2_ld = a->res2_ld; p.alpha2 = a->alpha2;
  p.tiles_x = ceil_div(a->W, TW); p.tiles_y = ceil_div(a->H, TH);
  bool ov = aligned16(a->out) && a->out_ld % 4 == 0;
  if (a->out2) ov = ov && aligned16(a->out2) && a->out2_ld % 4 == 0;
  if (a->res1) ov = ov && aligned16(a->res1) && a->res1_ld % 4 == 0;
  if (a->res2) ov = ov && aligned16(a->res2) && a->res2_ld % 4 == 0;
  p.out_vec = ov ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int bn = a->npad >= 64 ? 64 : a->npad;
  if (a->ks == 3) {
    if (bn == 64) return launch<3, 64>(p, st);
    if (bn == 32) return launch<3, 32>(p, st);
    return launch<3, 16>(p, st);
  }
  if (bn == 64) return launch<1, 64>(p, st);
  if (bn == 32) return launch<1, 32>(p, st);
  return launch<1, 16>(p, st);
}
