// 3x3 convolution as an implicit GEMM on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   D[128 pixels, N] += A[128 pixels, 32 ch] * B[N, 32 ch]^T     per (tap, 32-channel chunk)
//
// One CTA owns a 16x8 pixel tile of one image (UMMA M = 128) and all N = ceil16(Cout) <= 64
// output channels; the fp32 accumulator lives in TMEM (N columns).  Per 32-channel chunk the
// producer thread issues
//   * ONE 4-D TMA tile load of the 18x10 halo tile [18][10][32 ch] (128-byte swizzle, out of
//     bounds -> 0, which is exactly the conv's zero padding), and
//   * one bulk copy of the pre-swizzled weight slab [9 taps][N][32 ch];
// the MMA thread then issues 9 taps x 4 K-steps of tcgen05.mma.kind::tf32.  The nine taps
// do NOT reload the activations: tap (dy,dx) is a shared-memory descriptor whose start is
// shifted by (dy*10+dx) 128-byte rows into the same halo tile and whose 8-row-group stride
// (SBO) is the halo row pitch (10*128 B) -- the 128B-swizzle XOR is a function of the
// absolute smem address bits, so a shifted view of a TMA-written tile stays consistent.
// (HCF_TC_SAFE_A=1 selects a diagnostic variant that loads nine separate aligned tiles.)
//
// passes = 3 runs the K loop three times (A_raw*B_raw, A_raw*B_lo, A_lo*B_raw): the tensor
// core reads an fp32 word as TF32 by ignoring the low 13 mantissa bits, so "hi" parts are
// free, B_lo is precomputed on the host and A_lo = a - trunc(a) is formed in place in smem
// by the epilogue warps before the MMA of that stage (3xTF32 split, ~fp32 accuracy).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = A_lo conversion during the main loop, then epilogue (tcgen05.ld -> bias /
// scale / activation / residuals -> global).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace hcf {
namespace tc {

constexpr int TH = 16, TW = 8;               // pixel tile (UMMA M = 128)
constexpr int HALO_W = TW + 2, HALO_H = TH + 2;
constexpr int KCH = 32;                      // channels per K chunk (= 128 B rows)
constexpr int ROW_BYTES = KCH * 4;           // 128
constexpr int A_HALO_BYTES = HALO_H * HALO_W * ROW_BYTES;   // 23040
constexpr int A_HALO_STAGE = 23552;                         // padded to a 1024 B multiple
constexpr int A_SAFE_TILE = TH * TW * ROW_BYTES;            // 16384 per tap
constexpr int NTHREADS = 192;
constexpr int SMEM_LIMIT = 227 * 1024;

struct Params {
  int B, H, W;
  int kchunks;   // Cin / 32
  int N;         // UMMA N (multiple of 16, <= 64)
  int cout;
  int passes;    // 1 or 3
  int stages;
  int safe_a;    // diagnostic: nine aligned A tiles per stage instead of one halo tile
  int tiles_x, tiles_y;
  const float* wimg;   // [2][kchunks][9][N][32] pre-swizzled (raw, lo)
  const float* bias;
  const float* scale;
  int act;
  float* out; int out_ld;
  float* out2; int out2_ld;
  const float* res1; int res1_ld; float alpha1;
  const float* res2; int res2_ld; float alpha2;
  int out_vec;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, 128B-swizzled operand: rows of 128 B, 8-row groups `sbo_bytes` apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address          bits [0,14)
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32; // stride byte offset     bits [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(NTHREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap amap, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A | B)] then barriers
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_stage_bytes = p.safe_a ? 9u * A_SAFE_TILE : (uint32_t)A_HALO_STAGE;
  const uint32_t b_stage_bytes = 9u * p.N * ROW_BYTES;
  const uint32_t stage_bytes = a_stage_bytes + b_stage_bytes;
  const uint32_t bar_base = smem_base + p.stages * stage_bytes;
  // barriers: full[s], empty[s], conv[s], tmem_full ; then the TMEM base address slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto conv_bar = [&](int s) { return bar_base + 8u * (2 * p.stages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (3 * p.stages);
  const uint32_t tmem_slot = tmem_full_bar + 8u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));   // generic pointer to smem_base

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tile = blockIdx.x;
  const int tile_x = tile % p.tiles_x;
  tile /= p.tiles_x;
  const int tile_y = tile % p.tiles_y;
  const int b = tile / p.tiles_y;
  const int y0 = tile_y * TH, x0 = tile_x * TW;
  const int iters = p.passes * p.kchunks;
  const uint32_t tmem_cols = p.N <= 32 ? 32u : 64u;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&amap) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), p.passes == 3 ? 129u : 1u);   // MMA commit (+ the 128 converter threads)
      mbar_init(conv_bar(s), 128);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const int pass = it / p.kchunks, kc = it % p.kchunks;
        const uint32_t a_dst = smem_base + s * stage_bytes;
        const uint32_t b_dst = a_dst + a_stage_bytes;
        const uint32_t a_bytes = p.safe_a ? 9u * A_SAFE_TILE : (uint32_t)A_HALO_BYTES;
        mbar_expect_tx(full_bar(s), a_bytes + b_stage_bytes);
        if (!p.safe_a) {
          tma_load_4d(a_dst, &amap, full_bar(s), kc * KCH, x0 - 1, y0 - 1, b);
        } else {
          for (int t = 0; t < 9; ++t)
            tma_load_4d(a_dst + t * A_SAFE_TILE, &amap, full_bar(s), kc * KCH, x0 + (t % 3) - 1, y0 + (t / 3) - 1, b);
        }
        const int bpart = (pass == 1) ? 1 : 0;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wimg) +
                             (size_t)(bpart * p.kchunks + kc) * b_stage_bytes;
        bulk_load(b_dst, src, b_stage_bytes, full_bar(s));
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
      uint32_t conv_phase = 0;   // bit s = parity of the next completion of conv_bar(s)
      for (int it = 0; it < iters; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        const int pass = it / p.kchunks;
        mbar_wait(full_bar(s), ph);
        if (pass == 2) {
          mbar_wait(conv_bar(s), (conv_phase >> s) & 1u);
          conv_phase ^= 1u << s;
        }
        tc_fence_after();
        const uint32_t a_src = smem_base + s * stage_bytes;
        const uint32_t b_src = a_src + a_stage_bytes;
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
          const int dy = t / 3, dx = t % 3;
          const uint32_t a_tap = p.safe_a ? a_src + t * A_SAFE_TILE : a_src + (dy * HALO_W + dx) * ROW_BYTES;
          const uint32_t a_sbo = p.safe_a ? 8u * ROW_BYTES : (uint32_t)(HALO_W * ROW_BYTES);
          const uint32_t b_tap = b_src + t * p.N * ROW_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_tf32(tmem_base, make_desc(a_tap + k * 32, a_sbo), make_desc(b_tap + k * 32, 8u * ROW_BYTES), idesc,
                      (it > 0 || t > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty_bar(s));   // frees the smem stage when these MMAs retire
      }
      umma_commit(tmem_full_bar);    // accumulator complete
    }
  } else {
    // ===================== converter (3-pass only), then epilogue =====================
    const int et = threadIdx.x - 64;   // 0..127
    if (p.passes == 3) {
      // The converters follow EVERY stage phase in order (an mbarrier waiter may never fall
      // two phases behind) and release the stage together with the MMA commit.
      for (int it = 0; it < iters; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(full_bar(s), ph);
        if (it >= 2 * p.kchunks) {
          float4* a = reinterpret_cast<float4*>(gen_base + (size_t)s * stage_bytes);
          const int n4 = (p.safe_a ? 9 * A_SAFE_TILE : A_HALO_BYTES) / 16;
          for (int i = et; i < n4; i += 128) {
            float4 v = a[i];
            v.x -= __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
            v.y -= __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
            v.z -= __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
            v.w -= __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
            a[i] = v;
          }
          fence_async_smem();
          mbar_arrive(conv_bar(s));
        }
        mbar_arrive(empty_bar(s));
      }
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;                  // accumulator row = pixel within the tile
    const int gy = y0 + m / TW, gx = x0 + m % TW;
    const bool inb = (gy < p.H) && (gx < p.W);
    const size_t pix = ((size_t)b * p.H + (inb ? gy : 0)) * p.W + (inb ? gx : 0);
    for (int c0 = 0; c0 < p.N; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (!inb || c0 >= p.cout) continue;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = c0 + j;
        float t = v[j];
        if (p.bias) t += __ldg(p.bias + c);
        if (p.scale) t *= __ldg(p.scale + c);
        if (p.act == HCF_ACT_RELU) t = fmaxf(t, 0.f);
        else if (p.act == HCF_ACT_LRELU) t = t > 0.f ? t : 0.2f * t;
        v[j] = t;
      }
      if (p.out_vec && c0 + 15 < p.cout) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (p.res1) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(p.res1 + pix * p.res1_ld + c0 + j));
            o.x = o.x * p.alpha1 + r.x; o.y = o.y * p.alpha1 + r.y; o.z = o.z * p.alpha1 + r.z; o.w = o.w * p.alpha1 + r.w;
          }
          if (p.res2) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(p.res2 + pix * p.res2_ld + c0 + j));
            o.x = o.x * p.alpha2 + r.x; o.y = o.y * p.alpha2 + r.y; o.z = o.z * p.alpha2 + r.z; o.w = o.w * p.alpha2 + r.w;
          }
          *reinterpret_cast<float4*>(p.out + pix * p.out_ld + c0 + j) = o;
          if (p.out2) *reinterpret_cast<float4*>(p.out2 + pix * p.out2_ld + c0 + j) = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = c0 + j;
          if (c < p.cout) {
            float t = v[j];
            if (p.res1) t = t * p.alpha1 + __ldg(p.res1 + pix * p.res1_ld + c);
            if (p.res2) t = t * p.alpha2 + __ldg(p.res2 + pix * p.res2_ld + c);
            p.out[pix * p.out_ld + c] = t;
            if (p.out2) p.out2[pix * p.out2_ld + c] = t;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static int n_for(int cout) { return (cout + 15) / 16 * 16; }

static int stages_for(int N, int safe_a) {
  const int a = safe_a ? 9 * A_SAFE_TILE : A_HALO_STAGE;
  const int per = a + 9 * N * ROW_BYTES;
  int s = (SMEM_LIMIT - 2048) / per;
  if (s > 4) s = 4;
  return s;
}

}  // namespace tc
}  // namespace hcf

struct hcf_conv_tc_plan {
  CUtensorMap amap;
  hcf::tc::Params p;
  size_t smem_bytes;
  dim3 grid;
};

extern "C" int hcf_conv_tc_supported(const hcf_conv_args* a) {
  if (!a) return 0;
  if (a->ks != 3 || a->nseg != 1) return 0;
  if (a->seg[0].up_shift != 0 || a->seg[0].C % 32 != 0 || a->seg[0].C < 32) return 0;
  if (a->seg[0].ld % 4 != 0 || !hcf::aligned16(a->seg[0].ptr)) return 0;
  if (a->cout < 1 || a->cout > 64) return 0;
  return 1;
}

extern "C" int64_t hcf_conv_tc_weight_bytes(int32_t kin, int32_t cout) {
  if (kin % 32 != 0 || cout < 1 || cout > 64) return 0;
  return (int64_t)2 * (kin / 32) * 9 * hcf::tc::n_for(cout) * 128;
}

extern "C" int hcf_conv_tc_pack_weights(const float* w, int32_t kin, int32_t cout, float* image) {
  using namespace hcf;
  HCF_REQUIRE(w && image && kin % 32 == 0 && cout >= 1 && cout <= 64, "tc_pack: bad args");
  const int N = tc::n_for(cout), KC = kin / 32;
  memset(image, 0, (size_t)hcf_conv_tc_weight_bytes(kin, cout));
  for (int part = 0; part < 2; ++part)
    for (int kc = 0; kc < KC; ++kc)
      for (int t = 0; t < 9; ++t)
        for (int n = 0; n < cout; ++n)
          for (int j = 0; j < 32; ++j) {
            const float v = w[(((size_t)n * kin + kc * 32 + j) * 3 + t / 3) * 3 + t % 3];
            uint32_t bits;
            memcpy(&bits, &v, 4);
            bits &= 0xFFFFE000u;
            float hi;
            memcpy(&hi, &bits, 4);
            const float val = part == 0 ? v : v - hi;
            const int chunk = (j / 4) ^ (n & 7);   // 128B swizzle: 16-byte chunk index XOR row-in-atom
            image[((((size_t)part * KC + kc) * 9 + t) * N + n) * 32 + chunk * 4 + (j & 3)] = val;
          }
  return 0;
}

extern "C" int hcf_conv_tc_plan_create(const hcf_conv_args* a, const float* wtc, int32_t passes,
                                       hcf_conv_tc_plan** out) {
  using namespace hcf;
  HCF_REQUIRE(out != nullptr, "tc_plan: null out");
  *out = nullptr;
  int rc = validate_conv_args(a);
  if (rc) return rc;
  HCF_REQUIRE(hcf_conv_tc_supported(a), "tc_plan: unsupported shape");
  HCF_REQUIRE(wtc && aligned16(wtc), "tc_plan: weight image alignment");
  HCF_REQUIRE(passes == 1 || passes == 3, "tc_plan: passes %d", passes);
  tc::EncodeTiledFn enc = tc::get_encode();
  HCF_REQUIRE(enc != nullptr, "tc_plan: cuTensorMapEncodeTiled entry point not found");
  hcf_conv_tc_plan* pl = new hcf_conv_tc_plan();
  tc::Params& p = pl->p;
  const char* env = getenv("HCF_TC_SAFE_A");
  p.safe_a = (env && env[0] == '1') ? 1 : 0;
  p.B = a->B; p.H = a->H; p.W = a->W;
  p.kchunks = a->seg[0].C / 32;
  p.N = tc::n_for(a->cout);
  p.cout = a->cout;
  p.passes = passes;
  p.stages = tc::stages_for(p.N, p.safe_a);
  if (p.stages < 1) {
    delete pl;
    set_error("tc_plan: tile does not fit in shared memory");
    return HCF_ENOTSUP;
  }
  p.tiles_x = ceil_div(a->W, tc::TW); p.tiles_y = ceil_div(a->H, tc::TH);
  p.wimg = wtc; p.bias = a->bias; p.scale = a->scale; p.act = a->act;
  p.out = a->out; p.out_ld = a->out_ld; p.out2 = a->out2; p.out2_ld = a->out2_ld;
  p.res1 = a->res1; p.res1_ld = a->res1_ld; p.alpha1 = a->alpha1;
  p.res2 = a->res2; p.res2_ld = a->res2_ld; p.alpha2 = a->alpha2;
  bool ov = aligned16(a->out) && a->out_ld % 4 == 0;
  if (a->out2) ov = ov && aligned16(a->out2) && a->out2_ld % 4 == 0;
  if (a->res1) ov = ov && aligned16(a->res1) && a->res1_ld % 4 == 0;
  if (a->res2) ov = ov && aligned16(a->res2) && a->res2_ld % 4 == 0;
  p.out_vec = ov ? 1 : 0;

  const cuuint64_t dims[4] = {(cuuint64_t)a->seg[0].C, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->B};
  const cuuint64_t ld_b = (cuuint64_t)a->seg[0].ld * 4;
  const cuuint64_t strides[3] = {ld_b, ld_b * a->W, ld_b * a->W * a->H};
  const cuuint32_t box_halo[4] = {(cuuint32_t)tc::KCH, (cuuint32_t)tc::HALO_W, (cuuint32_t)tc::HALO_H, 1};
  const cuuint32_t box_safe[4] = {(cuuint32_t)tc::KCH, (cuuint32_t)tc::TW, (cuuint32_t)tc::TH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&pl->amap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a->seg[0].ptr), dims, strides,
                   p.safe_a ? box_safe : box_halo, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    delete pl;
    set_error("tc_plan: cuTensorMapEncodeTiled failed with %d", (int)r);
    return HCF_EINVAL;
  }
  const size_t a_stage = p.safe_a ? 9 * tc::A_SAFE_TILE : tc::A_HALO_STAGE;
  pl->smem_bytes = 1024 + p.stages * (a_stage + 9 * p.N * tc::ROW_BYTES) + 8 * (3 * p.stages + 1) + 16;
  pl->grid = dim3((unsigned)(p.tiles_x * p.tiles_y * p.B));
  cudaError_t e = cudaFuncSetAttribute(tc::conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tc::SMEM_LIMIT);
  if (e != cudaSuccess) {
    delete pl;
    set_error("tc_plan: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return (int)e;
  }
  *out = pl;
  return 0;
}

extern "C" int hcf_conv_tc_run(const hcf_conv_tc_plan* pl, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(pl != nullptr, "tc_run: null plan");
  tc::conv3x3_tc_kernel<<<pl->grid, tc::NTHREADS, pl->smem_bytes, (cudaStream_t)stream>>>(pl->amap, pl->p);
  return finish_launch("hcf_conv_tc_run");
}

extern "C" void hcf_conv_tc_plan_destroy(hcf_conv_tc_plan* p) { delete p; }
