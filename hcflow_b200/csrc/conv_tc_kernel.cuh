// Device side of the tcgen05 convolution (conv_tc.cu includes this file; the host side -- ring sizing, tensor maps,
// layer tables, the C ABI -- lives there).
#pragma once
// 3x3 / 1x1 convolutions as a persistent implicit GEMM on the 5th-generation tensor cores
// (tcgen05 / TMEM / TMA, sm_100a), single or CHAINED, with fused epilogues up to a whole FlowStep tail.
//
//   D[128 pixels, N] += A[128 pixels, K chunk] * B[N, K chunk]^T     per (tap, chunk of one 128-byte row per pixel:
//                                                                      32 fp32 channels read as TF32, or 64 fp16)
//
// Work item = MT vertically adjacent 16x8 pixel tiles of one image (UMMA M = 128 each) x all
// N = ceil16(Cout) <= 128 output channels.  CTAs are persistent (grid = #SMs, items strided by
// gridDim.x) and warp-specialised.  A launch executes either one convolution or a CHAIN of
// dependent convolutions on the same pixel grid (the 213 convs of an RRDB encoder level; the shared
// conditioning convs + all coupling sub-nets + FlowStep tails of a level): item = (layer, tile) in
// layer-major order; a tile of layer l may start once the 3x3 tile neighbourhood of layer l-1 is
// complete, tracked by per-tile counters in global memory (release after the epilogue, acquire by
// the producer), so there is no kernel boundary, no ramp-up / drain and no wave quantisation
// between the convs.  Template: MT, PASSES (3: some layer uses the operand split), KS (3: 3x3 with
// 1x1 layers through the centre tap), F16 (fp16 hi / lo operand planes instead of TF32 words).  Roles:
//   warp 0      TMA producer.  Per chunk ONE 4-D TMA tile load brings the (16*MT+2) x 10 halo tile
//               [rows][10][128 B] (128-byte swizzle; out-of-bounds -> 0 is exactly the conv's zero
//               padding) into the A ring (fp16 split chunks: the hi and the lo plane); the
//               pre-swizzled weights stream through a second ring in slabs of 9 / 3 / 1 taps.  In a
//               chain it first polls the dependency counters (relaxed loads, prefetched one item
//               ahead), fences, and tells the epilogue groups that the item's inputs may be read.
//   warp 1      TMEM owner + MMA issuer: taps x K-steps of tcgen05.mma per chunk (only the K-steps
//               that hold real channels).  The nine taps do NOT reload activations: tap (dy,dx) is a
//               smem descriptor whose start is shifted by (dy*10+dx) 128-byte rows into the same halo
//               tile and whose 8-row-group stride (SBO) is the halo row pitch (1280 B).  The
//               128B-swizzle XOR is a function of the absolute smem address, so a shifted view
//               of a TMA-written tile stays consistent (verified on B200).  Split chunks issue
//               A_hi x [B_hi ; B_lo] (2N accumulator columns: main | correction) and A_lo x B_hi.
//   warps 2..5  epilogue group 0 (and warps 6..9 = group 1 in the fp16 kernels: group g drains TMEM
//               accumulator buffer g, i.e. every other item): tcgen05.ld -> per-warp staging transpose
//               in shared memory -> 8 lanes per pixel: pre-activation addend, bias / scale,
//               activation, residuals (prefetched tiles), full-line stores of the fp32 and / or
//               fp16 hi / lo representations -- or, for the last conv of a coupling sub-net, the
//               FlowStep inverse (affine coupling, W^-1, ActNorm) on z in place (thread = pixel).
//   warps 6..9  (TF32 kernels with PASSES == 3) form A_lo = a - trunc(a) next to every A stage; the
//               tensor core reads an fp32 word as TF32 by ignoring the low 13 mantissa bits.
//   warps 10,11 (fp16 kernels) publishers: one per epilogue group, take the gpu-scope release of a
//               finished tile (MEMBAR + counter increment) off the epilogue's path.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace hcf {
namespace tc {

constexpr int TH = 16, TW = 8;               // sub-tile (UMMA M = 128)
constexpr int KCH = 32;                      // channels per K chunk (= 128 B rows)
constexpr int ROW_BYTES = KCH * 4;           // 128
constexpr int KCH16 = 64;                    // fp16 kernels: channels per 128-byte row
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int NUM_SMS_FALLBACK = 148;
constexpr int BAR_BYTES = 512;               // mbarriers + TMEM slot
constexpr int STAGE_BYTES = 4 * 32 * 32 * 4;  // epilogue staging: 4 warps x 32 pixels x 32 channels fp32
constexpr int TAIL_BYTES = BAR_BYTES + 2 * (1024 + STAGE_BYTES);   // + (per-layer bias / scale + staging) x 2 epilogue groups

// ks = 3: (16*mt+2) x 10 halo tile; ks = 1: plain 16*mt x 8 tile
__host__ __device__ constexpr int halo_w(int ks) { return TW + (ks - 1); }
__host__ __device__ constexpr int halo_rows(int mt, int ks) { return TH * mt + (ks - 1); }
__host__ __device__ constexpr int a_bytes(int mt, int ks) { return halo_rows(mt, ks) * halo_w(ks) * ROW_BYTES; }
__host__ __device__ constexpr int a_part(int mt, int ks) { return (a_bytes(mt, ks) + 1023) / 1024 * 1024; }

constexpr int MAX_MAPS = 20;
struct Maps { CUtensorMap m[MAX_MAPS]; };

// One convolution of a chain.  Lives in global memory; every warp role reads the fields it needs
// at the start of a work item.
struct LayerDesc {
  int nseg;
  int map_idx[3];     // tensor map of each segment (fp16 kernels: the hi plane)
  int map_lo[3];      // fp16 kernels, split layers: tensor map of the lo plane
  int seg_coff[3];    // channel offset of the segment inside its tensor map (views of one buffer share a map)
  int seg_end[3];     // chunk index where segment i ends (prefix sums)
  int seg_last_k[3];  // K-steps (of 4 per chunk) that hold real channels in the LAST chunk of each segment
  int kchunks;        // total 32-channel chunks
  int N;              // UMMA N (multiple of 16, <= 128)
  int cout;
  int act;
  int out_vec;
  int parts;          // 1: one TF32 pass (B rows = N);  2: 3xTF32 split (B rows = [raw N ; lo N], plus A_lo x B)
  int split_kc;       // parts == 2: the split covers K chunks [0, split_kc) only (fp16 chains: the residual-stream
                      // channels of an RDB's conv5); the remaining chunks run one pass into the main columns
  int slab_taps;      // taps of this layer per B ring slot: KS*KS, KS or 1 (largest that fits the slot)
  int taps, tap0;     // taps of this layer (KS*KS, or 1 for a 1x1 conv inside a 3x3 chain) and the first tap index
  const float* wimg;  // [kchunks][ks dy][ks dx][NB][32] pre-swizzled; NB = N * parts
  const float* bias;
  const float* scale;
  float* out; int out_ld;
  float* out2; int out2_ld;
  __half* out_hi; __half* out_lo;     // fp16 kernels: hi / lo planes shadowing out (same ld), may be null
  __half* out2_hi; __half* out2_lo;
  const float* res1; int res1_ld; float alpha1;
  const float* res2; int res2_ld; float alpha2;
  const float* pre; int pre_ld;   // added to the accumulator before bias / scale / activation
  float* raw2; int raw2_ld;       // accumulator columns [32, cout) stored RAW as fp32 (a later conv's partial sum)
  // fused FlowStep inverse (FlowStep.py:55-64): this conv is the sub-net's last layer; its output h is not stored,
  // the epilogue applies  z2 = z2 * exp(-ls(h)) - shift(h);  z = W^-1 z;  z = z * exp(-logs) - bias  in place
  float* step_z; int step_z_ld, step_C, step_npass;
  const float* step_w; const float* step_sc; const float* step_b;
  __half* step_z16; int step_z16_ld;   // fp16 chains: hi plane of z[:, :n_pass] for the next step's first conv
  __half* step_z16_lo;                 // ... and its lo plane (that conv runs split), may be null
};

struct Params {
  int B, H, W;
  int tiles_x, tiles_y;
  int n_tiles;        // B * tiles_x * tiles_y  (work items per layer)
  int n_layers;       // > 1: a chain of dependent convs executed by ONE persistent launch
  int n_items;        // n_layers * n_tiles
  int tpc;            // 0: items rotate over the CTAs (item = blockIdx + q * grid); > 0: every CTA OWNS tpc tiles
                      // (n_tiles == tpc * grid) and walks the layers over them -- few-tile grids (40x40 level)
  int nb_max;         // max over layers of B rows per tap (N, or 2N in 3-pass mode)
  int sa, sb;         // ring depths
  int slot_bytes;     // bytes of one B ring slot
  int debug;          // timing experiments only (HCF_TC_DEBUG, wrong results): 1 aligned A descriptors,
                      // 2 no MMAs, 4 no loads, 8 no epilogue stores, 16 launch only, 32 prologue only,
                      // 64 no dependency waits, 128 no FlowStep arithmetic, 512 direct hi-plane stores off
  const LayerDesc* layers;
  int* done;          // chain mode: per-tile count of completed layers (zeroed before the launch)
  const float* epi;   // [n_layers][256]: bias (0) | scale (1) of every layer, padded: one coalesced load per layer change
  int step_tab;       // 1: the chain has fused FlowStep layers (shared memory carries their W^-1 / ActNorm tables)
  long long* prof;    // HCF_TC_PROF=1: cycles per role / wait class summed over CTAs (see PROF_* below)
  // weight-stationary schedule (conv_ws_kernel.cuh): image groups ("phases") of ipp images each, per-layer chunk tables
  int n_phases, ipp;
  const void* ws_layers;
  int* status;        // sticky device word (may be null): bit 0 = an fp16 operand plane saturated (|x| > 65504),
                      // bit 1 = a dependency wait timed out (the chain's CTAs were not co-resident)
};
constexpr int STATUS_F16_OVERFLOW = 1, STATUS_DEP_TIMEOUT = 2;

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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred;
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

// fp16 operand format: a = hi + lo / 2048 with hi = fp16(a), lo = fp16((a - hi) * 2048) (the scaling keeps the
// residual out of the fp16 subnormal range); hi alone carries 11 significant bits (round-to-nearest), hi + lo 22.
// Range guard: values beyond the fp16 range saturate to +-65504 (never inf: the stale x zero-weight-row trick of the
// ragged chunks would turn one inf into NaNs) and raise a sticky flag the host checks after the pass.
__device__ __forceinline__ uint2 split_hi(float4& o, int* __restrict__ status) {
  const float m = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w)));
  if (!(m <= 65504.0f)) {   // (also catches NaN)
    if (status) atomicOr(status, 1);
    o.x = fminf(fmaxf(o.x, -65504.0f), 65504.0f); o.y = fminf(fmaxf(o.y, -65504.0f), 65504.0f);
    o.z = fminf(fmaxf(o.z, -65504.0f), 65504.0f); o.w = fminf(fmaxf(o.w, -65504.0f), 65504.0f);
  }
  const __half2 a = __floats2half2_rn(o.x, o.y), b = __floats2half2_rn(o.z, o.w);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
__device__ __forceinline__ uint2 split_lo(const float4 o, const uint2 hi) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hi.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi.y));
  const __half2 la = __floats2half2_rn((o.x - a.x) * 2048.0f, (o.y - a.y) * 2048.0f);
  const __half2 lb = __floats2half2_rn((o.z - b.x) * 2048.0f, (o.w - b.y) * 2048.0f);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&la), *reinterpret_cast<const uint32_t*>(&lb));
}
struct Out16 { __half* hi; __half* lo; __half* hi2; __half* lo2; };

// Coalesced-domain store of one 32-column group when every chunk is a full float4 and the layer has a single
// output view: which representations are written is a template parameter (one specialisation per layer kind of
// a chain), so the per-pixel code is straight-line.  8 lanes cover the 32 channels of a pixel.
// per-lane epilogue constants of one 32-column group (the lane's 4 channels)
struct Chan4 { float4 bias, scale; float slope; };
// (o + bias) * scale, then max(o, slope * o): slope 1 = no activation, 0 = ReLU, 0.2 = LeakyReLU.  The inline table
// holds bias 0 / scale 1 for layers without them, so the arithmetic is branch-free.
__device__ __forceinline__ float4 chan_apply(float4 o, const Chan4& c) {
  o.x = (o.x + c.bias.x) * c.scale.x; o.y = (o.y + c.bias.y) * c.scale.y;
  o.z = (o.z + c.bias.z) * c.scale.z; o.w = (o.w + c.bias.w) * c.scale.w;
  o.x = fmaxf(o.x, c.slope * o.x); o.y = fmaxf(o.y, c.slope * o.y);
  o.z = fmaxf(o.z, c.slope * o.z); o.w = fmaxf(o.w, c.slope * o.w);
  return o;
}

template <bool W32, bool WHI, bool WLO>
__device__ __forceinline__ void coal_store_fast(const float4* __restrict__ stage, const uint32_t (&pixv)[8], int lane,
                                                int ch, int ld, float* __restrict__ out, __half* __restrict__ hi_p,
                                                __half* __restrict__ lo_p, const Chan4& cc, bool has_pre, bool has_r1,
                                                bool has_r2, const float4 (&r1v)[8], const float4 (&r2v)[8],
                                                float alpha1, float alpha2, int* __restrict__ status) {
  const int cidx = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int pl = it * 4 + (lane >> 3);
    if (pixv[it] != 0xffffffffu) {
      const uint32_t e1 = pixv[it] * (uint32_t)ld + ch;
      float4 o = stage[pl * 8 + (cidx ^ (pl & 7))];
      if (has_pre) {
        const float4 rr = r1v[it];
        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
      }
      o = chan_apply(o, cc);
      if (has_r1) {
        const float4 rr = r1v[it];
        o.x = fmaf(o.x, alpha1, rr.x); o.y = fmaf(o.y, alpha1, rr.y); o.z = fmaf(o.z, alpha1, rr.z); o.w = fmaf(o.w, alpha1, rr.w);
      }
      if (has_r2) {
        const float4 rr = r2v[it];
        o.x = fmaf(o.x, alpha2, rr.x); o.y = fmaf(o.y, alpha2, rr.y); o.z = fmaf(o.z, alpha2, rr.z); o.w = fmaf(o.w, alpha2, rr.w);
      }
      // st.global.cg: the consumers are other SMs (TMA / L2-coherent loads); a plain store through a pointer that was
      // loaded from memory compiles to a generic ST
      if (W32) __stcg(reinterpret_cast<float4*>(out + e1), o);
      if (WHI) {
        const uint2 hi = split_hi(o, status);
        __stcg(reinterpret_cast<uint2*>(hi_p + e1), hi);
        if (WLO) __stcg(reinterpret_cast<uint2*>(lo_p + e1), split_lo(o, hi));
      }
    }
  }
}

// Fused FlowStep inverse on one pixel (the thread's accumulator row): h (N <= 32 values, after bias / scale) is
// parked in the warp's staging region as scratch[c * 32 + lane] (conflict-free), z lives in registers.
// Arithmetic follows step_inverse_kernel (flow_ops.cu): AffineCouplings.py:73-87, Permutations.py:103-108,
// ActNorms.py:66-69.
constexpr int STEP_MAXC = 24;
constexpr int STEP_TAB_BYTES = (STEP_MAXC * STEP_MAXC + 2 * STEP_MAXC) * 4 + 64;   // W^-1 | scale | bias, per epilogue group
__device__ __forceinline__ void step_inverse_pixel(float* __restrict__ scratch, int lane, const float4 (&zq)[8],
                                                   float* __restrict__ zp, int C, int n_pass, bool has_w,
                                                   const float* __restrict__ s_w, const float* __restrict__ s_sc,
                                                   const float* __restrict__ s_b, __half* __restrict__ z16p,
                                                   __half* __restrict__ z16lo) {
  float z[STEP_MAXC];   // z arrives in the (otherwise unused) residual prefetch registers: 6 x float4
#pragma unroll
  for (int i = 0; i < STEP_MAXC / 4; ++i) {
    z[4 * i] = zq[i].x; z[4 * i + 1] = zq[i].y; z[4 * i + 2] = zq[i].z; z[4 * i + 3] = zq[i].w;
  }
#pragma unroll
  for (int i = 0; i < STEP_MAXC; ++i) {
    if (i >= n_pass && i < C) {
      const int j = i - n_pass;
      const float shift = scratch[(2 * j) * 32 + lane], scale = scratch[(2 * j + 1) * 32 + lane];
      z[i] = z[i] * expf(-coupling_logscale(scale)) - shift;
    }
  }
  for (int i = 0; i < C; ++i) {   // (scratch columns are private to the thread: no warp sync needed)
    float acc = 0.f;
    if (has_w) {
#pragma unroll
      for (int j = 0; j < STEP_MAXC; ++j)
        if (j < C) acc = fmaf(s_w[i * C + j], z[j], acc);
    } else {
#pragma unroll
      for (int j = 0; j < STEP_MAXC; ++j) acc = (j == i) ? z[j] : acc;
    }
    scratch[i * 32 + lane] = acc * s_sc[i] - s_b[i];
  }
  for (int i = 0; i < C; ++i) {
    const float v = scratch[i * 32 + lane];
    zp[i] = v;
    if (z16p && i < n_pass) {
      const __half hh = __float2half_rn(v);
      z16p[i] = hh;
      if (z16lo) z16lo[i] = __float2half_rn((v - __half2float(hh)) * 2048.0f);
    }
  }
}

// The same with the channel count as a template parameter (the nets use C = 6, 12, 21, 24): exact unrolling, the
// C x C mix with independent accumulators per row and the rows' weights as 16-byte broadcast loads when C % 4 == 0,
// no scratch round trip for the result.  (The generic version above spends most of its time in one dependent FMA
// chain per row and in predicated-off iterations: 0.7 ms of a 10.7 ms step.)
template <int C>
__device__ __forceinline__ void step_inverse_pixel_t(const float* __restrict__ scratch, int lane, const float4 (&zq)[8],
                                                     float* __restrict__ zp, int n_pass, bool has_w,
                                                     const float* __restrict__ s_w, const float* __restrict__ s_sc,
                                                     const float* __restrict__ s_b, __half* __restrict__ z16p,
                                                     __half* __restrict__ z16lo) {
  float z[C];
#pragma unroll
  for (int i = 0; i < C; ++i) {
    const float4 q4 = zq[i >> 2];
    z[i] = (i & 3) == 0 ? q4.x : ((i & 3) == 1 ? q4.y : ((i & 3) == 2 ? q4.z : q4.w));
  }
#pragma unroll
  for (int i = 0; i < C; ++i) {
    if (i >= n_pass) {
      const int j = i - n_pass;
      const float shift = scratch[(2 * j) * 32 + lane], scale = scratch[(2 * j + 1) * 32 + lane];
      z[i] = z[i] * expf(-coupling_logscale(scale)) - shift;
    }
  }
  float y[C];
  if (has_w) {
#pragma unroll
    for (int i = 0; i < C; ++i) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      if (C % 4 == 0) {
#pragma unroll
        for (int j = 0; j < C; j += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(s_w + i * C + j);
          a0 = fmaf(w4.x, z[j], a0); a1 = fmaf(w4.y, z[j + 1], a1);
          a2 = fmaf(w4.z, z[j + 2], a2); a3 = fmaf(w4.w, z[j + 3], a3);
        }
      } else {
#pragma unroll
        for (int j = 0; j < C; ++j) {
          const float w = s_w[i * C + j];
          if ((j & 3) == 0) a0 = fmaf(w, z[j], a0);
          else if ((j & 3) == 1) a1 = fmaf(w, z[j], a1);
          else if ((j & 3) == 2) a2 = fmaf(w, z[j], a2);
          else a3 = fmaf(w, z[j], a3);
        }
      }
      y[i] = (a0 + a1) + (a2 + a3);
    }
  } else {
#pragma unroll
    for (int i = 0; i < C; ++i) y[i] = z[i];
  }
#pragma unroll
  for (int i = 0; i < C; ++i) {
    const float v = y[i] * s_sc[i] - s_b[i];
    zp[i] = v;
    if (z16p && i < n_pass) {
      const __half hh = __float2half_rn(v);
      z16p[i] = hh;
      if (z16lo) z16lo[i] = __float2half_rn((v - __half2float(hh)) * 2048.0f);
    }
  }
}

// ------------------------------------------------------------------ kernel
// profile slots (HCF_TC_PROF=1): cycles summed over CTAs, printed by hcf_conv_tc_plan_destroy
enum { PROF_P_TOTAL = 0, PROF_P_DEPS, PROF_P_EMPTYA, PROF_P_EMPTYB, PROF_M_TOTAL, PROF_M_TMEM, PROF_M_FULLA, PROF_M_FULLB,
       PROF_M_CONVA, PROF_E_TOTAL, PROF_E_TMEMFULL, PROF_E_BODY, PROF_E_PUBLISH, PROF_E_LAYER, PROF_E_ROW, PROF_E_COAL, PROF_M_ISSUE, PROF_LAUNCHES, PROF_N };
// compiled in only with -DHCF_TC_PROF_BUILD (HCF_BUILD_PROF=1 python -m hcflow_b200.build --force): the accumulators
// cost ~30 registers per thread
#ifdef HCF_TC_PROF_BUILD
#define HCF_T(var) const long long var = prof_on ? clock64() : 0ll
#define HCF_ACC(slot, a, b) do { if (prof_on) pacc[slot] += (b) - (a); } while (0)
#define HCF_PROF_FLUSH(lo, hi) do { if (prof_on) for (int i_ = (lo); i_ <= (hi); ++i_) \
    atomicAdd((unsigned long long*)p.prof + i_, (unsigned long long)pacc[i_]); } while (0)
#else
#define HCF_T(var) do { } while (0)
#define HCF_ACC(slot, a, b) do { } while (0)
#define HCF_PROF_FLUSH(lo, hi) do { } while (0)
#endif
template <typename T>
__device__ __forceinline__ T* ldg_ptr(T* const* p) {
  return reinterpret_cast<T*>(__ldg(reinterpret_cast<const unsigned long long*>(p)));
}

// Dependency counters are polled with RELAXED loads (nine in flight at once) and ordered by ONE
// gpu-scope fence after the poll succeeds: nine ld.acquire in a row serialise into nine L2 round
// trips (~5k cycles per work item, measured), which is what used to bound the chained kernel.
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acquire_gpu() { asm volatile("fence.acquire.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// completed-layer counters of the 3x3 tile neighbourhood of (b, ty, tx); missing neighbours read as "done"
struct Deps { int v[9]; };
__device__ __forceinline__ void load_deps(Deps& d, const int* done, int base, int ty, int tx, int tiles_y, int tiles_x) {
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = ty + dy, xx = tx + dx;
      const bool in = yy >= 0 && yy < tiles_y && xx >= 0 && xx < tiles_x;
      d.v[(dy + 1) * 3 + dx + 1] = in ? ld_relaxed(done + base + yy * tiles_x + xx) : 0x7fffffff;
    }
}
__device__ __forceinline__ bool deps_ready(const Deps& d, int layer) {
  int m = d.v[0];
#pragma unroll
  for (int i = 1; i < 9; ++i) m = min(m, d.v[i]);
  return m >= layer;
}

// F16: operands are fp16 (hi / lo planes, 64 channels per 128-byte row, kind::f16); otherwise fp32 words read as TF32.
// STEP: the chain contains fused FlowStep layers (only those variants carry the per-pixel FlowStep code: its
// registers cost the encoder kernels 4 % when it was compiled into all of them).
// DIRECT: the variant carries the direct fp16-plane store path of the epilogue (chains with hi-plane-only layers, i.e.
// the RRDB encoder chains); the other chains keep the smaller binary -- code growth in one role costs every role.
template <int MT, int PASSES, int KS, bool F16, bool STEP = false, bool DIRECT = false>
__global__ void __launch_bounds__(F16 ? 384 : (PASSES == 3 ? 320 : 192), 1)
conv_tc_kernel(const __grid_constant__ Maps maps, const Params p) {
  constexpr int HALO = KS / 2;
  constexpr int HALO_W = halo_w(KS);
  constexpr int A_BYTES = a_bytes(MT, KS);
  constexpr int A_PART = a_part(MT, KS);
  constexpr int A_STAGE = A_PART * (PASSES == 3 ? 2 : 1);   // [raw | lo]  (fp16: [hi | lo], both loaded by TMA)
  constexpr int KCHX = F16 ? KCH16 : KCH;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if (p.debug & 16) return;   // timing experiment: launch cost only
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t slot_bytes = (uint32_t)p.slot_bytes;                        // B ring slot
  const uint32_t b_base = smem_base + p.sa * A_STAGE;
  const uint32_t bar_base = b_base + p.sb * slot_bytes;
  auto fullA = [&](int s) { return bar_base + 8u * s; };
  auto emptyA = [&](int s) { return bar_base + 8u * (p.sa + s); };
  auto convA = [&](int s) { return bar_base + 8u * (2 * p.sa + s); };
  auto fullB = [&](int s) { return bar_base + 8u * (3 * p.sa + s); };
  auto emptyB = [&](int s) { return bar_base + 8u * (3 * p.sa + p.sb + s); };
  const uint32_t tbar = bar_base + 8u * (3 * p.sa + 2 * p.sb);
  auto tmem_full = [&](int a) { return tbar + 8u * a; };
  auto tmem_empty = [&](int a) { return tbar + 16u + 8u * a; };
  // producer -> epilogue: number of this CTA's work items whose inputs have been acquired (monotonic counter; an
  // mbarrier would need the producer to be at most one phase ahead)
  const uint32_t dep_seq = tbar + 32u;
  // fp16 kernels: items published by the publisher warp of epilogue group g (tbar + 36 + 4 g)
  auto pub_seq = [&](int g) { return tbar + 36u + 4u * g; };
  const uint32_t tmem_slot = tbar + 48u;
  auto map_ptr = [&](int i) -> const CUtensorMap* { return &maps.m[i]; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t need_cols = 2u * MT * p.nb_max;
  const uint32_t tmem_cols =
      need_cols <= 32 ? 32u : (need_cols <= 64 ? 64u : (need_cols <= 128 ? 128u : (need_cols <= 256 ? 256u : 512u)));

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[0]) : "memory");
    for (int s = 0; s < p.sa; ++s) {
      mbar_init(fullA(s), 1);
      mbar_init(emptyA(s), 1);
      mbar_init(convA(s), 128);
    }
    for (int s = 0; s < p.sb; ++s) {
      mbar_init(fullB(s), 1);
      mbar_init(emptyB(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full(a), 1);
      mbar_init(tmem_empty(a), 128);
    }
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(dep_seq), "r"(0u) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(pub_seq(0)), "r"(0u) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(pub_seq(1)), "r"(0u) : "memory");
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

  const int per_img = p.tiles_x * p.tiles_y;
  const int n_items = (p.debug & 32) ? 0 : p.n_items;   // timing experiment: prologue + teardown only
  // q-th work item of this CTA (-1 = none)
  auto item_at = [&](int q) -> int {
    if (p.tpc > 0) {
      const int layer = q / p.tpc;
      if (layer >= p.n_layers || n_items == 0) return -1;
      return layer * p.n_tiles + (int)blockIdx.x + (q - layer * p.tpc) * (int)gridDim.x;
    }
    const int item = (int)blockIdx.x + q * (int)gridDim.x;
    return item < n_items ? item : -1;
  };
  const bool chain = p.done != nullptr;
#ifdef HCF_TC_PROF_BUILD
  const bool prof_on = p.prof != nullptr;
  long long pacc[PROF_N];
#pragma unroll
  for (int i = 0; i < PROF_N; ++i) pacc[i] = 0;
#endif

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t p_it = 0;
      int sA = 0, sB = 0;           // ring positions as (slot, phase) counters (see the MMA warp)
      uint32_t phA = 0, phB = 0;
      // per-layer fields stay in registers; a CTA sees the same layer for ~n_tiles/gridDim.x items in a row
      int cur_layer = -1, kchunks = 0, se0 = 0, se1 = 0, m0 = 0, m1 = 0, m2 = 0, slab_taps = 1, slabs = 1;
      int l0 = 0, l1 = 0, l2 = 0, lparts = 1, ltaps = KS * KS, lsplit = 0, co0 = 0, co1 = 0, co2 = 0;
      uint32_t tap_n = 0;   // bytes of one tap of B at one row block (N rows)
      const uint8_t* wimg = nullptr;
      Deps deps;
      HCF_T(tp0);
      if (chain && item_at(0) >= 0) {   // counters of the first item (prefetched one item ahead below)
        const int tile = item_at(0) % p.n_tiles;
        const int b = tile / per_img, r = tile % per_img;
        load_deps(deps, p.done, b * per_img, r / p.tiles_x, r % p.tiles_x, p.tiles_y, p.tiles_x);
      }
      for (int seq = 0, item; (item = item_at(seq)) >= 0; ++seq) {
        const int layer = item / p.n_tiles, tile = item - layer * p.n_tiles;
        const int b = tile / per_img, r = tile % per_img;
        const int ty = r / p.tiles_x, tx = r % p.tiles_x;
        const int y0 = ty * TH * MT, x0 = tx * TW;
        if (layer != cur_layer) {
          cur_layer = layer;
          const LayerDesc* L = p.layers + layer;
          kchunks = __ldg(&L->kchunks);
          se0 = __ldg(&L->seg_end[0]); se1 = __ldg(&L->seg_end[1]);
          m0 = __ldg(&L->map_idx[0]); m1 = __ldg(&L->map_idx[1]); m2 = __ldg(&L->map_idx[2]);
          l0 = __ldg(&L->map_lo[0]); l1 = __ldg(&L->map_lo[1]); l2 = __ldg(&L->map_lo[2]);
          co0 = __ldg(&L->seg_coff[0]); co1 = __ldg(&L->seg_coff[1]); co2 = __ldg(&L->seg_coff[2]);
          lparts = __ldg(&L->parts);
          lsplit = lparts == 2 ? __ldg(&L->split_kc) : 0;
          tap_n = (uint32_t)__ldg(&L->N) * ROW_BYTES;
          slab_taps = __ldg(&L->slab_taps);
          ltaps = __ldg(&L->taps);
          slabs = ltaps / slab_taps;
          wimg = reinterpret_cast<const uint8_t*>(ldg_ptr(&L->wimg));
        }
        auto issue_b = [&](int kc, int sl) {
          HCF_T(tb0);
          mbar_wait(emptyB(sB), phB ^ 1u);
          HCF_T(tb1);
          HCF_ACC(PROF_P_EMPTYB, tb0, tb1);
          if (p.debug & 4) {
            mbar_arrive(fullB(sB));
          } else {
            // weight image: chunks [0, lsplit) carry [hi ; lo] row blocks (2N rows per tap), the rest N rows
            const uint32_t tap_bytes = kc < lsplit ? 2u * tap_n : tap_n;
            const uint32_t b_slab = (uint32_t)slab_taps * tap_bytes;
            const size_t chunk_off = ((size_t)min(kc, lsplit) * 2u + (size_t)max(kc - lsplit, 0)) * (uint32_t)ltaps * tap_n;
            mbar_expect_tx(fullB(sB), b_slab);
            bulk_load(b_base + sB * slot_bytes, wimg + chunk_off + (size_t)sl * b_slab, b_slab, fullB(sB));
          }
          if (++sB == p.sb) { sB = 0; phB ^= 1u; }
        };
        // (issuing the first chunk's weight slabs BEFORE the dependency wait was measured slower: with a two-slot B
        //  ring the activation tile then queues behind the wait for a free slot)
        constexpr int n_pre = 0;
        if (chain) {
          if (layer > 0 && !(p.debug & 64)) {
            // wait until layer-1 is complete on the 3x3 tile neighbourhood (halo + WAR safety)
            HCF_T(td0);
            uint32_t spins = 0;
            while (!deps_ready(deps, layer)) {
              __nanosleep(32);
              load_deps(deps, p.done, b * per_img, ty, tx, p.tiles_y, p.tiles_x);
              // the launch is cooperative (all CTAs co-resident), so a wait is bounded by the chain's own run time;
              // ~2 s of polling means a producer CTA is missing or crashed: flag it and stop instead of hanging
              if (++spins > (1u << 24)) {
                if (p.status) atomicOr(p.status, STATUS_DEP_TIMEOUT);
                __trap();
              }
            }
            fence_acquire_gpu();                                      // acquire side of the epilogue's release
            asm volatile("fence.proxy.async.global;" ::: "memory");   // order the TMA (async proxy) reads after it
            HCF_T(td1);
            HCF_ACC(PROF_P_DEPS, td0, td1);
          }
          asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(dep_seq), "r"(p_it + 1u) : "memory");   // epilogue may prefetch
          const int nitem = item_at(seq + 1);                    // counters of the next item: in flight during this item's loads
          if (nitem >= 0) {
            const int nt = nitem % p.n_tiles;
            const int nb = nt / per_img, nr = nt % per_img;
            load_deps(deps, p.done, nb * per_img, nr / p.tiles_x, nr % p.tiles_x, p.tiles_y, p.tiles_x);
          }
        } else {
          asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(dep_seq), "r"(p_it + 1u) : "memory");
        }
        ++p_it;
        for (int kc = 0; kc < kchunks; ++kc) {
          HCF_T(ta0);
          mbar_wait(emptyA(sA), phA ^ 1u);
          HCF_T(ta1);
          HCF_ACC(PROF_P_EMPTYA, ta0, ta1);
          if (p.debug & 4) {
            mbar_arrive(fullA(sA));
          } else {
            const bool two = F16 && PASSES == 3 && kc < lsplit;   // split chunk of an fp16 chain: hi and lo planes
            mbar_expect_tx(fullA(sA), two ? 2 * A_BYTES : A_BYTES);
            const int mi = kc < se0 ? m0 : (kc < se1 ? m1 : m2);
            const int kl = kc < se0 ? kc : (kc < se1 ? kc - se0 : kc - se1);
            const int cch = kl * KCHX + (kc < se0 ? co0 : (kc < se1 ? co1 : co2));   // channel coordinate in the map
            tma_load_4d(smem_base + sA * A_STAGE, map_ptr(mi), fullA(sA), cch, x0 - HALO, y0 - HALO, b);
            if (two) {
              const int li = kc < se0 ? l0 : (kc < se1 ? l1 : l2);
              tma_load_4d(smem_base + sA * A_STAGE + A_PART, map_ptr(li), fullA(sA), cch, x0 - HALO, y0 - HALO, b);
            }
          }
          for (int sl = (kc == 0 ? n_pre : 0); sl < slabs; ++sl) issue_b(kc, sl);
          if (++sA == p.sa) { sA = 0; phA ^= 1u; }
        }
      }
      HCF_T(tp1);
      HCF_ACC(PROF_P_TOTAL, tp0, tp1);
      HCF_PROF_FLUSH(PROF_P_TOTAL, PROF_P_EMPTYB);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp follows the barriers; one elected lane issues (warp-uniform control flow
    // lets ptxas keep descriptors in uniform registers without a per-instruction election loop).
    // Descriptor templates: everything but the 14-bit start-address field (addr >> 4).  A shared
    // memory address is < 2^18, so adding (bytes >> 4) never carries out of the field; the
    // per-MMA work is one add per operand (every dependent ALU op of the single issuing thread
    // costs its full latency).
    const uint64_t a_tmpl = make_desc(0, HALO_W * ROW_BYTES);
    const uint64_t b_tmpl = make_desc(0, 8u * ROW_BYTES);
    uint32_t t_it = 0;
    // ring positions as (slot, phase) counters: `it % depth` / `it / depth` with a run-time depth are two integer
    // divisions (~60 instructions) per chunk and per slab on the one warp whose scalar latency idles the tensor pipe
    int sA = 0, sB = 0;
    uint32_t phA = 0, phB = 0;
    int cur_layer = -1, kchunks = 0, slab_taps = 1, slabs = 1, tap0 = 0;
    uint32_t parts = 1, nb = 0, nb_n = 0, idesc_n = 0, idesc = 0, n_cols = 0;
    int split_kc = 0, e0 = 0, e1 = 0, lk0 = 4, lk1 = 4, lk2 = 4;
    HCF_T(tm0);
    for (int seq = 0, item; (item = item_at(seq)) >= 0; ++seq, ++t_it) {
      const int layer = item / p.n_tiles;
      if (layer != cur_layer) {
        cur_layer = layer;
        const LayerDesc* L = p.layers + layer;
        kchunks = __ldg(&L->kchunks);
        const uint32_t N = (uint32_t)__ldg(&L->N);
        parts = (uint32_t)__ldg(&L->parts);
        const uint32_t NB = N * parts;
        slab_taps = __ldg(&L->slab_taps);
        slabs = __ldg(&L->taps) / slab_taps;
        tap0 = __ldg(&L->tap0);
        nb = NB * (ROW_BYTES >> 4);      // one tap of B in 16-byte units (split chunk)
        nb_n = N * (ROW_BYTES >> 4);     // ... of a one-pass chunk
        split_kc = parts == 2 ? __ldg(&L->split_kc) : 0;
        e0 = __ldg(&L->seg_end[0]); e1 = __ldg(&L->seg_end[1]);
        lk0 = __ldg(&L->seg_last_k[0]); lk1 = __ldg(&L->seg_last_k[1]); lk2 = __ldg(&L->seg_last_k[2]);
        const uint32_t fmt = F16 ? 0u : 2u;   // A / B format: F16 = 0, TF32 = 2; D = F32
        idesc_n = (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((128u >> 4) << 24);
        idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((NB >> 3) << 17) | ((128u >> 4) << 24);
        n_cols = N;
      }
      const uint32_t acc = t_it & 1u;
      HCF_T(tt0);
      mbar_wait(tmem_empty(acc), ((t_it >> 1) & 1u) ^ 1u);   // all lanes poll: measured faster than lane 0 + syncwarp
      tc_fence_after();
      HCF_T(tt1);
      HCF_ACC(PROF_M_TMEM, tt0, tt1);
      const uint32_t d0 = tmem_base + acc * MT * p.nb_max;
      uint32_t accum = 0u;   // first MMA of the item overwrites the accumulator
      for (int kc = 0; kc < kchunks; ++kc) {
        HCF_T(tfa0);
        mbar_wait(fullA(sA), phA);
        HCF_T(tfa1);
        if (PASSES == 3 && !F16) mbar_wait(convA(sA), phA);
        HCF_T(tfa2);
        HCF_ACC(PROF_M_FULLA, tfa0, tfa1);
        HCF_ACC(PROF_M_CONVA, tfa1, tfa2);
        const uint64_t a0 = a_tmpl + ((smem_base + sA * A_STAGE) >> 4);
        const bool split = PASSES == 3 && parts == 2 && kc < split_kc;   // this chunk: hi + lo on both operands
        // a segment's last chunk may be partly padding (96 = 64 + 32 channels, the 3..12 channels of z1): only the
        // K-steps that hold real channels are issued
        const int kmax = (kc == kchunks - 1) ? lk2 : ((kc == e0 - 1) ? lk0 : ((kc == e1 - 1) ? lk1 : 4));
        const uint32_t nb_kc = split ? nb : nb_n;
        const uint32_t idesc_kc = split ? idesc : idesc_n;
        for (int sl = 0; sl < slabs; ++sl) {
          HCF_T(tfb0);
          mbar_wait(fullB(sB), phB);
          tc_fence_after();
          HCF_T(tfb1);
          HCF_ACC(PROF_M_FULLB, tfb0, tfb1);
          HCF_T(tis0);
          if (elect_one()) {
            const uint64_t b0 = b_tmpl + ((b_base + sB * slot_bytes) >> 4);
            for (int t = 0; t < ((p.debug & 2) ? 0 : slab_taps); ++t) {
              const int tap = tap0 + sl * slab_taps + t;
              const int dy = tap / KS, dx = tap - dy * KS;
              uint64_t a_tap = a0 + (uint32_t)((dy * HALO_W + dx) * (ROW_BYTES >> 4));
              if (p.debug & 1) a_tap = make_desc(0, 8u * ROW_BYTES) + ((smem_base + sA * A_STAGE) >> 4);
              const uint64_t b_tap = b0 + (uint32_t)t * nb_kc;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k >= kmax) break;
                const uint64_t bd = b_tap + 2u * k;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  const uint32_t d = d0 + mt * p.nb_max;
                  const uint64_t ad = a_tap + (uint32_t)((mt * TH * HALO_W * ROW_BYTES + k * 32) >> 4);
                  if (F16) {
                    // cols [0,N) += A_hi x B_hi, cols [N,2N) += A_hi x B_lo' + A_lo' x B_hi (lo' = lo * 2048)
                    umma_f16(d, ad, bd, idesc_kc, accum);
                    if (split) umma_f16(d + n_cols, ad + (A_PART >> 4), bd, idesc_n, 1u);
                  } else {
                    umma_tf32(d, ad, bd, idesc_kc, accum);                               // A x [B ; B_lo]
                    if (split) umma_tf32(d, ad + (A_PART >> 4), bd, idesc_n, 1u);        // A_lo x B
                  }
                }
                accum = 1u;
              }
            }
            umma_commit(emptyB(sB));
            if (sl == slabs - 1) {
              umma_commit(emptyA(sA));
              if (kc == kchunks - 1) umma_commit(tmem_full(acc));
            }
          }
          __syncwarp();
          HCF_T(tis1);
          HCF_ACC(PROF_M_ISSUE, tis0, tis1);
          accum = 1u;
          if (++sB == p.sb) { sB = 0; phB ^= 1u; }
        }
        if (++sA == p.sa) { sA = 0; phA ^= 1u; }
      }
    }
    HCF_T(tm1);
    HCF_ACC(PROF_M_TOTAL, tm0, tm1);
    if (lane == 0) { HCF_PROF_FLUSH(PROF_M_TOTAL, PROF_M_CONVA); HCF_PROF_FLUSH(PROF_M_ISSUE, PROF_M_ISSUE); }
  } else if (warp < 6 || (F16 && warp < 10)) {
    // ===================== epilogue =====================
    // fp16 kernels run TWO epilogue groups of four warps (warps 2-5 and 6-9): group g drains accumulator buffer g,
    // i.e. every other work item, so the single-warp-per-scheduler latency of the store code is halved.
    constexpr int EG = F16 ? 2 : 1;
    const int grp = (warp - 2) >> 2;              // epilogue group of this warp
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int et = (threadIdx.x - 64) & 127;      // 0..127 within the group
    float* s_bias = reinterpret_cast<float*>(gen_base + (bar_base + BAR_BYTES - smem_base)) + grp * (256 + STAGE_BYTES / 4);
    float* s_scale = s_bias + 128;                // [128] bias | [128] scale | staging, per group
    float4* stage = reinterpret_cast<float4*>(s_bias + 256) + q * 256;   // this warp's [32 pixels][8 x 16 B]
    float* st_w = reinterpret_cast<float*>(gen_base + (bar_base + TAIL_BYTES - smem_base) + grp * STEP_TAB_BYTES);
    float* st_sc = st_w + STEP_MAXC * STEP_MAXC;   // (only present when p.step_tab)
    float* st_b = st_sc + STEP_MAXC;
    Out16 o16 = {nullptr, nullptr, nullptr, nullptr};
    uint32_t t_it = grp;
    // per-layer fields stay in registers, bias / scale in shared memory (every thread needs all N of them)
    int cur_layer = -1, N = 0, cout = 0, act = 0, out_vec = 0, parts = 1;
    int out_ld = 0, out2_ld = 0, res1_ld = 0, res2_ld = 0;
    bool has_bias = false, has_scale = false;
    int fast = 0;   // straight-line store path: bit0 fp32, bit1 hi plane, bit2 lo plane (0 = generic path)
    bool direct_hi = false;
    const bool getenv_direct = !(p.debug & 512);   // (HCF_TC_DEBUG=512 switches the direct path off: A/B timing)
    float* out = nullptr; float* out2 = nullptr;
    const float* res1 = nullptr; const float* res2 = nullptr;
    float alpha1 = 0.f, alpha2 = 0.f;
    bool is_pre = false;   // res1 holds the pre-activation addend (prefetched through the same registers)
    float* step_z = nullptr; int step_z_ld = 0, step_C = 0, step_npass = 0, step_z16_ld = 0;
    const float* step_w = nullptr; const float* step_sc = nullptr; const float* step_b = nullptr;
    __half* step_z16 = nullptr; __half* step_z16_lo = nullptr;
    float* raw2 = nullptr; int raw2_ld = 0;
    HCF_T(te0);
    for (int seq = grp, item; (item = item_at(seq)) >= 0; seq += EG, t_it += EG) {
      const int layer = item / p.n_tiles, tile = item - layer * p.n_tiles;
      const int b = tile / per_img, r = tile % per_img;
      const int y0 = (r / p.tiles_x) * TH * MT, x0 = (r % p.tiles_x) * TW;
      HCF_T(tl0);
      if (layer != cur_layer) {
        cur_layer = layer;
        const LayerDesc* L = p.layers + layer;
        N = __ldg(&L->N); cout = __ldg(&L->cout); act = __ldg(&L->act); out_vec = __ldg(&L->out_vec);
        parts = __ldg(&L->parts);
        const float* bias = ldg_ptr(&L->bias);
        const float* scale = ldg_ptr(&L->scale);
        has_bias = bias != nullptr; has_scale = scale != nullptr;
        out = ldg_ptr(&L->out); out2 = ldg_ptr(&L->out2);
        if (F16) {
          o16.hi = ldg_ptr(&L->out_hi); o16.lo = ldg_ptr(&L->out_lo);
          o16.hi2 = ldg_ptr(&L->out2_hi); o16.lo2 = ldg_ptr(&L->out2_lo);
        }
        res1 = ldg_ptr(&L->res1); res2 = ldg_ptr(&L->res2);
        out_ld = __ldg(&L->out_ld); out2_ld = __ldg(&L->out2_ld);
        res1_ld = __ldg(&L->res1_ld); res2_ld = __ldg(&L->res2_ld);
        alpha1 = __ldg(&L->alpha1); alpha2 = __ldg(&L->alpha2);
        step_z = STEP ? ldg_ptr(&L->step_z) : nullptr;
        if (STEP && step_z) {
          step_z_ld = __ldg(&L->step_z_ld); step_C = __ldg(&L->step_C); step_npass = __ldg(&L->step_npass);
          step_w = ldg_ptr(&L->step_w); step_sc = ldg_ptr(&L->step_sc); step_b = ldg_ptr(&L->step_b);
          step_z16 = ldg_ptr(&L->step_z16); step_z16_ld = __ldg(&L->step_z16_ld);
          step_z16_lo = ldg_ptr(&L->step_z16_lo);
        }
        raw2 = ldg_ptr(&L->raw2);
        if (raw2) { raw2_ld = __ldg(&L->raw2_ld); cout = 32; }   // the main path sees columns [0, 32) only
        is_pre = false;
        if (const float* pre = ldg_ptr(&L->pre)) {   // host guarantees: no res1 / res2 on such a layer
          res1 = pre; res1_ld = __ldg(&L->pre_ld); is_pre = true;
        }
        fast = 0;
        if (out_vec && cout % 32 == 0 && out2 == nullptr && o16.hi2 == nullptr && o16.lo2 == nullptr &&
            (o16.lo == nullptr || o16.hi != nullptr))
          fast = (out ? 1 : 0) | (o16.hi ? 2 : 0) | (o16.lo ? 4 : 0);
        // hi plane only, no residual / addend / raw partial sums, one pass, 16-byte aligned rows: direct stores
        direct_hi = DIRECT && F16 && fast == 2 && res1 == nullptr && res2 == nullptr && raw2 == nullptr && parts == 1 &&
                    !(STEP && step_z) && out_ld % 8 == 0 && getenv_direct;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");   // everyone is done with the previous layer's bias / scale
        s_bias[et] = __ldg(p.epi + (size_t)layer * 256 + et);          // inline table: no pointer chase
        s_scale[et] = __ldg(p.epi + (size_t)layer * 256 + 128 + et);
        if (STEP && step_z) {
          if (step_w)
            for (int i = et; i < step_C * step_C; i += 128) st_w[i] = __ldg(step_w + i);
          if (et < step_C) { st_sc[et] = __ldg(step_sc + et); st_b[et] = __ldg(step_b + et); }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      }
      const uint32_t acc = t_it & 1u;
      // residual tiles in the coalesced-domain mapping, prefetched one column group ahead (the first group before
      // the accumulator wait): L2-coherent loads cost a full round trip each when they are issued one by one
      const bool res_pf = out_vec && (res1 != nullptr || res2 != nullptr) && !(p.debug & 8);
      float4 r1v[8], r2v[8];
      uint32_t pixv[8];   // coalesced-domain pixel of iteration `it` (global pixel index, ~0 = outside the image)
      auto pix_setup = [&](int mt_) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int mm = q * 32 + it * 4 + (lane >> 3);
          const int gy = y0 + mt_ * TH + mm / TW, gx = x0 + mm % TW;
          pixv[it] = (gy < p.H && gx < p.W) ? (uint32_t)((b * p.H + gy) * p.W + gx) : 0xffffffffu;
        }
      };
      // a layer with ONE residual / addend tensor and two column groups (N = 64: an RDB's conv5 without the RRDB
      // residual, the sub-net's first conv) fetches the second group's tile into the idle r2v registers up front
      const bool ahead2 = res_pf && res2 == nullptr && N > 32 && N <= 64 && MT == 1;
      auto res_prefetch_2nd = [&]() {
        const int ch_ = 32 + (lane & 7) * 4;
        if (ch_ + 3 < cout) {
#pragma unroll
          for (int it = 0; it < 8; ++it)
            if (pixv[it] != 0xffffffffu)
              r2v[it] = __ldcg(reinterpret_cast<const float4*>(res1 + (pixv[it] * (uint32_t)res1_ld + ch_)));
        }
      };
      auto res_prefetch = [&](int c0_) {
        const int ch_ = c0_ + (lane & 7) * 4;
        if (ch_ + 3 < cout) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (pixv[it] != 0xffffffffu) {
              if (res1) r1v[it] = __ldcg(reinterpret_cast<const float4*>(res1 + (pixv[it] * (uint32_t)res1_ld + ch_)));
              if (res2) r2v[it] = __ldcg(reinterpret_cast<const float4*>(res2 + (pixv[it] * (uint32_t)res2_ld + ch_)));
            }
          }
        }
      };
      pix_setup(0);
      if (res_pf || (STEP && step_z)) {
        // the producer has acquired this item's inputs (dependency counters + fence): from here on residuals, the
        // pre-activation addend and z may be read -- long before the accumulator is ready, so the latency hides
        uint32_t seen;
        do {
          asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(seen) : "r"(dep_seq) : "memory");
        } while (seen <= t_it);
        if (res_pf) res_prefetch(0);
        if (ahead2) res_prefetch_2nd();
        if (STEP && step_z) {
          const int mm = q * 32 + lane;
          const int gy = y0 + mm / TW, gx = x0 + mm % TW;    // (fused steps run with MT == 1)
          const bool in = gy < p.H && gx < p.W;
          const float* zp = step_z + (size_t)((b * p.H + (in ? gy : 0)) * p.W + (in ? gx : 0)) * step_z_ld;
#pragma unroll
          for (int i = 0; i < STEP_MAXC / 4; ++i) {   // a fused-step layer has no residuals: z rides in r1v
            r1v[i].x = (in && 4 * i < step_C) ? __ldcg(zp + 4 * i) : 0.f;
            r1v[i].y = (in && 4 * i + 1 < step_C) ? __ldcg(zp + 4 * i + 1) : 0.f;
            r1v[i].z = (in && 4 * i + 2 < step_C) ? __ldcg(zp + 4 * i + 2) : 0.f;
            r1v[i].w = (in && 4 * i + 3 < step_C) ? __ldcg(zp + 4 * i + 3) : 0.f;
          }
        }
      }
      HCF_T(tl1);
      mbar_wait(tmem_full(acc), (t_it >> 1) & 1u);
      tc_fence_after();
      HCF_T(tl2);
      HCF_ACC(PROF_E_LAYER, tl0, tl1);
      HCF_ACC(PROF_E_TMEMFULL, tl1, tl2);
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          const int gw = min(32, N - c0);          // columns of this group (16 or 32)
          HCF_T(tr0);
          if (STEP && step_z) {
            // ---- fused FlowStep inverse (N <= 32, single group): h never leaves the SM
            float* scratch = reinterpret_cast<float*>(stage);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (h * 16 < gw) {
                float v[16];
                const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (acc * MT + mt) * p.nb_max + (uint32_t)(h * 16);
                tmem_ld16(tcol, v);
                if (PASSES == 3 && parts == 2) {   // split sub-net conv: main + correction columns
                  float lo[16];
                  tmem_ld16(tcol + (uint32_t)N, lo);
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[j] = F16 ? fmaf(lo[j], 1.0f / 2048.0f, v[j]) : v[j] + lo[j];
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  float t = v[j];
                  if (has_bias) t += s_bias[h * 16 + j];
                  if (has_scale) t *= s_scale[h * 16 + j];
                  scratch[(h * 16 + j) * 32 + lane] = t;
                }
              }
            }
            // the accumulator has been read (tcgen05.wait::ld inside tmem_ld16): hand the TMEM buffer back NOW, the
            // per-pixel FlowStep arithmetic below runs while the MMA warp already fills it with a later item
            tc_fence_before();
            mbar_arrive(tmem_empty(acc));
            const int mm = q * 32 + lane;
            const int gy = y0 + mt * TH + mm / TW, gx = x0 + mm % TW;
            if (gy < p.H && gx < p.W && !(p.debug & (8 | 128))) {   // 128: timing experiment, no FlowStep arithmetic
              const uint32_t pix = (uint32_t)((b * p.H + gy) * p.W + gx);
              float* zp = step_z + (size_t)pix * step_z_ld;
              __half* z16p = step_z16 ? step_z16 + (size_t)pix * step_z16_ld : nullptr;
              __half* z16lo = (step_z16 && step_z16_lo) ? step_z16_lo + (size_t)pix * step_z16_ld : nullptr;
              const bool hw = step_w != nullptr;
              switch (step_C) {
                case 6: step_inverse_pixel_t<6>(scratch, lane, r1v, zp, step_npass, hw, st_w, st_sc, st_b, z16p, z16lo); break;
                case 12: step_inverse_pixel_t<12>(scratch, lane, r1v, zp, step_npass, hw, st_w, st_sc, st_b, z16p, z16lo); break;
                case 21: step_inverse_pixel_t<21>(scratch, lane, r1v, zp, step_npass, hw, st_w, st_sc, st_b, z16p, z16lo); break;
                case 24: step_inverse_pixel_t<24>(scratch, lane, r1v, zp, step_npass, hw, st_w, st_sc, st_b, z16p, z16lo); break;
                default: step_inverse_pixel(scratch, lane, r1v, zp, step_C, step_npass, hw, st_w, st_sc, st_b, z16p, z16lo); break;
              }
            }
            __syncwarp();
            continue;
          }
          if (DIRECT && F16 && direct_hi && !(p.debug & 8)) {
            // ---- hi-plane-only layers without residuals (an RDB's growth convs: four of five items): no staging
            // transpose -- the thread keeps its pixel, applies bias / scale / activation to its 32 accumulator columns and
            // writes the pixel's 64 contiguous bytes of the hi plane with four 16-byte stores (both 32-byte sectors are
            // completed by the same thread).  ~170 instead of ~860 warp-instructions per item and no shared-memory
            // traffic for the transpose.
            const int mm = q * 32 + lane;
            const int gy = y0 + mt * TH + mm / TW, gx = x0 + mm % TW;
            const bool in = gy < p.H && gx < p.W;
            __half* dst = o16.hi + ((size_t)((b * p.H + (in ? gy : 0)) * p.W + (in ? gx : 0)) * (size_t)out_ld + c0);
            const float slope = act == HCF_ACT_RELU ? 0.f : (act == HCF_ACT_LRELU ? 0.2f : 1.f);
            float amax = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float v[16];
              const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (acc * MT + mt) * p.nb_max + (uint32_t)(c0 + h * 16);
              tmem_ld16(tcol, v);
              if (h == 1 && mt == MT - 1 && c0 + 32 >= N) {   // last TMEM read of the item
                tc_fence_before();
                mbar_arrive(tmem_empty(acc));
              }
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                float a0 = (v[j] + s_bias[c0 + h * 16 + j]) * s_scale[c0 + h * 16 + j];
                float a1 = (v[j + 1] + s_bias[c0 + h * 16 + j + 1]) * s_scale[c0 + h * 16 + j + 1];
                a0 = fmaxf(a0, slope * a0); a1 = fmaxf(a1, slope * a1);
                amax = fmaxf(amax, fmaxf(fabsf(a0), fabsf(a1)));
                if (!(fabsf(a0) <= 65504.0f)) a0 = fminf(fmaxf(a0, -65504.0f), 65504.0f);
                if (!(fabsf(a1) <= 65504.0f)) a1 = fminf(fmaxf(a1, -65504.0f), 65504.0f);
                const __half2 hh = __floats2half2_rn(a0, a1);
                pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
              }
              if (in) {
                __stcg(reinterpret_cast<uint4*>(dst + h * 16), make_uint4(pk[0], pk[1], pk[2], pk[3]));
                __stcg(reinterpret_cast<uint4*>(dst + h * 16 + 8), make_uint4(pk[4], pk[5], pk[6], pk[7]));
              }
            }
            if (in && !(amax <= 65504.0f) && p.status) atomicOr(p.status, STATUS_F16_OVERFLOW);
            continue;
          }
          // ---- row domain (thread = pixel): TMEM accumulator -> staging (transpose through shared memory)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h * 16 < gw) {
              float v[16];
              const int cc = c0 + h * 16;
              const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (acc * MT + mt) * p.nb_max + (uint32_t)cc;
              tmem_ld16(tcol, v);
              if (PASSES == 3 && parts == 2) {
                float lo[16];
                tmem_ld16(tcol + (uint32_t)N, lo);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = F16 ? fmaf(lo[j], 1.0f / 2048.0f, v[j]) : v[j] + lo[j];
              }
#pragma unroll
              for (int j = 0; j < 4; ++j)   // 16-byte chunk index XOR (pixel & 7): conflict-free both ways
                stage[lane * 8 + ((h * 4 + j) ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
          __syncwarp();
          if (mt == MT - 1 && c0 + 32 >= N) {
            // last TMEM read of the item done: release the accumulator buffer before the (long) store phase
            tc_fence_before();
            mbar_arrive(tmem_empty(acc));
          }
          HCF_T(tr1);
          HCF_ACC(PROF_E_ROW, tr0, tr1);
          // ---- coalesced domain (8 lanes = the 32 channels of one pixel, 4 pixels per instruction):
          //      residuals -> full-line global stores (fp32 and / or fp16 hi / lo planes)
          if (raw2 && c0 >= 32 && !(p.debug & 8)) {
            // partial sum of a LATER conv over the inputs it shares with this one (an RDB's conv2 / conv4 over the
            // channels conv1 / conv3 read): stored raw, that conv adds it before its bias (hcf_conv_args.pre)
            const int chr = c0 - 32 + (lane & 7) * 4;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int pl = it * 4 + (lane >> 3);
              if (pixv[it] != 0xffffffffu)
                __stcg(reinterpret_cast<float4*>(raw2 + (pixv[it] * (uint32_t)raw2_ld + chr)), stage[pl * 8 + ((lane & 7) ^ (pl & 7))]);
            }
          } else if (!(p.debug & 8)) {
            const int cidx = lane & 7;
            const int ch = c0 + cidx * 4;
            const bool ch_ok = ch < cout && cidx * 4 < gw;
            const bool vec = out_vec && ch + 3 < cout;
            const bool hr1 = res1 != nullptr && !is_pre, hr2 = res2 != nullptr;
            Chan4 cc;   // bias / scale are padded to N entries; ch < N always
            cc.bias = *reinterpret_cast<const float4*>(s_bias + ch);
            cc.scale = *reinterpret_cast<const float4*>(s_scale + ch);
            cc.slope = act == HCF_ACT_RELU ? 0.f : (act == HCF_ACT_LRELU ? 0.2f : 1.f);
#define HCF_COAL(A, B_, C_) coal_store_fast<A, B_, C_>(stage, pixv, lane, ch, out_ld, out, o16.hi, o16.lo, cc, is_pre, hr1, hr2, \
                                                     r1v, r2v, alpha1, alpha2, p.status)
            switch (fast) {
              case 1: HCF_COAL(true, false, false); break;
              case 2: HCF_COAL(false, true, false); break;
              case 3: HCF_COAL(true, true, false); break;
              case 6: HCF_COAL(false, true, true); break;
              case 7: HCF_COAL(true, true, true); break;
              default: break;
            }
#undef HCF_COAL
#pragma unroll
            for (int it = 0; it < (fast ? 0 : 8); ++it) {
              const int pl = it * 4 + (lane >> 3);
              if (pixv[it] != 0xffffffffu && ch_ok) {
                const uint32_t e1 = pixv[it] * (uint32_t)out_ld + ch;     // element offsets fit 32 bits (checked on the host)
                const uint32_t e2 = pixv[it] * (uint32_t)out2_ld + ch;
                float4 o = stage[pl * 8 + (cidx ^ (pl & 7))];
                if (vec) {
                  if (is_pre) {
                    const float4 rr = r1v[it];
                    o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                  }
                  o = chan_apply(o, cc);
                  if (hr1) {
                    const float4 rr = r1v[it];
                    o.x = o.x * alpha1 + rr.x; o.y = o.y * alpha1 + rr.y; o.z = o.z * alpha1 + rr.z; o.w = o.w * alpha1 + rr.w;
                  }
                  if (res2) {
                    const float4 rr = r2v[it];
                    o.x = o.x * alpha2 + rr.x; o.y = o.y * alpha2 + rr.y; o.z = o.z * alpha2 + rr.z; o.w = o.w * alpha2 + rr.w;
                  }
                  if (out) *reinterpret_cast<float4*>(out + e1) = o;
                  if (out2) *reinterpret_cast<float4*>(out2 + e2) = o;
                  if (F16) {
                    if (o16.hi || o16.hi2) {
                      const uint2 hi = split_hi(o, p.status);
                      if (o16.hi) *reinterpret_cast<uint2*>(o16.hi + e1) = hi;
                      if (o16.hi2) *reinterpret_cast<uint2*>(o16.hi2 + e2) = hi;
                      if (o16.lo || o16.lo2) {
                        const uint2 lo = split_lo(o, hi);
                        if (o16.lo) *reinterpret_cast<uint2*>(o16.lo + e1) = lo;
                        if (o16.lo2) *reinterpret_cast<uint2*>(o16.lo2 + e2) = lo;
                      }
                    }
                  }
                } else {
                  // residuals may have been written by another SM earlier in this launch: L2-coherent loads
                  const float e4[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    if (ch + e < cout) {
                      float t = e4[e];
                      if (is_pre) t += __ldcg(res1 + (pixv[it] * (uint32_t)res1_ld + ch + e));
                      if (has_bias) t += s_bias[ch + e];
                      if (has_scale) t *= s_scale[ch + e];
                      t = act == HCF_ACT_RELU ? fmaxf(t, 0.f) : (act == HCF_ACT_LRELU ? (t > 0.f ? t : 0.2f * t) : t);
                      if (hr1) t = t * alpha1 + __ldcg(res1 + (pixv[it] * (uint32_t)res1_ld + ch + e));
                      if (res2) t = t * alpha2 + __ldcg(res2 + (pixv[it] * (uint32_t)res2_ld + ch + e));
                      if (out) out[e1 + e] = t;
                      if (out2) out2[e2 + e] = t;
                      if (F16) {
                        if (!(fabsf(t) <= 65504.0f)) {   // range guard (see split_hi)
                          if (p.status) atomicOr(p.status, STATUS_F16_OVERFLOW);
                          t = fminf(fmaxf(t, -65504.0f), 65504.0f);
                        }
                        const __half hh = __float2half_rn(t);
                        const __half hl = __float2half_rn((t - __half2float(hh)) * 2048.0f);
                        if (o16.hi) o16.hi[e1 + e] = hh;
                        if (o16.hi2) o16.hi2[e2 + e] = hh;
                        if (o16.lo) o16.lo[e1 + e] = hl;
                        if (o16.lo2) o16.lo2[e2 + e] = hl;
                      }
                    }
                  }
                }
              }
            }
            if (res_pf) {   // next column group's residuals: in flight during its row phase
              if (ahead2) {
#pragma unroll
                for (int it = 0; it < 8; ++it) r1v[it] = r2v[it];
              } else if (c0 + 32 < N) {
                res_prefetch(c0 + 32);
              } else if (mt + 1 < MT) {
                pix_setup(mt + 1);
                res_prefetch(0);
              }
            } else if (c0 + 32 >= N && mt + 1 < MT) {
              pix_setup(mt + 1);
            }
          }
          __syncwarp();   // staging is reused by the next column group
          HCF_T(tr2);
          HCF_ACC(PROF_E_COAL, tr1, tr2);
        }
      }
      HCF_T(tl3);
      HCF_ACC(PROF_E_BODY, tl2, tl3);
      if (chain) {
        // publish: 128-thread barrier (the stores of every epilogue thread happen before it), then ONE thread
        // makes them visible at gpu scope and bumps the tile's counter (release side of the producer's acquire)
        if (F16) {
          // hand the release to the group's publisher warp (the gpu-scope MEMBAR costs ~1.5k cycles): wait until it
          // has published this group's previous item (so that the barrier is at most one phase ahead), then arrive
          // without blocking
          const uint32_t own = t_it >> 1;   // index of this item among the group's items
          uint32_t seen;
          do {
            asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(seen) : "r"(pub_seq(grp)) : "memory");
          } while (seen < own);
          asm volatile("bar.arrive %0, 160;" ::"r"(3 + grp) : "memory");
        } else {
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          if (et == 0) red_release_add(p.done + tile, 1);
        }
      }
      HCF_T(tl4);
      HCF_ACC(PROF_E_PUBLISH, tl3, tl4);
    }
    HCF_T(te1);
    HCF_ACC(PROF_E_TOTAL, te0, te1);
    if (et == 0 && grp == 0) {
      HCF_PROF_FLUSH(PROF_E_TOTAL, PROF_E_COAL);
#ifdef HCF_TC_PROF_BUILD
      if (prof_on && blockIdx.x == 0) atomicAdd((unsigned long long*)p.prof + PROF_LAUNCHES, 1ull);
#endif
    }
  } else if (F16) {
    // ===================== publishers (fp16 kernels, warps 10 and 11) =====================
    // one warp per epilogue group: joins the group's 128 threads on a named barrier (their stores happen before it),
    // then ONE lane makes them visible at gpu scope and bumps the tile's counter (release side of the producer's
    // acquire) -- off the epilogue's critical path
    if (chain) {
      const int grp = warp - 10;
      uint32_t own = 0;
      for (int seq = grp, item; (item = item_at(seq)) >= 0; seq += 2, ++own) {
        const int tile = item % p.n_tiles;
        asm volatile("bar.sync %0, 160;" ::"r"(3 + grp) : "memory");
        if (lane == 0) {
          red_release_add(p.done + tile, 1);
          asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(pub_seq(grp)), "r"(own + 1u) : "memory");
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== A_lo converters (PASSES == 3) =====================
    if (PASSES == 3 && !F16) {
      const int et = threadIdx.x - 192;   // 0..127
      uint32_t a_it = 0;
      for (int seq = 0, item; (item = item_at(seq)) >= 0; ++seq) {
        const LayerDesc* L = p.layers + item / p.n_tiles;
        const int kchunks = __ldg(&L->kchunks);
        const int n4 = __ldg(&L->parts) == 2 ? A_BYTES / 16 : 0;   // one-pass layers need no A_lo
        for (int kc = 0; kc < kchunks; ++kc, ++a_it) {
          const int sA = a_it % p.sa;
          mbar_wait(fullA(sA), (a_it / p.sa) & 1u);
          const float4* src = reinterpret_cast<const float4*>(gen_base + (size_t)sA * A_STAGE);
          float4* dst = reinterpret_cast<float4*>(gen_base + (size_t)sA * A_STAGE + A_PART);
          for (int i = et; i < n4; i += 128) {
            float4 v = src[i];
            v.x -= __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
            v.y -= __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
            v.z -= __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
            v.w -= __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
            dst[i] = v;
          }
          fence_async_smem();
          mbar_arrive(convA(sA));
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

// ------------------------------------------------------------------ host helpers shared by conv_tc.cu / flowstep_tc.cu
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
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

inline int n_for(int cout) { return (cout + 15) / 16 * 16; }

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = NUM_SMS_FALLBACK;
  }
  return n;
}

}  // namespace tc
}  // namespace hcf
