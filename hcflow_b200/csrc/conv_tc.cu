// 3x3 convolution as a persistent implicit GEMM on the 5th-generation tensor cores
// (tcgen05 / TMEM / TMA, sm_100a).
//
//   D[128 pixels, N] += A[128 pixels, 32 ch] * B[N, 32 ch]^T     per (tap, 32-channel chunk)
//
// Work item = MT vertically adjacent 16x8 pixel tiles of one image (UMMA M = 128 each) x all
// N = ceil16(Cout) <= 128 output channels.  CTAs are persistent (grid = #SMs, items strided by
// gridDim.x) and warp-specialised.  A launch executes either one convolution or a CHAIN of
// dependent convolutions on the same pixel grid (e.g. the 211 convs of an RRDB encoder level):
// item = (layer, tile) in layer-major order; a tile of layer l may start once the 3x3 tile
// neighbourhood of layer l-1 is complete, tracked by per-tile counters in global memory
// (release by the epilogue, acquire by the producer), so there is no kernel boundary, no
// ramp-up / drain and no wave quantisation between the convs.  Roles:
//   warp 0      TMA producer.  Per 32-channel chunk ONE 4-D TMA tile load brings the
//               (16*MT+2) x 10 halo tile [rows][10][32 ch] (128-byte swizzle; out-of-bounds -> 0
//               is exactly the conv's zero padding) into the A ring; the pre-swizzled weights
//               stream through a second ring in (chunk, dy) slabs [3 taps][N][32 ch].
//   warp 1      TMEM owner + MMA issuer: 9 taps x 4 K-steps of tcgen05.mma.kind::tf32 per chunk
//               and sub-tile.  The nine taps do NOT reload activations: tap (dy,dx) is a smem
//               descriptor whose start is shifted by (dy*10+dx) 128-byte rows into the same halo
//               tile and whose 8-row-group stride (SBO) is the halo row pitch (1280 B).  The
//               128B-swizzle XOR is a function of the absolute smem address, so a shifted view
//               of a TMA-written tile stays consistent (verified on B200).
//   warps 2..5  epilogue: tcgen05.ld -> bias / scale / activation / residuals -> global.  The
//               accumulator is double-buffered in TMEM (2 x MT x N columns), so the epilogue of
//               item i overlaps the loads and MMAs of item i+1.
//   warps 6..9  (PASSES == 3 only) 3xTF32 split: the tensor core reads an fp32 word as TF32 by
//               ignoring the low 13 mantissa bits, so the "hi" parts are free; these warps form
//               A_lo = a - trunc(a) next to every A stage and B_lo = w - trunc(w) is precomputed on
//               the host.  The three products take TWO MMAs per K-step: A x [B ; B_lo] as one
//               2N-row operand (columns [0,N) and [N,2N) of the accumulator) and A_lo x B into
//               columns [0,N); the epilogue adds the two column groups.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace hcf {
namespace tc {

constexpr int TH = 16, TW = 8;               // sub-tile (UMMA M = 128)
constexpr int KCH = 32;                      // channels per K chunk (= 128 B rows)
constexpr int ROW_BYTES = KCH * 4;           // 128
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int NUM_SMS_FALLBACK = 148;

// ks = 3: (16*mt+2) x 10 halo tile; ks = 1: plain 16*mt x 8 tile
__host__ __device__ constexpr int halo_w(int ks) { return TW + (ks - 1); }
__host__ __device__ constexpr int halo_rows(int mt, int ks) { return TH * mt + (ks - 1); }
__host__ __device__ constexpr int a_bytes(int mt, int ks) { return halo_rows(mt, ks) * halo_w(ks) * ROW_BYTES; }
__host__ __device__ constexpr int a_part(int mt, int ks) { return (a_bytes(mt, ks) + 1023) / 1024 * 1024; }

constexpr int MAX_MAPS = 8;

// One convolution of a chain.  Lives in global memory; every warp role reads the fields it needs
// at the start of a work item.
struct LayerDesc {
  int nseg;
  int map_idx[3];     // tensor map of each segment
  int seg_end[3];     // chunk index where segment i ends (prefix sums)
  int kchunks;        // total 32-channel chunks
  int N;              // UMMA N (multiple of 16, <= 128)
  int cout;
  int act;
  int out_vec;
  int parts;          // 1: one TF32 pass (B rows = N);  2: 3xTF32 split (B rows = [raw N ; lo N], plus A_lo x B)
  int slab_taps;      // taps of this layer per B ring slot: KS*KS, KS or 1 (largest that fits the slot)
  const float* wimg;  // [kchunks][ks dy][ks dx][NB][32] pre-swizzled; NB = N * parts
  const float* bias;
  const float* scale;
  float* out; int out_ld;
  float* out2; int out2_ld;
  const float* res1; int res1_ld; float alpha1;
  const float* res2; int res2_ld; float alpha2;
};

struct Params {
  int B, H, W;
  int tiles_x, tiles_y;
  int n_tiles;        // B * tiles_x * tiles_y  (work items per layer)
  int n_layers;       // > 1: a chain of dependent convs executed by ONE persistent launch
  int n_items;        // n_layers * n_tiles
  int nb_max;         // max over layers of B rows per tap (N, or 2N in 3-pass mode)
  int sa, sb;         // ring depths
  int slot_bytes;     // bytes of one B ring slot
  int debug;          // timing experiments only (HCF_TC_DEBUG, wrong results): 1 aligned A descriptors,
                      // 2 no MMAs, 4 no loads, 8 no epilogue stores, 16 launch only, 32 prologue only
  const LayerDesc* layers;
  int* done;          // chain mode: per-tile count of completed layers (zeroed before the launch)
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

// ------------------------------------------------------------------ kernel
template <typename T>
__device__ __forceinline__ T* ldg_ptr(T* const* p) {
  return reinterpret_cast<T*>(__ldg(reinterpret_cast<const unsigned long long*>(p)));
}

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int MT, int PASSES, int KS>
__global__ void __launch_bounds__(PASSES == 3 ? 320 : 192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap maps0, const __grid_constant__ CUtensorMap maps1,
               const __grid_constant__ CUtensorMap maps2, const __grid_constant__ CUtensorMap maps3,
               const __grid_constant__ CUtensorMap maps4, const __grid_constant__ CUtensorMap maps5,
               const __grid_constant__ CUtensorMap maps6, const __grid_constant__ CUtensorMap maps7,
               const Params p) {
  constexpr int HALO = KS / 2;
  constexpr int HALO_W = halo_w(KS);
  constexpr int A_BYTES = a_bytes(MT, KS);
  constexpr int A_PART = a_part(MT, KS);
  constexpr int A_STAGE = A_PART * (PASSES == 3 ? 2 : 1);   // [raw | lo]
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
  const uint32_t tmem_slot = tbar + 32u;
  auto map_ptr = [&](int i) -> const CUtensorMap* {
    switch (i) {
      case 0: return &maps0; case 1: return &maps1; case 2: return &maps2; case 3: return &maps3;
      case 4: return &maps4; case 5: return &maps5; case 6: return &maps6; default: return &maps7;
    }
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t need_cols = 2u * MT * p.nb_max;
  const uint32_t tmem_cols =
      need_cols <= 32 ? 32u : (need_cols <= 64 ? 64u : (need_cols <= 128 ? 128u : (need_cols <= 256 ? 256u : 512u)));

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps0) : "memory");
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
  const bool chain = p.done != nullptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t a_it = 0, b_it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int layer = item / p.n_tiles, tile = item - layer * p.n_tiles;
        const LayerDesc* L = p.layers + layer;
        const int b = tile / per_img, r = tile % per_img;
        const int ty = r / p.tiles_x, tx = r % p.tiles_x;
        const int y0 = ty * TH * MT, x0 = tx * TW;
        const int kchunks = __ldg(&L->kchunks);
        const int se0 = __ldg(&L->seg_end[0]), se1 = __ldg(&L->seg_end[1]);
        const int m0 = __ldg(&L->map_idx[0]), m1 = __ldg(&L->map_idx[1]), m2 = __ldg(&L->map_idx[2]);
        const uint32_t tap_bytes = (uint32_t)(__ldg(&L->N) * __ldg(&L->parts)) * ROW_BYTES;
        const int slab_taps = __ldg(&L->slab_taps);
        const int slabs = (KS * KS) / slab_taps;
        const uint32_t b_slab = (uint32_t)slab_taps * tap_bytes;
        const uint8_t* wimg = reinterpret_cast<const uint8_t*>(ldg_ptr(&L->wimg));
        if (chain && layer > 0) {
          // wait until layer-1 is complete on the 3x3 tile neighbourhood (halo + WAR safety)
          const int base = b * per_img;
          for (;;) {
            int ok = 1;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
              for (int dx = -1; dx <= 1; ++dx) {
                const int yy = ty + dy, xx = tx + dx;
                if (yy >= 0 && yy < p.tiles_y && xx >= 0 && xx < p.tiles_x)
                  ok &= (ld_acquire(p.done + base + yy * p.tiles_x + xx) >= layer) ? 1 : 0;
              }
            if (ok) break;
            __nanosleep(64);
          }
          asm volatile("fence.proxy.async;" ::: "memory");   // order the TMA reads after the acquire
        }
        for (int kc = 0; kc < kchunks; ++kc) {
          const int sA = a_it % p.sa;
          mbar_wait(emptyA(sA), ((a_it / p.sa) & 1u) ^ 1u);
          if (p.debug & 4) {
            mbar_arrive(fullA(sA));
          } else {
            mbar_expect_tx(fullA(sA), A_BYTES);
            const int mi = kc < se0 ? m0 : (kc < se1 ? m1 : m2);
            const int kl = kc < se0 ? kc : (kc < se1 ? kc - se0 : kc - se1);
            tma_load_4d(smem_base + sA * A_STAGE, map_ptr(mi), fullA(sA), kl * KCH, x0 - HALO, y0 - HALO, b);
          }
          ++a_it;
          for (int sl = 0; sl < slabs; ++sl) {
            const int sB = b_it % p.sb;
            mbar_wait(emptyB(sB), ((b_it / p.sb) & 1u) ^ 1u);
            if (p.debug & 4) {
              mbar_arrive(fullB(sB));
            } else {
              mbar_expect_tx(fullB(sB), b_slab);
              bulk_load(b_base + sB * slot_bytes,
                        wimg + (size_t)kc * ((uint32_t)(KS * KS) * tap_bytes) + (size_t)sl * b_slab, b_slab, fullB(sB));
            }
            ++b_it;
          }
        }
      }
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
    uint32_t a_it = 0, b_it = 0, t_it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++t_it) {
      const int layer = item / p.n_tiles;
      const LayerDesc* L = p.layers + layer;
      const int kchunks = __ldg(&L->kchunks);
      const uint32_t N = (uint32_t)__ldg(&L->N);
      const uint32_t parts = (uint32_t)__ldg(&L->parts);
      const uint32_t NB = N * parts;
      const int slab_taps = __ldg(&L->slab_taps);
      const int slabs = (KS * KS) / slab_taps;
      const uint32_t nb = NB * (ROW_BYTES >> 4);      // one tap of B in 16-byte units
      const uint32_t idesc_n = (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((NB >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t acc = t_it & 1u;
      mbar_wait(tmem_empty(acc), ((t_it >> 1) & 1u) ^ 1u);   // all lanes poll: measured faster than lane 0 + syncwarp
      tc_fence_after();
      const uint32_t d0 = tmem_base + acc * MT * p.nb_max;
      uint32_t accum = 0u;   // first MMA of the item overwrites the accumulator
      for (int kc = 0; kc < kchunks; ++kc) {
        const int sA = a_it % p.sa;
        const uint32_t phA = (a_it / p.sa) & 1u;
        mbar_wait(fullA(sA), phA);
        if (PASSES == 3) mbar_wait(convA(sA), phA);
        const uint64_t a0 = a_tmpl + ((smem_base + sA * A_STAGE) >> 4);
        for (int sl = 0; sl < slabs; ++sl) {
          const int sB = b_it % p.sb;
          mbar_wait(fullB(sB), (b_it / p.sb) & 1u);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t b0 = b_tmpl + ((b_base + sB * slot_bytes) >> 4);
            for (int t = 0; t < ((p.debug & 2) ? 0 : slab_taps); ++t) {
              const int tap = sl * slab_taps + t;
              const int dy = tap / KS, dx = tap - dy * KS;
              uint64_t a_tap = a0 + (uint32_t)((dy * HALO_W + dx) * (ROW_BYTES >> 4));
              if (p.debug & 1) a_tap = make_desc(0, 8u * ROW_BYTES) + ((smem_base + sA * A_STAGE) >> 4);
              const uint64_t b_tap = b0 + (uint32_t)t * nb;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t bd = b_tap + 2u * k;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  const uint32_t d = d0 + mt * p.nb_max;
                  const uint64_t ad = a_tap + (uint32_t)((mt * TH * HALO_W * ROW_BYTES + k * 32) >> 4);
                  umma_tf32(d, ad, bd, idesc, accum);                                  // A x [B ; B_lo]
                  if (PASSES == 3 && parts == 2) umma_tf32(d, ad + (A_PART >> 4), bd, idesc_n, 1u);  // A_lo x B
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
          accum = 1u;
          ++b_it;
        }
        ++a_it;
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;                  // accumulator row = pixel within the sub-tile
    uint32_t t_it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++t_it) {
      const int layer = item / p.n_tiles, tile = item - layer * p.n_tiles;
      const LayerDesc* L = p.layers + layer;
      const int b = tile / per_img, r = tile % per_img;
      const int y0 = (r / p.tiles_x) * TH * MT, x0 = (r % p.tiles_x) * TW;
      const int N = __ldg(&L->N), cout = __ldg(&L->cout), act = __ldg(&L->act), out_vec = __ldg(&L->out_vec);
      const int parts = __ldg(&L->parts);
      const float* bias = ldg_ptr(&L->bias);
      const float* scale = ldg_ptr(&L->scale);
      float* out = ldg_ptr(&L->out);
      float* out2 = ldg_ptr(&L->out2);
      const float* res1 = ldg_ptr(&L->res1);
      const float* res2 = ldg_ptr(&L->res2);
      const int out_ld = __ldg(&L->out_ld), out2_ld = __ldg(&L->out2_ld);
      const int res1_ld = __ldg(&L->res1_ld), res2_ld = __ldg(&L->res2_ld);
      const float alpha1 = __ldg(&L->alpha1), alpha2 = __ldg(&L->alpha2);
      const uint32_t acc = t_it & 1u;
      mbar_wait(tmem_full(acc), (t_it >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int gy = y0 + mt * TH + m / TW, gx = x0 + m % TW;
        const bool inb = (gy < p.H) && (gx < p.W);
        const size_t pix = ((size_t)b * p.H + (inb ? gy : 0)) * p.W + (inb ? gx : 0);
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 16) {
          float v[16];
          const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (acc * MT + mt) * p.nb_max + (uint32_t)c0;
          tmem_ld16(tcol, v);
          if (PASSES == 3 && parts == 2) {
            float lo[16];
            tmem_ld16(tcol + (uint32_t)N, lo);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += lo[j];
          }
          if (!inb || c0 >= cout || (p.debug & 8)) continue;
          if (bias) {   // bias / scale are padded to >= N entries and 16-byte aligned
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(bias + c0 + j));
              v[j] += t4.x; v[j + 1] += t4.y; v[j + 2] += t4.z; v[j + 3] += t4.w;
            }
          }
          if (scale) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(scale + c0 + j));
              v[j] *= t4.x; v[j + 1] *= t4.y; v[j + 2] *= t4.z; v[j + 3] *= t4.w;
            }
          }
          if (act == HCF_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (act == HCF_ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
          }
          // residuals may have been written by another SM earlier in this launch: L2-coherent loads
          if (out_vec && c0 + 15 < cout) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              if (res1) {
                const float4 rr = __ldcg(reinterpret_cast<const float4*>(res1 + pix * res1_ld + c0 + j));
                o.x = o.x * alpha1 + rr.x; o.y = o.y * alpha1 + rr.y;
                o.z = o.z * alpha1 + rr.z; o.w = o.w * alpha1 + rr.w;
              }
              if (res2) {
                const float4 rr = __ldcg(reinterpret_cast<const float4*>(res2 + pix * res2_ld + c0 + j));
                o.x = o.x * alpha2 + rr.x; o.y = o.y * alpha2 + rr.y;
                o.z = o.z * alpha2 + rr.z; o.w = o.w * alpha2 + rr.w;
              }
              *reinterpret_cast<float4*>(out + pix * out_ld + c0 + j) = o;
              if (out2) *reinterpret_cast<float4*>(out2 + pix * out2_ld + c0 + j) = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c = c0 + j;
              if (c < cout) {
                float t = v[j];
                if (res1) t = t * alpha1 + __ldcg(res1 + pix * res1_ld + c);
                if (res2) t = t * alpha2 + __ldcg(res2 + pix * res2_ld + c);
                out[pix * out_ld + c] = t;
                if (out2) out2[pix * out2_ld + c] = t;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty(acc));   // all TMEM reads of this item are complete (wait::ld above)
      if (chain) {
        // publish: every epilogue thread's stores -> gpu-scope fence -> 128-thread barrier -> counter
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) atomicAdd(p.done + tile, 1);
      }
    }
  } else {
    // ===================== A_lo converters (PASSES == 3) =====================
    if (PASSES == 3) {
      const int et = threadIdx.x - 192;   // 0..127
      uint32_t a_it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
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

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = NUM_SMS_FALLBACK;
  }
  return n;
}

// ring depths and B slot granularity that fit in shared memory; false if nothing fits
static bool pick_rings(int mt, int passes, int ks, int NB, int* sa, int* sb, int* slab_taps, size_t* smem) {
  const int a_stage = a_part(mt, ks) * (passes == 3 ? 2 : 1);
  const int budget = SMEM_LIMIT - 1024 - 512;
  const int tap = NB * ROW_BYTES;
  const int taps = ks * ks;
  auto done = [&](int a, int b, int st) {
    *sa = a; *sb = b; *slab_taps = st;
    *smem = 1024 + (size_t)a * a_stage + (size_t)b * st * tap + 512;
    return true;
  };
  // 1) whole-chunk slots (one barrier round trip per chunk): >= 2 of them beside >= 2 A stages
  for (int b = 3; b >= 2; --b)
    for (int a = 4; a >= 2; --a)
      if (a * a_stage + b * taps * tap <= budget && (a >= 3 || b == 2)) return done(a, b, taps);
  if (ks == 1) return false;
  // 2) one dy row of taps per slot beside >= 2 A stages.  Measured (tools/gpu_rings.sh): double-buffered A
  //    with only two row slots beats a single A stage with a whole chunk of B in flight, and per-tap
  //    slots lose to both (every slot costs a barrier round trip).
  for (int a = 4; a >= 2; --a)
    if (a * a_stage + 2 * ks * tap <= budget) {
      int b = (budget - a * a_stage) / (ks * tap);
      if (b > 9) b = 9;
      if (a > 2 && b < 4) {
        --a;
        b = (budget - a * a_stage) / (ks * tap);
        if (b > 9) b = 9;
      }
      return done(a, b, ks);
    }
  // 3) single-tap slots: two A stages if at least 6 taps of B still fit, else one
  for (int a = 2; a >= 1; --a)
    if (a * a_stage + (a == 2 ? 6 : 3) * tap <= budget) {
      int b = (budget - a * a_stage) / tap;
      if (b > 18) b = 18;
      return done(a, b, 1);
    }
  return false;
}

typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                         const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const Params);

static KernelFn pick_kernel(int mt, int passes, int ks) {
  if (ks == 1) return passes == 3 ? conv_tc_kernel<1, 3, 1> : conv_tc_kernel<1, 1, 1>;
  if (passes == 3) return conv_tc_kernel<1, 3, 3>;
  return mt == 2 ? conv_tc_kernel<2, 1, 3> : conv_tc_kernel<1, 1, 3>;
}

}  // namespace tc
}  // namespace hcf

struct hcf_conv_tc_plan {
  CUtensorMap maps[hcf::tc::MAX_MAPS];
  hcf::tc::Params p;
  hcf::tc::KernelFn fn;
  hcf::tc::LayerDesc* d_layers;
  size_t smem_bytes;
  int threads;
  dim3 grid;
};

// Any channel count works per segment: the tensor map's channel extent is the segment's C and the
// 32-channel box is zero-filled beyond it (the packed weights carry zero rows there too).
extern "C" int hcf_conv_tc_supported(const hcf_conv_args* a) {
  if (!a) return 0;
  if (a->ks != 3 && a->ks != 1) return 0;
  if (a->nseg < 1 || a->nseg > 3) return 0;
  for (int i = 0; i < a->nseg; ++i) {
    if (a->seg[i].up_shift != 0 || a->seg[i].C < 1) return 0;
    if (a->seg[i].ld % 4 != 0 || !hcf::aligned16(a->seg[i].ptr)) return 0;
  }
  if (a->cout < 1 || a->cout > 128) return 0;
  return 1;
}

// kin = number of (segment-padded) input channels, a multiple of 32
extern "C" int64_t hcf_conv_tc_weight_bytes(int32_t kin, int32_t cout, int32_t ks, int32_t passes) {
  if (kin % 32 != 0 || cout < 1 || cout > 128 || (ks != 1 && ks != 3) || (passes != 1 && passes != 3)) return 0;
  return (int64_t)(kin / 32) * ks * ks * hcf::tc::n_for(cout) * (passes == 3 ? 2 : 1) * 128;
}

// w: [cout][kin][ks][ks] fp32 (host), kin already padded per segment to multiples of 32
extern "C" int hcf_conv_tc_pack_weights(const float* w, int32_t kin, int32_t cout, int32_t ks, int32_t passes,
                                        float* image) {
  using namespace hcf;
  HCF_REQUIRE(w && image && kin % 32 == 0 && cout >= 1 && cout <= 128 && (ks == 1 || ks == 3) &&
                  (passes == 1 || passes == 3), "tc_pack: bad args");
  const int N = tc::n_for(cout), KC = kin / 32, parts = passes == 3 ? 2 : 1, NB = N * parts;
  memset(image, 0, (size_t)hcf_conv_tc_weight_bytes(kin, cout, ks, passes));
  for (int kc = 0; kc < KC; ++kc)
    for (int dy = 0; dy < ks; ++dy)
      for (int dx = 0; dx < ks; ++dx)
        for (int part = 0; part < parts; ++part)
          for (int n = 0; n < cout; ++n)
            for (int j = 0; j < 32; ++j) {
              const float v = w[(((size_t)n * kin + kc * 32 + j) * ks + dy) * ks + dx];
              uint32_t bits;
              memcpy(&bits, &v, 4);
              bits &= 0xFFFFE000u;
              float hi;
              memcpy(&hi, &bits, 4);
              const float val = part == 0 ? v : v - hi;
              const int row = part * N + n;
              const int chunk = (j / 4) ^ (row & 7);   // 128B swizzle: 16-byte chunk index XOR row-in-atom
              image[((((size_t)kc * ks + dy) * ks + dx) * NB + row) * 32 + chunk * 4 + (j & 3)] = val;
            }
  return 0;
}

// A chain of n convolutions on the same [B,H,W] grid with the same kernel size, each TC-eligible,
// executed by one persistent launch.  Conv i may read anything convs < i wrote (dependencies are
// tracked per 3x3 tile neighbourhood, which also covers write-after-read).  `done_flags`
// (device, B*ceil(H/16)*ceil(W/8) int32) must be zeroed before every run; NULL is allowed for n == 1.
extern "C" int hcf_conv_chain_create(const hcf_conv_args* args, const float* const* wtc, const int32_t* layer_passes,
                                     int32_t n, int32_t* done_flags, hcf_conv_tc_plan** out) {
  using namespace hcf;
  HCF_REQUIRE(out != nullptr, "tc_chain: null out");
  *out = nullptr;
  HCF_REQUIRE(args && wtc && layer_passes && n >= 1, "tc_chain: bad args");
  HCF_REQUIRE(n == 1 || done_flags != nullptr, "tc_chain: a chain needs the done-flag array");
  int passes = 1;   // kernel variant: 3 as soon as one layer uses the 3xTF32 split
  for (int i = 0; i < n; ++i) {
    HCF_REQUIRE(layer_passes[i] == 1 || layer_passes[i] == 3, "tc_chain: conv %d: passes %d", i, layer_passes[i]);
    if (layer_passes[i] == 3) passes = 3;
  }
  const int ks = args[0].ks;
  for (int i = 0; i < n; ++i) {
    int rc = validate_conv_args(&args[i]);
    if (rc) return rc;
    HCF_REQUIRE(hcf_conv_tc_supported(&args[i]), "tc_chain: conv %d: unsupported shape", i);
    HCF_REQUIRE(wtc[i] && aligned16(wtc[i]), "tc_chain: conv %d: weight image alignment", i);
    HCF_REQUIRE(args[i].ks == ks && args[i].B == args[0].B && args[i].H == args[0].H && args[i].W == args[0].W,
                "tc_chain: conv %d: all convs of a chain share ks and [B,H,W]", i);
  }
  tc::EncodeTiledFn enc = tc::get_encode();
  HCF_REQUIRE(enc != nullptr, "tc_chain: cuTensorMapEncodeTiled entry point not found");
  const hcf_conv_args* a0 = &args[0];
  hcf_conv_tc_plan* pl = new hcf_conv_tc_plan();
  memset(pl, 0, sizeof(*pl));
  tc::Params& p = pl->p;
  p.B = a0->B; p.H = a0->H; p.W = a0->W;
  const int sms = tc::num_sms();
  int nmax = 0, nbmax = 0;
  for (int i = 0; i < n; ++i) {
    const int nn = tc::n_for(args[i].cout), nb = nn * (layer_passes[i] == 3 ? 2 : 1);
    nmax = nmax > nn ? nmax : nn;
    nbmax = nbmax > nb ? nbmax : nb;
  }
  p.nb_max = nbmax;
  // sub-tiles per work item: 2 halves the weight traffic per pixel but quantises worse on small images
  int mt = 1;
  const char* env = getenv("HCF_TC_MT");
  if (passes == 1 && ks == 3 && n == 1) {
    const long items1 = (long)a0->B * ceil_div(a0->H, 16) * ceil_div(a0->W, 8);
    const long items2 = (long)a0->B * ceil_div(a0->H, 32) * ceil_div(a0->W, 8);
    const double t1 = (double)ceil_div((int)items1, sms) * (tc::a_part(1, 3) + 9.0 * nmax * 128);
    const double t2 = (double)ceil_div((int)items2, sms) * (tc::a_part(2, 3) + 9.0 * nmax * 128);
    mt = (t2 < 0.80 * t1) ? 2 : 1;
    if (env && (env[0] == '1' || env[0] == '2')) mt = env[0] - '0';
  }
  {
    const char* dbg = getenv("HCF_TC_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  int slot_taps = 0;
  if (!tc::pick_rings(mt, passes, ks, p.nb_max, &p.sa, &p.sb, &slot_taps, &pl->smem_bytes)) {
    delete pl;
    set_error("tc_chain: tile does not fit in shared memory");
    return HCF_ENOTSUP;
  }
  if (const char* rings = getenv(n > 1 ? "HCF_TC_RINGS" : "HCF_TC_RINGS_SINGLE")) {   // tuning: "sa,sb,slab_taps"
    int ra = 0, rb = 0, rs = 0;
    if (sscanf(rings, "%d,%d,%d", &ra, &rb, &rs) == 3 && ra >= 1 && ra <= 4 && rb >= 2 && rb <= 18 &&
        (rs == 1 || rs == ks || rs == ks * ks)) {
      const size_t need = 1024 + (size_t)ra * tc::a_part(mt, ks) * (passes == 3 ? 2 : 1) +
                          (size_t)rb * rs * p.nb_max * tc::ROW_BYTES + 512;
      if (need <= (size_t)tc::SMEM_LIMIT) {
        p.sa = ra; p.sb = rb; slot_taps = rs;
        pl->smem_bytes = need;
      }
    }
  }
  p.slot_bytes = slot_taps * p.nb_max * tc::ROW_BYTES;
  p.tiles_x = ceil_div(a0->W, tc::TW); p.tiles_y = ceil_div(a0->H, tc::TH * mt);
  p.n_tiles = p.tiles_x * p.tiles_y * a0->B;
  p.n_layers = n;
  p.n_items = p.n_tiles * n;
  p.done = n > 1 ? done_flags : nullptr;

  // ---- tensor maps, de-duplicated.  A segment whose channel count is a multiple of 32 never reads
  // beyond its last chunk, so such segments of one buffer share a map of the widest extent seen.
  struct MapKey { const float* ptr; int ld; int C; bool ragged; };
  std::vector<MapKey> keys;
  std::vector<tc::LayerDesc> layers(n);
  for (int i = 0; i < n; ++i) {
    const hcf_conv_args& a = args[i];
    tc::LayerDesc& L = layers[i];
    memset(&L, 0, sizeof(L));
    L.nseg = a.nseg;
    int kc = 0;
    for (int s = 0; s < 3; ++s) {
      if (s < a.nseg) {
        const hcf_seg& sg = a.seg[s];
        const bool ragged = (sg.C % 32) != 0;
        int found = -1;
        for (size_t k = 0; k < keys.size(); ++k)
          if (keys[k].ptr == sg.ptr && keys[k].ld == sg.ld && keys[k].ragged == ragged && (!ragged || keys[k].C == sg.C))
            found = (int)k;
        if (found < 0) {
          keys.push_back({sg.ptr, sg.ld, sg.C, ragged});
          found = (int)keys.size() - 1;
        } else if (!ragged && keys[found].C < sg.C) {
          keys[found].C = sg.C;
        }
        L.map_idx[s] = found;
        kc += (sg.C + 31) / 32;
      }
      L.seg_end[s] = s < a.nseg ? kc : (1 << 30);
    }
    L.seg_end[a.nseg - 1] = 1 << 30;
    L.kchunks = kc;
    L.N = tc::n_for(a.cout);
    L.parts = layer_passes[i] == 3 ? 2 : 1;
    {   // largest tap count (whole chunk, one dy row, one tap) of this layer that fits a ring slot
      const int tap = L.N * L.parts * tc::ROW_BYTES;
      L.slab_taps = (ks * ks * tap <= p.slot_bytes) ? ks * ks : ((ks * tap <= p.slot_bytes) ? ks : 1);
    }
    L.cout = a.cout;
    L.act = a.act;
    L.wimg = wtc[i]; L.bias = a.bias; L.scale = a.scale;
    L.out = a.out; L.out_ld = a.out_ld; L.out2 = a.out2; L.out2_ld = a.out2_ld;
    L.res1 = a.res1; L.res1_ld = a.res1_ld; L.alpha1 = a.alpha1;
    L.res2 = a.res2; L.res2_ld = a.res2_ld; L.alpha2 = a.alpha2;
    bool ov = aligned16(a.out) && a.out_ld % 4 == 0;
    if (a.out2) ov = ov && aligned16(a.out2) && a.out2_ld % 4 == 0;
    if (a.res1) ov = ov && aligned16(a.res1) && a.res1_ld % 4 == 0;
    if (a.res2) ov = ov && aligned16(a.res2) && a.res2_ld % 4 == 0;
    L.out_vec = ov ? 1 : 0;
  }
  if ((int)keys.size() > tc::MAX_MAPS) {
    delete pl;
    set_error("tc_chain: %d distinct input views (max %d)", (int)keys.size(), tc::MAX_MAPS);
    return HCF_ENOTSUP;
  }
  const cuuint32_t box[4] = {(cuuint32_t)tc::KCH, (cuuint32_t)tc::halo_w(ks), (cuuint32_t)tc::halo_rows(mt, ks), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < tc::MAX_MAPS; ++i) {
    const MapKey& k = keys[i < (int)keys.size() ? i : 0];   // unused maps alias map 0 (never dereferenced)
    const cuuint64_t dims[4] = {(cuuint64_t)k.C, (cuuint64_t)a0->W, (cuuint64_t)a0->H, (cuuint64_t)a0->B};
    const cuuint64_t ld_b = (cuuint64_t)k.ld * 4;
    const cuuint64_t strides[3] = {ld_b, ld_b * a0->W, ld_b * a0->W * a0->H};
    CUresult r = enc(&pl->maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(k.ptr), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete pl;
      set_error("tc_chain: cuTensorMapEncodeTiled failed with %d (view %d, C %d, ld %d)", (int)r, i, k.C, k.ld);
      return HCF_EINVAL;
    }
  }
  cudaError_t e = cudaMalloc(&pl->d_layers, sizeof(tc::LayerDesc) * n);
  if (e == cudaSuccess)
    e = cudaMemcpy(pl->d_layers, layers.data(), sizeof(tc::LayerDesc) * n, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (pl->d_layers) cudaFree(pl->d_layers);
    delete pl;
    set_error("tc_chain: layer table upload: %s", cudaGetErrorString(e));
    return (int)e;
  }
  p.layers = pl->d_layers;
  pl->grid = dim3((unsigned)(p.n_tiles < sms ? p.n_tiles : sms));
  pl->threads = passes == 3 ? 320 : 192;
  pl->fn = tc::pick_kernel(mt, passes, ks);
  e = cudaFuncSetAttribute(reinterpret_cast<const void*>(pl->fn), cudaFuncAttributeMaxDynamicSharedMemorySize,
                           tc::SMEM_LIMIT);
  if (e != cudaSuccess) {
    cudaFree(pl->d_layers);
    delete pl;
    set_error("tc_chain: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return (int)e;
  }
  *out = pl;
  return 0;
}

extern "C" int hcf_conv_tc_plan_create(const hcf_conv_args* a, const float* wtc, int32_t passes,
                                       hcf_conv_tc_plan** out) {
  return hcf_conv_chain_create(a, &wtc, &passes, 1, nullptr, out);
}

extern "C" int32_t hcf_conv_tc_plan_layers(const hcf_conv_tc_plan* pl) { return pl ? pl->p.n_layers : 0; }

extern "C" int hcf_conv_tc_run(const hcf_conv_tc_plan* pl, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(pl != nullptr, "tc_run: null plan");
  pl->fn<<<pl->grid, pl->threads, pl->smem_bytes, (cudaStream_t)stream>>>(
      pl->maps[0], pl->maps[1], pl->maps[2], pl->maps[3], pl->maps[4], pl->maps[5], pl->maps[6], pl->maps[7], pl->p);
  return finish_launch("hcf_conv_tc_run");
}

extern "C" void hcf_conv_tc_plan_destroy(hcf_conv_tc_plan* p) {
  if (!p) return;
  if (p->d_layers) cudaFree(p->d_layers);
  delete p;
}
