// Weight-stationary schedule of a chained fp16 convolution launch (conv_tc.cu picks it for encoder-like chains; the
// per-item schedule of conv_tc_kernel.cuh stays the fallback and runs everything else).
//
// What the per-item kernel pays per (tile, layer) work item and this schedule does not:
//   * the layer's weights (36 .. 295 KB) were streamed into shared memory once per 16x8 tile: two thirds of the
//     L2 -> SM bytes of an RRDB chain.  Here a CTA OWNS up to three tiles per image group and walks a layer over
//     them chunk by chunk: every weight slab is loaded once per (layer, group) and multiplies all of the group's
//     activation tiles before it leaves (weight-stationary), each tile accumulating in its own TMEM slot;
//   * one dependency poll + gpu-scope acquire fence per item: one per group;
//   * activation stages that carried [hi | lo] although only an RDB's conv5 reads a lo plane: the ring holds single
//     planes (five stages instead of two pairs), the lo plane of a split chunk is its own K chunk (A_lo x B_hi into
//     the correction columns), and weights have their own producer warp, so neither ring waits for the other.
// STATUS: opt-in (HCF_TC_WS=1).  Parity-green and it does what it was built for (weight fills -2/3, MMA issue time per
// tile -19 %, L2 -> SM traffic 31.5 -> 20.0 GB on the 80x80 encoder chain), but the launch is slower than the per-item
// schedule (6.0 vs 5.2 ms): three tiles of a pass complete together and need two epilogue rounds before the pass can be
// published, static ownership pays six tile slots per layer for 5.4 tiles of work -- DESIGN.md section 4.2.
// Image groups ("phases") alternate: images are independent, so the tiles of group p at layer l + 1 depend only on
// group p at layer l, which finished one whole group pass earlier -- the dependency wait is off the critical path
// instead of being hidden by luck.
//
// Roles (384 threads): warp 0 activation producer (dependency poll, one 4-D TMA halo tile per chunk and tile),
// warp 1 MMA issuer (TMEM owner), warps 2..9 two epilogue groups (same code as the per-item kernel: TMEM -> staging
// transpose -> coalesced bias / activation / residual / fp32 + fp16 hi / lo stores), warp 10 publisher (gpu-scope
// release of finished tiles), warp 11 weight producer.
#pragma once
#include "conv_tc_kernel.cuh"

namespace hcf {
namespace tc {

constexpr int WS_SA = 5;                  // activation ring: single-plane halo tiles
constexpr int WS_SB = 2;                  // weight ring slots
constexpr int WS_SLOT = 36864;            // bytes per weight slot: all nine taps of a 32-row chunk
constexpr int WS_GMAX = 3;                // tiles a CTA owns per image group
constexpr int WS_MAX_CHUNKS = 12;
constexpr int WS_BIG_COLS = 128, WS_SMALL_COLS = 32;   // TMEM slots: three of each
constexpr int WS_TAB_BYTES = 448;         // per-role copy of the current layer's WsLayer record

// One K chunk of a layer = one activation plane tile per owned tile + its weight slabs.
struct WsChunk {
  int map;            // tensor map of the plane (hi, or lo for the A_lo x B_hi pass of a split chunk)
  int cch;            // channel coordinate in the map
  int kst;            // kmax | taps-per-slab << 8 | slabs << 16
  int rows;           // weight rows per tap = UMMA N of this chunk (2N for [B_hi ; B_lo], N otherwise)
  int dcol;           // first accumulator column (N for the lo pass: the correction columns)
  uint32_t idesc;
  uint32_t b_off;     // byte offset of tap 0 in the layer's weight image
  uint32_t b_tap_src; // bytes between taps in the image (> rows * 128 when only the hi rows are loaded)
};
struct WsLayer {
  int n_chunks, taps, tap0, nb;   // nb = accumulator columns (N * parts)
  unsigned long long wimg;
  int pad[2];
  WsChunk ch[WS_MAX_CHUNKS];
};
static_assert(sizeof(WsLayer) <= WS_TAB_BYTES, "WsLayer record does not fit its shared-memory copy");

// Every wait of this kernel is bounded: a protocol bug (or a missing CTA) prints where it stopped, raises the
// sticky status bit and lets the whole grid drain instead of hanging the GPU.
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
constexpr long long WS_TIMEOUT_CYCLES = 2000000000ll;   // ~1 s
__device__ __noinline__ void ws_wait_failed(int code, int a, int b, int c, volatile int* abort_flag, int* status) {
  if (atomicExch((int*)abort_flag, 1) == 0) {
    printf("[hcf ws] CTA %d warp %d: wait %d timed out (group %d, %d, %d)\n", (int)blockIdx.x, (int)(threadIdx.x >> 5),
           code, a, b, c);
    if (status) atomicOr(status, STATUS_DEP_TIMEOUT);
  }
}
#define WS_WAIT(bar, parity, code, a, b, c)                                                  \
  do {                                                                                       \
    uint32_t n_ = 0;                                                                         \
    long long t0_ = 0;                                                                       \
    while (!mbar_try(bar, parity)) {                                                         \
      if ((++n_ & 1023u) == 0) {                                                             \
        if (t0_ == 0) t0_ = clock64();                                                       \
        if (*abort_flag) break;                                                              \
        if (clock64() - t0_ > WS_TIMEOUT_CYCLES) {                                           \
          ws_wait_failed(code, a, b, c, abort_flag, p.status);                               \
          break;                                                                             \
        }                                                                                    \
      }                                                                                      \
    }                                                                                        \
  } while (0)

__device__ __forceinline__ void ws_load_tab(const WsLayer* __restrict__ src, uint32_t* dst, int lane) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(WsLayer) / 4 + 31) / 32; ++i) {
    const int idx = lane + 32 * i;
    if (idx < (int)(sizeof(WsLayer) / 4)) dst[idx] = __ldg(s + idx);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(384, 1)
conv_ws_kernel(const __grid_constant__ Maps maps, const Params p) {
  constexpr int HALO_W = TW + 2;
  constexpr int A_BYTES = a_bytes(1, 3);
  constexpr int A_PART = a_part(1, 3);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t b_base = smem_base + WS_SA * A_PART;
  const uint32_t bar_base = b_base + WS_SB * WS_SLOT;
  auto fullA = [&](int s) { return bar_base + 8u * s; };
  auto emptyA = [&](int s) { return bar_base + 48u + 8u * s; };
  auto fullB = [&](int s) { return bar_base + 96u + 8u * s; };
  auto emptyB = [&](int s) { return bar_base + 120u + 8u * s; };
  // accumulator-ready barriers: one per (TMEM slot, epilogue group that drains this use) -- a parity wait is only
  // meaningful for a waiter that observes every phase of its barrier, and consecutive uses of a slot alternate
  // between the two epilogue groups
  auto tfull = [&](int s, int e) { return bar_base + 144u + 8u * (2 * s + e); };
  auto tempty = [&](int s) { return bar_base + 240u + 8u * s; };
  const uint32_t dep_seq = bar_base + 288u;
  auto pub_seq = [&](int g) { return bar_base + 292u + 4u * g; };
  const uint32_t tmem_slot = bar_base + 300u;
  volatile int* abort_flag = reinterpret_cast<volatile int*>(gen_base + (bar_base + 304u - smem_base));
  uint32_t* tabs = reinterpret_cast<uint32_t*>(gen_base + (bar_base + TAIL_BYTES - smem_base));   // 3 x WS_TAB_BYTES

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[0]) : "memory");
    for (int s = 0; s < WS_SA; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
    for (int s = 0; s < WS_SB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
    for (int s = 0; s < 6; ++s) { mbar_init(tfull(s, 0), 1); mbar_init(tfull(s, 1), 1); mbar_init(tempty(s), 128); }
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(dep_seq), "r"(0u) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(pub_seq(0)), "r"(0u) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(pub_seq(1)), "r"(0u) : "memory");
    *abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int per_img = p.tiles_x * p.tiles_y;
  const int cta = (int)blockIdx.x, grid = (int)gridDim.x;
  // tiles of image group ph this CTA processes at `layer`: local indices c, c + grid, c + 2 grid with
  // c = (cta + layer * 53) mod grid -- the rotation spreads the CTAs that get one tile more than the others
  // (400 tiles on 148 CTAs: 104 x 3 + 44 x 2) evenly over the layers
  auto group_c = [&](int layer) -> int { return (cta + layer * 53) % grid; };
  auto group_n = [&](int ph, int c) -> int {
    const int nimg = min(p.ipp, p.B - ph * p.ipp);
    const int tp = nimg * per_img;
    return tp > c ? (tp - c - 1) / grid + 1 : 0;
  };
  auto group_tile = [&](int ph, int c, int g) -> int { return ph * p.ipp * per_img + c + g * grid; };
  const int n_groups = p.n_layers * p.n_phases;   // group index q = layer * n_phases + ph
  const WsLayer* wsl = reinterpret_cast<const WsLayer*>(p.ws_layers);
#ifdef HCF_TC_PROF_BUILD
  const bool prof_on = p.prof != nullptr;
  long long pacc[PROF_N];
#pragma unroll
  for (int i = 0; i < PROF_N; ++i) pacc[i] = 0;
#endif

  if (warp == 0) {
    // ===================== activation producer =====================
    uint32_t* tab = tabs;
    const WsLayer* T = reinterpret_cast<const WsLayer*>(tab);
    int sA = 0;
    uint32_t phA = 0, p_it = 0;
    int cur_layer = -1;
    HCF_T(tp0);
    for (int q = 0; q < n_groups; ++q) {
      const int layer = q / p.n_phases, ph = q - layer * p.n_phases;
      const int gc = group_c(layer);
      const int ng = group_n(ph, gc);
      if (ng == 0) continue;
      if (layer != cur_layer) {
        __syncwarp();
        ws_load_tab(wsl + layer, tab, lane);
        cur_layer = layer;
      }
      if (lane == 0) {
        int tb[WS_GMAX], ty[WS_GMAX], tx[WS_GMAX];
#pragma unroll
        for (int g = 0; g < WS_GMAX; ++g) {
          const int tile = group_tile(ph, gc, g < ng ? g : 0);
          tb[g] = tile / per_img;
          const int r = tile - tb[g] * per_img;
          ty[g] = r / p.tiles_x; tx[g] = r - ty[g] * p.tiles_x;
        }
        if (p.done != nullptr && layer > 0 && !(p.debug & 64)) {
          // layer - 1 complete on the 3x3 tile neighbourhood of every tile of the group (halo + WAR safety); the
          // group's previous layer ran one whole group pass ago, so this normally succeeds at the first poll
          HCF_T(td0);
          uint32_t spins = 0;
          for (;;) {
            Deps d0, d1, d2;
            load_deps(d0, p.done, tb[0] * per_img, ty[0], tx[0], p.tiles_y, p.tiles_x);
            if (ng > 1) load_deps(d1, p.done, tb[1] * per_img, ty[1], tx[1], p.tiles_y, p.tiles_x);
            if (ng > 2) load_deps(d2, p.done, tb[2] * per_img, ty[2], tx[2], p.tiles_y, p.tiles_x);
            bool ok = deps_ready(d0, layer);
            if (ng > 1) ok = ok && deps_ready(d1, layer);
            if (ng > 2) ok = ok && deps_ready(d2, layer);
            if (ok) break;
            __nanosleep(32);
            if (*abort_flag) break;
            if (++spins > (1u << 22)) {
              ws_wait_failed(0, q, layer, ng, abort_flag, p.status);
              break;
            }
          }
          fence_acquire_gpu();
          asm volatile("fence.proxy.async.global;" ::: "memory");
          HCF_T(td1);
          HCF_ACC(PROF_P_DEPS, td0, td1);
        }
        p_it += (uint32_t)ng;
        asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(dep_seq), "r"(p_it) : "memory");   // epilogues may prefetch
        const int nch = T->n_chunks;
        for (int j = 0; j < nch; ++j) {
          const int mi = T->ch[j].map, cch = T->ch[j].cch;
#pragma unroll
          for (int g = 0; g < WS_GMAX; ++g) {
            if (g < ng) {
              HCF_T(ta0);
              WS_WAIT(emptyA(sA), phA ^ 1u, 1, q, j, g);
              HCF_T(ta1);
              HCF_ACC(PROF_P_EMPTYA, ta0, ta1);
              if (p.debug & 4) {
                mbar_arrive(fullA(sA));
              } else {
                mbar_expect_tx(fullA(sA), A_BYTES);
                tma_load_4d(smem_base + sA * A_PART, &maps.m[mi], fullA(sA), cch, tx[g] * TW - 1, ty[g] * TH - 1, tb[g]);
              }
              if (++sA == WS_SA) { sA = 0; phA ^= 1u; }
            }
          }
        }
      }
    }
    HCF_T(tp1);
    HCF_ACC(PROF_P_TOTAL, tp0, tp1);
    if (lane == 0) HCF_PROF_FLUSH(PROF_P_TOTAL, PROF_P_EMPTYA);
  } else if (warp == 11) {
    // ===================== weight producer =====================
    uint32_t* tab = tabs + WS_TAB_BYTES / 4;
    const WsLayer* T = reinterpret_cast<const WsLayer*>(tab);
    int sB = 0;
    uint32_t phB = 0;
    int cur_layer = -1;
    for (int q = 0; q < n_groups; ++q) {
      const int layer = q / p.n_phases, ph = q - layer * p.n_phases;
      if (group_n(ph, group_c(layer)) == 0) continue;
      if (layer != cur_layer) {
        __syncwarp();
        ws_load_tab(wsl + layer, tab, lane);
        cur_layer = layer;
      }
      if (lane == 0) {
        const uint8_t* wimg = reinterpret_cast<const uint8_t*>(T->wimg);
        const int nch = T->n_chunks, taps = T->taps;
        for (int j = 0; j < nch; ++j) {
          const int kst = T->ch[j].kst;
          const int tps = (kst >> 8) & 0xff, slabs = (kst >> 16) & 0xff;
          const uint32_t tap_bytes = (uint32_t)T->ch[j].rows * ROW_BYTES;
          const uint32_t tap_src = T->ch[j].b_tap_src;
          const uint8_t* src = wimg + T->ch[j].b_off;
          for (int s = 0; s < slabs; ++s) {
            const int nt = min(tps, taps - s * tps);
            HCF_T(tb0);
            WS_WAIT(emptyB(sB), phB ^ 1u, 2, q, j, s);
            HCF_T(tb1);
            HCF_ACC(PROF_P_EMPTYB, tb0, tb1);
            if (p.debug & 4) {
              mbar_arrive(fullB(sB));
            } else {
              mbar_expect_tx(fullB(sB), (uint32_t)nt * tap_bytes);
              const uint32_t dst = b_base + sB * WS_SLOT;
              if (tap_src == tap_bytes) {
                bulk_load(dst, src + (size_t)(s * tps) * tap_src, (uint32_t)nt * tap_bytes, fullB(sB));
              } else {
                for (int t = 0; t < nt; ++t)
                  bulk_load(dst + (uint32_t)t * tap_bytes, src + (size_t)(s * tps + t) * tap_src, tap_bytes, fullB(sB));
              }
            }
            if (++sB == WS_SB) { sB = 0; phB ^= 1u; }
          }
        }
      }
    }
    if (lane == 0) HCF_PROF_FLUSH(PROF_P_EMPTYB, PROF_P_EMPTYB);
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    uint32_t* tab = tabs + 2 * (WS_TAB_BYTES / 4);
    const WsLayer* T = reinterpret_cast<const WsLayer*>(tab);
    const uint64_t a_tmpl = make_desc(0, HALO_W * ROW_BYTES);
    const uint64_t b_tmpl = make_desc(0, 8u * ROW_BYTES);
    int sA = 0, sB = 0;
    uint32_t phA = 0, phB = 0;
    uint32_t slot_par = 0;    // bit s: uses of TMEM slot s so far, mod 2
    uint32_t t_it = 0;        // running item index: item t_it is drained by epilogue group t_it & 1
    bool prev_big = false;
    int cur_layer = -1;
    HCF_T(tm0);
    for (int q = 0; q < n_groups; ++q) {
      const int layer = q / p.n_phases, ph = q - layer * p.n_phases;
      const int gc = group_c(layer);
      const int ng = group_n(ph, gc);
      if (ng == 0) continue;
      if (layer != cur_layer) {
        __syncwarp();
        ws_load_tab(wsl + layer, tab, lane);
        cur_layer = layer;
      }
      const int nch = T->n_chunks, taps = T->taps, tap0 = T->tap0;
      const bool big = T->nb > WS_SMALL_COLS || !prev_big;
      prev_big = big;
      const int slot0 = big ? 0 : 3;
      const uint32_t col0 = big ? 0u : 3u * WS_BIG_COLS, colw = big ? WS_BIG_COLS : WS_SMALL_COLS;
      for (int j = 0; j < nch; ++j) {
        const int kst = T->ch[j].kst;
        const int kmax = kst & 0xff, tps = (kst >> 8) & 0xff, slabs = (kst >> 16) & 0xff;
        const uint32_t rows8 = (uint32_t)T->ch[j].rows * (ROW_BYTES >> 4);
        const uint32_t idesc = T->ch[j].idesc;
        const uint32_t dcol = (uint32_t)T->ch[j].dcol;
        for (int s = 0; s < slabs; ++s) {
          const int nt = (p.debug & 2) ? 0 : min(tps, taps - s * tps);
          HCF_T(tfb0);
          WS_WAIT(fullB(sB), phB, 3, q, j, s);
          tc_fence_after();
          HCF_T(tfb1);
          HCF_ACC(PROF_M_FULLB, tfb0, tfb1);
          const uint64_t b0 = b_tmpl + ((b_base + sB * WS_SLOT) >> 4);
          for (int g = 0; g < ng; ++g) {
            int st = sA + g;
            uint32_t par = phA;
            if (st >= WS_SA) { st -= WS_SA; par ^= 1u; }
            if (s == 0) {
              HCF_T(tfa0);
              WS_WAIT(fullA(st), par, 4, q, j, g);
              HCF_T(tfa1);
              HCF_ACC(PROF_M_FULLA, tfa0, tfa1);
              if (j == 0) {
                WS_WAIT(tempty(slot0 + g), ((slot_par >> (slot0 + g)) & 1u) ^ 1u, 5, q, slot0, g);
                HCF_T(tfa2);
                HCF_ACC(PROF_M_TMEM, tfa1, tfa2);
              }
              tc_fence_after();
            }
            HCF_T(tis0);
            if (elect_one()) {
              const uint64_t a0 = a_tmpl + ((smem_base + st * A_PART) >> 4);
              const uint32_t d = tmem_base + col0 + (uint32_t)g * colw + dcol;
              uint32_t accum = (j == 0 && s == 0) ? 0u : 1u;
              for (int t = 0; t < nt; ++t) {
                const int tap = tap0 + s * tps + t;
                const int dy = tap / 3, dx = tap - dy * 3;
                const uint64_t a_tap = a0 + (uint32_t)((dy * HALO_W + dx) * (ROW_BYTES >> 4));
                const uint64_t b_tap = b0 + (uint32_t)t * rows8;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (k >= kmax) break;
                  umma_f16(d, a_tap + 2u * k, b_tap + 2u * k, idesc, accum);
                  accum = 1u;
                }
              }
              if (s == slabs - 1) {
                umma_commit(emptyA(st));
                if (j == nch - 1) umma_commit(tfull(slot0 + g, (int)((t_it + (uint32_t)g) & 1u)));
              }
              if (g == ng - 1) umma_commit(emptyB(sB));
            }
            __syncwarp();
            HCF_T(tis1);
            HCF_ACC(PROF_M_ISSUE, tis0, tis1);
          }
          if (++sB == WS_SB) { sB = 0; phB ^= 1u; }
        }
        sA += ng;
        if (sA >= WS_SA) { sA -= WS_SA; phA ^= 1u; }
      }
      for (int g = 0; g < ng; ++g) slot_par ^= 1u << (slot0 + g);
      t_it += (uint32_t)ng;
    }
    HCF_T(tm1);
    HCF_ACC(PROF_M_TOTAL, tm0, tm1);
    if (lane == 0) { HCF_PROF_FLUSH(PROF_M_TOTAL, PROF_M_FULLB); HCF_PROF_FLUSH(PROF_M_ISSUE, PROF_M_ISSUE); }
  } else if (warp < 10) {
    // ===================== epilogue (two groups of four warps; the per-item kernel's store path) =====================
    const int grp = (warp - 2) >> 2;
    const int qd = warp & 3;                      // TMEM lane quarter this warp may access
    const int et = (threadIdx.x - 64) & 127;
    float* s_bias = reinterpret_cast<float*>(gen_base + (bar_base + BAR_BYTES - smem_base)) + grp * (256 + STAGE_BYTES / 4);
    float* s_scale = s_bias + 128;
    float4* stage = reinterpret_cast<float4*>(s_bias + 256) + qd * 256;
    uint32_t t_it = 0;            // running item index over (layer, group, tile): item t_it belongs to group t_it & 1
    uint32_t slot_par = 0;        // bit 2 * slot + e: uses of accumulator-ready barrier (slot, e) so far, mod 2
    bool prev_big = false;
    int cur_layer = -1, N = 0, cout = 0, act = 0, out_vec = 0, parts = 1, nb = 0;
    int out_ld = 0, res1_ld = 0, res2_ld = 0;
    int fast = 0;
    float* out = nullptr;
    Out16 o16 = {nullptr, nullptr, nullptr, nullptr};
    const float* res1 = nullptr; const float* res2 = nullptr;
    float alpha1 = 0.f, alpha2 = 0.f;
    bool is_pre = false;
    int grp_layer = -1;           // layer whose bias / scale this group has staged
    HCF_T(te0);
    for (int q = 0; q < n_groups; ++q) {
      const int layer = q / p.n_phases, ph = q - layer * p.n_phases;
      const int gc = group_c(layer);
      const int ng = group_n(ph, gc);
      if (ng == 0) continue;
      if (layer != cur_layer) {
        cur_layer = layer;
        nb = __ldg(&wsl[layer].nb);
      }
      const bool big = nb > WS_SMALL_COLS || !prev_big;
      prev_big = big;
      const int slot0 = big ? 0 : 3;
      const uint32_t col0 = big ? 0u : 3u * WS_BIG_COLS, colw = big ? WS_BIG_COLS : WS_SMALL_COLS;
      for (int g = 0; g < ng; ++g, ++t_it) {
        const int fidx = 2 * (slot0 + g) + (int)(t_it & 1u);   // this use's accumulator-ready barrier
        const uint32_t fpar = (slot_par >> fidx) & 1u;
        slot_par ^= 1u << fidx;
        if ((int)(t_it & 1u) != grp) continue;
        const int tile = group_tile(ph, gc, g);
        const int b = tile / per_img, r = tile - b * per_img;
        const int y0 = (r / p.tiles_x) * TH, x0 = (r % p.tiles_x) * TW;
        HCF_T(tl0);
        if (layer != grp_layer) {
          grp_layer = layer;
          const LayerDesc* L = p.layers + layer;
          N = __ldg(&L->N); cout = __ldg(&L->cout); act = __ldg(&L->act); out_vec = __ldg(&L->out_vec);
          parts = __ldg(&L->parts);
          out = ldg_ptr(&L->out);
          o16.hi = ldg_ptr(&L->out_hi); o16.lo = ldg_ptr(&L->out_lo);
          const bool second = ldg_ptr(&L->out2) != nullptr || ldg_ptr(&L->out2_hi) != nullptr || ldg_ptr(&L->out2_lo) != nullptr;
          res1 = ldg_ptr(&L->res1); res2 = ldg_ptr(&L->res2);
          out_ld = __ldg(&L->out_ld);
          res1_ld = __ldg(&L->res1_ld); res2_ld = __ldg(&L->res2_ld);
          alpha1 = __ldg(&L->alpha1); alpha2 = __ldg(&L->alpha2);
          is_pre = false;
          if (const float* pre = ldg_ptr(&L->pre)) {
            res1 = pre; res1_ld = __ldg(&L->pre_ld); is_pre = true;
          }
          fast = 0;
          if (out_vec && cout % 32 == 0 && !second && (o16.lo == nullptr || o16.hi != nullptr))
            fast = (out ? 1 : 0) | (o16.hi ? 2 : 0) | (o16.lo ? 4 : 0);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          s_bias[et] = __ldg(p.epi + (size_t)layer * 256 + et);
          s_scale[et] = __ldg(p.epi + (size_t)layer * 256 + 128 + et);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        }
        const int slot = slot0 + g;
        const uint32_t tcol0 = tmem_base + ((uint32_t)(qd * 32) << 16) + col0 + (uint32_t)g * colw;
        const bool res_pf = out_vec && (res1 != nullptr || res2 != nullptr) && !(p.debug & 8);
        float4 r1v[8], r2v[8];
        uint32_t pixv[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int mm = qd * 32 + it * 4 + (lane >> 3);
          const int gy = y0 + mm / TW, gx = x0 + mm % TW;
          pixv[it] = (gy < p.H && gx < p.W) ? (uint32_t)((b * p.H + gy) * p.W + gx) : 0xffffffffu;
        }
        const bool ahead2 = res_pf && res2 == nullptr && N > 32 && N <= 64;
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
        if (res_pf) {
          uint32_t seen;
          do {
            asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(seen) : "r"(dep_seq) : "memory");
          } while (seen <= t_it && !*abort_flag);
          res_prefetch(0);
          if (ahead2) {
            const int ch_ = 32 + (lane & 7) * 4;
            if (ch_ + 3 < cout) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                if (pixv[it] != 0xffffffffu)
                  r2v[it] = __ldcg(reinterpret_cast<const float4*>(res1 + (pixv[it] * (uint32_t)res1_ld + ch_)));
            }
          }
        }
        HCF_T(tl1);
        WS_WAIT(tfull(slot, grp), fpar, 6, q, slot, g);
        tc_fence_after();
        HCF_T(tl2);
        HCF_ACC(PROF_E_LAYER, tl0, tl1);
        HCF_ACC(PROF_E_TMEMFULL, tl1, tl2);
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          const int gw = min(32, N - c0);
          HCF_T(tr0);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h * 16 < gw) {
              float v[16];
              const uint32_t tcol = tcol0 + (uint32_t)(c0 + h * 16);
              tmem_ld16(tcol, v);
              if (parts == 2) {
                float lo[16];
                tmem_ld16(tcol + (uint32_t)N, lo);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaf(lo[j], 1.0f / 2048.0f, v[j]);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j)
                stage[lane * 8 + ((h * 4 + j) ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
          __syncwarp();
          if (c0 + 32 >= N) {   // last TMEM read of the item: hand the slot back before the store phase
            tc_fence_before();
            mbar_arrive(tempty(slot));
          }
          HCF_T(tr1);
          HCF_ACC(PROF_E_ROW, tr0, tr1);
          if (!(p.debug & 8)) {
            const int cidx = lane & 7;
            const int ch = c0 + cidx * 4;
            const bool ch_ok = ch < cout && cidx * 4 < gw;
            const bool vec = out_vec && ch + 3 < cout;
            const bool hr1 = res1 != nullptr && !is_pre, hr2 = res2 != nullptr;
            Chan4 cc;
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
            // generic path (second output view, ragged channel count): a few layers per chain -- its extra fields
            // are fetched here instead of living in registers across the whole loop
            float* out2 = nullptr;
            int out2_ld = 0;
            bool has_bias = false, has_scale = false;
            if (!fast) {
              const LayerDesc* L = p.layers + layer;
              out2 = ldg_ptr(&L->out2); out2_ld = __ldg(&L->out2_ld);
              o16.hi2 = ldg_ptr(&L->out2_hi); o16.lo2 = ldg_ptr(&L->out2_lo);
              has_bias = ldg_ptr(&L->bias) != nullptr; has_scale = ldg_ptr(&L->scale) != nullptr;
            }
#pragma unroll
            for (int it = 0; it < (fast ? 0 : 8); ++it) {
              const int pl = it * 4 + (lane >> 3);
              if (pixv[it] != 0xffffffffu && ch_ok) {
                const uint32_t e1 = pixv[it] * (uint32_t)out_ld + ch;
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
                } else {
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
                      if (!(fabsf(t) <= 65504.0f)) {
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
            if (res_pf) {
              if (ahead2) {
#pragma unroll
                for (int it = 0; it < 8; ++it) r1v[it] = r2v[it];
              } else if (c0 + 32 < N) {
                res_prefetch(c0 + 32);
              }
            }
          }
          __syncwarp();
          HCF_T(tr2);
          HCF_ACC(PROF_E_COAL, tr1, tr2);
        }
        HCF_T(tl3);
        HCF_ACC(PROF_E_BODY, tl2, tl3);
        if (p.done != nullptr) {
          // hand the gpu-scope release to the publisher warp: wait until it has published this group's previous item
          // (so that the named barrier is at most one phase ahead), then arrive without blocking
          const uint32_t own = t_it >> 1;
          uint32_t seen;
          do {
            asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(seen) : "r"(pub_seq(grp)) : "memory");
          } while (seen < own && !*abort_flag);
          asm volatile("bar.arrive %0, 160;" ::"r"(3 + grp) : "memory");
        }
        HCF_T(tl4);
        HCF_ACC(PROF_E_PUBLISH, tl3, tl4);
      }
    }
    HCF_T(te1);
    HCF_ACC(PROF_E_TOTAL, te0, te1);
    if (et == 0 && grp == 0) {
      HCF_PROF_FLUSH(PROF_E_TOTAL, PROF_E_COAL);
#ifdef HCF_TC_PROF_BUILD
      if (prof_on && blockIdx.x == 0) atomicAdd((unsigned long long*)p.prof + PROF_LAUNCHES, 1ull);
#endif
    }
  } else {
    // ===================== publisher (warp 10) =====================
    // per group pass: join the epilogue group of every tile on its named barrier (the tile's stores happen before it),
    // then ONE gpu-scope fence makes the whole pass visible and the tiles' counters are bumped with relaxed adds
    // (release pattern: fence + relaxed atomic; the consumer's side is relaxed loads + fence.acquire.gpu).  One fence
    // per pass instead of one release per tile: the gpu-scope fence waits for the SM's outstanding stores (~2k cycles)
    // and a single publisher serialises them.
    if (p.done != nullptr) {
      uint32_t t_it = 0;
      for (int q = 0; q < n_groups; ++q) {
        const int layer = q / p.n_phases, ph = q - layer * p.n_phases;
        const int gc = group_c(layer);
        const int ng = group_n(ph, gc);
        if (ng == 0) continue;
        for (int g = 0; g < ng; ++g, ++t_it) {
          const int grp = (int)(t_it & 1u);
          asm volatile("bar.sync %0, 160;" ::"r"(3 + grp) : "memory");
          if (lane == 0)   // the group may arrive for its next tile
            asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(pub_seq(grp)), "r"((t_it >> 1) + 1u) : "memory");
          __syncwarp();
        }
        if (lane == 0) {
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          for (int g = 0; g < ng; ++g)
            asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p.done + group_tile(ph, gc, g)), "r"(1) : "memory");
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace tc
}  // namespace hcf
