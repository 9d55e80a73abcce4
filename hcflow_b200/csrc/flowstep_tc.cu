// One fused kernel per FlowStep chain (sm_100a: tcgen05 / TMEM / TMA).
//
// A FlowStep (FlowStep.py:40-64) with an FCN coupling sub-net (Basic.py:426-447) is
//     h = conv3x3( relu(an2( conv1x1( relu(an1( conv3x3(z1) [+ W_u * u] )) )) ) ) * exp(3 logs)        (sub-net)
//     inverse:  z2 = z2 * exp(-ls(h)) - shift(h);  z = W^-1 z;  z = z * exp(-logs) - b                  (tail)
//     forward:  z2 = (z2 + shift(h)) * exp(ls(h)); logdet += sum ls;  [next step's ActNorm + W]          (tail)
// Round 1 ran the three sub-net convs as three work items of a chained conv launch: three inter-CTA dependency
// hops per FlowStep (each: epilogue -> global -> release -> poll -> TMA), which left the FlowStep chains at 7 % of
// the HBM roofline / 10 % of the tensor peak (latency-bound: VERDICT r1 item 5).  Here ONE work item = one FlowStep on
// one 16x8-pixel tile, with a recomputed halo, so that h1 and h2 never leave the SM and a FlowStep is ONE hop:
//
//   z1 tile   [20][12] pixels (2-pixel halo), fp16 [hi 16 ch | lo 16 ch] per pixel, ONE TMA tile load (OOB -> 0 =
//             the convs' zero padding) from a ping-pong staging buffer the previous step's tail wrote
//   conv1     3x3 over z1 on the [18] x pitch-12 region (216 pixels = 2 UMMA M tiles; tap (dy,dx) = the same tile
//             through a descriptor shifted by dy*12+dx rows): 9 taps x K=16 into TMEM
//   epilogue1 TMEM -> registers: + W_u*u addend (conditional steps), ActNorm, ReLU -> fp16 hi / lo -> shared memory,
//             written directly in the 128B-swizzled K-major layout the tensor core reads (the A operand of conv2)
//   conv2     1x1, K=64, same 2 M tiles            epilogue2: ActNorm, ReLU, zero outside the image -> the SAME
//             shared-memory tile in place = h2 with its 1-pixel halo
//   conv3     3x3 over h2 (pitch-12 halo tile, SBO = 12 rows) for the 16x8 output pixels -> h in TMEM
//   tail      thread = pixel: coupling, C x C mix, ActNorm on z in place (fp32, global), fp16 hi / lo of the new z1
//             into the OTHER staging buffer, per-image log-det by warp shuffle + one fp64 atomic per warp (forward)
//
// Operand split (x3 modes): a = hi + lo / 2048 on both operands, A_hi x [B_hi ; B_lo] (main | correction columns)
// and A_lo x B_hi into the correction columns, exactly as in conv_tc_kernel.cuh -- all three convs run split, because
// a coupling amplifies the sub-net's rounding (tests/golden/*_stress.pt).
// Warp roles: 0 = TMA producer (+ dependency wait), 1 = MMA issuer / TMEM owner, 2-5 = epilogues 1 and 2,
// 6-9 = tail.  The tail of item i overlaps conv1 / conv2 of item i+1 (separate TMEM columns, separate warps).
// Weights of a step (136 KB fp16 [hi ; lo] images) stay in shared memory for all items of that step on the CTA.
// The launch is cooperative (all CTAs co-resident): tiles of step s wait for their 3x3 tile neighbourhood of step
// s-1 through per-tile counters, as in the chained conv kernel.
#include "conv_tc_kernel.cuh"

namespace hcf {
namespace fs {

using namespace hcf::tc;

constexpr int PITCH = 12;                       // pixels per row of the z1 / h tiles
constexpr int Z1_ROWS = 20, H_ROWS = 18;
constexpr int Z1_BYTES = Z1_ROWS * PITCH * 128;           // 30720
constexpr int HPIX = H_ROWS * PITCH;                      // 216 rows of the h1 / h2 tile
constexpr int H_PLANE = HPIX * 128;                       // 27648 (= 27 KB, 1024-aligned)
constexpr int W1_BYTES = 3 * 128 * 128;                   // 3 tap blocks x [64 hi ; 64 lo] rows x 128 B
constexpr int W2_BYTES = 128 * 128;
constexpr int W3_MAX = 9 * 64 * 128;
constexpr int OFF_Z1 = 0;
constexpr int OFF_HHI = OFF_Z1 + Z1_BYTES;
constexpr int OFF_HLO = OFF_HHI + H_PLANE;
constexpr int OFF_W1 = OFF_HLO + H_PLANE;
constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
constexpr int OFF_W3 = OFF_W2 + W2_BYTES;
constexpr int OFF_TAB = OFF_W3 + W3_MAX;
constexpr int MAXC = 24;
constexpr int EPI_FLOATS = 320;                           // bias1 64 | scale1 64 | bias2 64 | scale2 64 | bias3 32 | scale3 32
constexpr int TAIL_FLOATS = MAXC * MAXC + 2 * MAXC;       // W | scale | bias
constexpr int TAB_FLOATS = EPI_FLOATS + TAIL_FLOATS;      // per step, in global memory
constexpr int OFF_BAR = OFF_TAB + TAB_FLOATS * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;          // + base alignment slack
static_assert(OFF_HHI % 1024 == 0 && OFF_HLO % 1024 == 0 && OFF_W1 % 1024 == 0 && OFF_W2 % 1024 == 0 && OFF_W3 % 1024 == 0,
              "operand tiles must be 1024-byte aligned (128B swizzle atoms)");
static_assert(SMEM_BYTES <= SMEM_LIMIT, "flowstep kernel: shared memory budget");
constexpr int THREADS = 448;    // producer, MMA issuer, 2 x 4 epilogue warps (one group per M tile), 4 tail warps

struct StepDesc {
  const __half* w1; const __half* w2; const __half* w3;
  const float* pre; int pre_ld;     // conditional steps: W_u * u, fp32 [B,H,W,>=64] (may be null)
  int has_w;                        // tail mixes with a C x C matrix (inverse: W^-1 of this step; forward: W of the NEXT step)
  int has_next;                     // forward: the next step's ActNorm + W are applied by this tail; both: z1 is staged
};

struct Params {
  int B, H, W, tiles_x, tiles_y, n_tiles, n_steps, n_items;
  int C, n_pass, N3, split, forward;
  int w3_bytes;
  float* z; int z_ld;
  int z_vec;                        // z rows allow 16-byte accesses over ceil4(C) channels
  __half* z16[2];
  const StepDesc* steps;
  const float* tabs;                // [n_steps][TAB_FLOATS]
  int* done;
  int* status;
  double* logdet;
  long long* prof;                  // HCF_TC_PROF=1 (prof build): cycles per role / wait class summed over CTAs
};

struct Maps2 { CUtensorMap m[2]; };

// mbarrier wait with a watchdog: a protocol bug must not hang the GPU (a kernel that never ends takes the box down);
// after ~4 s of polling the status word gets HCF_STATUS_DEP_TIMEOUT and the kernel traps
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity, int* status) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    if ((it & 1023u) == 1023u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000ll) {
        if (status) atomicOr(status, STATUS_DEP_TIMEOUT);
        __trap();
      }
    }
  }
}
#define mbar_wait(bar, parity) mbar_wait_wd((bar), (parity), p.status)

// 32 accumulator columns of this thread's TMEM lane, WITHOUT the wait (several loads share one tcgen05.wait::ld)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// two floats -> packed fp16 pair, round to nearest, finite saturation (lo half = first argument)
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ------------------------------------------------------------------------------------------------ per-pixel tails
// h[2j] = shift, h[2j+1] = scale of coupled channel n_pass + j (AffineCouplings.py:52-57, 78-84: h[:, 0::2], h[:, 1::2])
template <int C>
__device__ __forceinline__ void mix(const float* __restrict__ s_w, const float (&z)[C], float (&y)[C]) {
#pragma unroll
  for (int i = 0; i < C; ++i) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // independent accumulators: no C-long dependent FMA chain
    if (C % 4 == 0) {
#pragma unroll
      for (int j = 0; j < C; j += 4) {               // 16-byte broadcast loads of the row
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
}

// returns the pixel's log-det contribution (forward) or 0
template <int C>
__device__ __forceinline__ float tail_pixel(const float (&h)[32], const float (&zq)[MAXC], float* __restrict__ zp, bool forward,
                                            bool has_w, bool has_next, const float* __restrict__ s_w,
                                            const float* __restrict__ s_sc, const float* __restrict__ s_b,
                                            __half* __restrict__ z16p, bool split, bool vec) {
  constexpr int n_pass = C / 2;      // AffineCoupling: channels_for_nn = in_channels // 2 (AffineCouplings.py:17-18)
  float z[C];
#pragma unroll
  for (int i = 0; i < C; ++i) z[i] = zq[i];
  float lsum = 0.f;
#pragma unroll
  for (int i = 0; i < C; ++i) {
    if (i >= n_pass) {
      const int j = i - n_pass;
      const float shift = h[2 * j], scale = h[2 * j + 1];
      const float ls = coupling_logscale(scale);
      if (forward) {
        z[i] = (z[i] + shift) * expf(ls);
        lsum += ls;
      } else {
        z[i] = z[i] * expf(-ls) - shift;
      }
    }
  }
  float y[C];
  if (forward) {
    if (has_next) {   // ActNorm then W of the next step (ActNorms.py:66-69, Permutations.py:94-101)
#pragma unroll
      for (int i = 0; i < C; ++i) z[i] = (z[i] + s_b[i]) * s_sc[i];
      if (has_w) mix<C>(s_w, z, y);
      else {
#pragma unroll
        for (int i = 0; i < C; ++i) y[i] = z[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < C; ++i) y[i] = z[i];
    }
  } else {            // W^-1 then ActNorm^-1 of this step (Permutations.py:103-108, ActNorms.py:90-93)
    if (has_w) mix<C>(s_w, z, y);
    else {
#pragma unroll
      for (int i = 0; i < C; ++i) y[i] = z[i];
    }
#pragma unroll
    for (int i = 0; i < C; ++i) y[i] = y[i] * s_sc[i] - s_b[i];
  }
  if (vec) {          // 16-byte stores for the full groups of four channels, scalar stores for the remainder
#pragma unroll
    for (int i = 0; i + 3 < C; i += 4)
      __stcg(reinterpret_cast<float4*>(zp + i), make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]));
#pragma unroll
    for (int i = C / 4 * 4; i < C; ++i) __stcg(zp + i, y[i]);
  } else {
#pragma unroll
    for (int i = 0; i < C; ++i) __stcg(zp + i, y[i]);
  }
  if (z16p) {         // the next step's conv1 operand: [hi 16 | lo 16] halves per pixel, four 16-byte stores
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const float a = i < n_pass ? fminf(fmaxf(y[i < C ? i : 0], -65504.0f), 65504.0f) : 0.f;
      const float c = i + 1 < n_pass ? fminf(fmaxf(y[i + 1 < C ? i + 1 : 0], -65504.0f), 65504.0f) : 0.f;
      const __half2 hh = __floats2half2_rn(a, c);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn((a - hf.x) * 2048.0f, (c - hf.y) * 2048.0f);
      hi[i >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[i >> 1] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    uint4* d = reinterpret_cast<uint4*>(z16p);
    __stcg(d, make_uint4(hi[0], hi[1], hi[2], hi[3]));
    __stcg(d + 1, make_uint4(hi[4], hi[5], hi[6], hi[7]));
    if (split) {
      __stcg(d + 2, make_uint4(lo[0], lo[1], lo[2], lo[3]));
      __stcg(d + 3, make_uint4(lo[4], lo[5], lo[6], lo[7]));
    }
  }
  return lsum;
}

// in-kernel wait profile (HCF_BUILD_PROF=1 build + HCF_TC_PROF=1): where each role's cycles go, per work item
enum { FP_P_TOTAL = 0, FP_P_WEIGHTS, FP_P_DEPS, FP_P_Z1EMPTY, FP_M_TOTAL, FP_M_Z1FULL, FP_M_ISSUE1, FP_M_A2READY, FP_M_ISSUE2,
       FP_M_A3READY, FP_M_ACC3EMPTY, FP_M_ISSUE3, FP_M_WFULL, FP_E_TOTAL, FP_E_ACC1, FP_E_BODY1, FP_E_ACC2, FP_E_BODY2, FP_E_TABLES,
       FP_T_TOTAL, FP_T_DEPSEQ, FP_T_ACC3, FP_T_BODY, FP_T_PUBLISH, FP_LAUNCHES, FP_N };
#ifdef HCF_TC_PROF_BUILD
#define FS_T(var) const long long var = prof_on ? clock64() : 0ll
#define FS_ACC(slot, a, b) do { if (prof_on) pacc[slot] += (b) - (a); } while (0)
#define FS_FLUSH(lo, hi) do { if (prof_on) for (int i_ = (lo); i_ <= (hi); ++i_) \
    atomicAdd((unsigned long long*)p.prof + i_, (unsigned long long)pacc[i_]); } while (0)
#else
#define FS_T(var) do { } while (0)
#define FS_ACC(slot, a, b) do { } while (0)
#define FS_FLUSH(lo, hi) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(THREADS, 1) flowstep_kernel(const __grid_constant__ Maps2 maps, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar0 = sbase + OFF_BAR;
  // mbarriers
  const uint32_t z1_full = bar0, z1_empty = bar0 + 8;
  auto w_full = [&](int k) { return bar0 + 16u + 8u * k; };
  auto w_empty = [&](int k) { return bar0 + 40u + 8u * k; };
  // conv1 -> epilogue 1 -> conv2 -> epilogue 2 hand-offs are per M tile (conv2 is 1x1: tile m's conv2 needs tile m's h1 only)
  auto acc1_full = [&](int m) { return bar0 + 64u + 8u * m; };
  auto a2_ready = [&](int m) { return bar0 + 80u + 8u * m; };
  auto acc2_full = [&](int m) { return bar0 + 96u + 8u * m; };
  const uint32_t a3_ready = bar0 + 112, acc3_full = bar0 + 120, acc3_empty = bar0 + 128;
  const uint32_t dep_seq = bar0 + 136, tmem_slot = bar0 + 140;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[1]) : "memory");
    mbar_init(z1_full, 1); mbar_init(z1_empty, 1);
    for (int k = 0; k < 3; ++k) { mbar_init(w_full(k), 1); mbar_init(w_empty(k), 1); }
    for (int m = 0; m < 2; ++m) { mbar_init(acc1_full(m), 1); mbar_init(a2_ready(m), 128); mbar_init(acc2_full(m), 1); }
    mbar_init(a3_ready, 256);
    mbar_init(acc3_full, 1); mbar_init(acc3_empty, 128);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(dep_seq), "r"(0u) : "memory");
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
  auto item_at = [&](int q) -> int {
    const int item = (int)blockIdx.x + q * (int)gridDim.x;
    return item < p.n_items ? item : -1;
  };
  const bool split = p.split != 0;
  const uint32_t NB3 = (uint32_t)p.N3 * (split ? 2u : 1u);
#ifdef HCF_TC_PROF_BUILD
  const bool prof_on = p.prof != nullptr;
  long long pacc[FP_N];
#pragma unroll
  for (int i = 0; i < FP_N; ++i) pacc[i] = 0;
#endif

  if (warp == 0) {
    // =============================================================== producer
    if (lane == 0) {
      int cur_step = -1;
      uint32_t w_it = 0;       // number of weight (re)loads so far
      Deps deps;
      if (item_at(0) >= 0) {
        const int tile = item_at(0) % p.n_tiles;
        const int b = tile / per_img, r = tile % per_img;
        load_deps(deps, p.done, b * per_img, r / p.tiles_x, r % p.tiles_x, p.tiles_y, p.tiles_x);
      }
      FS_T(tp0);
      for (int seq = 0, item; (item = item_at(seq)) >= 0; ++seq) {
        const int step = item / p.n_tiles, tile = item - step * p.n_tiles;
        const int b = tile / per_img, r = tile % per_img;
        const int ty = r / p.tiles_x, tx = r % p.tiles_x;
        FS_T(tw0);
        if (step != cur_step) {
          // the step's weight images: issued BEFORE the dependency wait (they depend on nothing)
          cur_step = step;
          const StepDesc* S = p.steps + step;
          const void* src[3] = {ldg_ptr(&S->w1), ldg_ptr(&S->w2), ldg_ptr(&S->w3)};
          // (the W1 image always has the [hi ; lo] block geometry; one-pass plans just never read the lo rows)
          const uint32_t bytes[3] = {(uint32_t)W1_BYTES, (uint32_t)(split ? W2_BYTES : W2_BYTES / 2), (uint32_t)p.w3_bytes};
          const uint32_t dst[3] = {sbase + OFF_W1, sbase + OFF_W2, sbase + OFF_W3};
          for (int k = 0; k < 3; ++k) {
            mbar_wait(w_empty(k), (w_it & 1u) ^ 1u);
            mbar_expect_tx(w_full(k), bytes[k]);
            bulk_load(dst[k], src[k], bytes[k], w_full(k));
          }
          ++w_it;
        }
        FS_T(tw1);
        FS_ACC(FP_P_WEIGHTS, tw0, tw1);
        if (step > 0) {
          uint32_t spins = 0;
          while (!deps_ready(deps, step)) {
            __nanosleep(32);
            load_deps(deps, p.done, b * per_img, ty, tx, p.tiles_y, p.tiles_x);
            if (++spins > (1u << 24)) {
              if (p.status) atomicOr(p.status, STATUS_DEP_TIMEOUT);
              __trap();
            }
          }
          fence_acquire_gpu();
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        FS_T(tw2);
        FS_ACC(FP_P_DEPS, tw1, tw2);
        asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(dep_seq), "r"((uint32_t)seq + 1u) : "memory");
        const int nitem = item_at(seq + 1);
        if (nitem >= 0) {
          const int nt = nitem % p.n_tiles;
          const int nb = nt / per_img, nr = nt % per_img;
          load_deps(deps, p.done, nb * per_img, nr / p.tiles_x, nr % p.tiles_x, p.tiles_y, p.tiles_x);
        }
        FS_T(tw3);
        mbar_wait(z1_empty, ((uint32_t)seq & 1u) ^ 1u);
        FS_T(tw4);
        FS_ACC(FP_P_Z1EMPTY, tw3, tw4);
        mbar_expect_tx(z1_full, Z1_BYTES);
        tma_load_4d(sbase + OFF_Z1, &maps.m[step & 1], z1_full, 0, tx * TW - 2, ty * TH - 2, b);
      }
      FS_T(tp1);
      FS_ACC(FP_P_TOTAL, tp0, tp1);
      FS_FLUSH(FP_P_TOTAL, FP_P_Z1EMPTY);
    }
  } else if (warp == 1) {
    // =============================================================== MMA issuer
    const uint64_t dense = make_desc(0, 8u * 128u);               // 8-row groups back to back (linear pixel rows)
    const uint64_t pitch12 = make_desc(0, (uint32_t)PITCH * 128u); // conv3: row group g = output row g of the tile
    const uint32_t idesc128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc3m = (1u << 4) | ((NB3 >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc3c = (1u << 4) | (((uint32_t)p.N3 >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc12 = split ? idesc128 : idesc64;
    const uint32_t W1_BLOCK16 = 128u * 8u;                         // one tap block of W1 ([hi 64 ; lo 64] rows) in 16-byte units
    int cur_step = -1;
    uint32_t w_it = 0;
    FS_T(tm0);
    for (int seq = 0, item; (item = item_at(seq)) >= 0; ++seq) {
      const int step = item / p.n_tiles;
      const int nitem = item_at(seq + 1);
      const bool last_of_step = nitem < 0 || nitem / p.n_tiles != step;
      const bool new_step = step != cur_step;
      if (new_step) { cur_step = step; ++w_it; }
      const uint32_t par = (uint32_t)seq & 1u, wpar = (w_it - 1u) & 1u;
      // ---------------- conv1: 9 taps x K=16 (hi) [+ lo] on two M tiles -> cols [m*128, m*128+128)
      FS_T(ta0);
      mbar_wait(z1_full, par);
      FS_T(ta1);
      if (new_step) mbar_wait(w_full(0), wpar);
      FS_T(ta2);
      FS_ACC(FP_M_Z1FULL, ta0, ta1);
      FS_ACC(FP_M_WFULL, ta1, ta2);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a0 = dense + ((sbase + OFF_Z1) >> 4);
        const uint64_t b0 = dense + ((sbase + OFF_W1) >> 4);
#pragma unroll 1
        for (int m = 0; m < 2; ++m) {
          const uint32_t d = tmem_base + (uint32_t)m * 128u;
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap - dy * 3;
            const uint64_t bd = b0 + (uint32_t)(tap >> 2) * W1_BLOCK16 + (uint32_t)(tap & 3) * 2u;
            const uint64_t ad = a0 + (uint32_t)((m * 128 + dy * PITCH + dx) * 8);
            umma_f16(d, ad, bd, idesc12, tap > 0 ? 1u : 0u);
            if (split) umma_f16(d + 64u, ad + 2u, bd, idesc64, 1u);      // A_lo (bytes 32..63 of the row) x B_hi
          }
          umma_commit(acc1_full(m));                                     // epilogue 1 of tile 0 runs under tile 1's MMAs
        }
        umma_commit(z1_empty);
        if (last_of_step) umma_commit(w_empty(0));
      }
      __syncwarp();
      FS_T(ta3);
      FS_ACC(FP_M_ISSUE1, ta2, ta3);
      // ---------------- conv2: 1x1, K = 64, tile by tile as its h1 rows arrive
      if (new_step) mbar_wait(w_full(1), wpar);
#pragma unroll 1
      for (int m = 0; m < 2; ++m) {
        FS_T(tq0);
        mbar_wait(a2_ready(m), par);
        FS_T(tq1);
        FS_ACC(FP_M_A2READY, tq0, tq1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ah = dense + ((sbase + OFF_HHI) >> 4), al = dense + ((sbase + OFF_HLO) >> 4);
          const uint64_t b0 = dense + ((sbase + OFF_W2) >> 4);
          const uint32_t d = tmem_base + (uint32_t)m * 128u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t off = (uint32_t)(m * 128 * 8 + k * 2);
            umma_f16(d, ah + off, b0 + 2u * k, idesc12, k > 0 ? 1u : 0u);
            if (split) umma_f16(d + 64u, al + off, b0 + 2u * k, idesc64, 1u);
          }
          umma_commit(acc2_full(m));
          if (m == 1 && last_of_step) umma_commit(w_empty(1));
        }
        __syncwarp();
        FS_T(tq2);
        FS_ACC(FP_M_ISSUE2, tq1, tq2);
      }
      FS_T(ta5);
      // ---------------- conv3: 3x3 over the h2 halo tile -> cols [256, 256 + NB3)
      mbar_wait(a3_ready, par);
      FS_T(ta6);
      FS_ACC(FP_M_A3READY, ta5, ta6);
      mbar_wait(acc3_empty, par ^ 1u);
      FS_T(ta7);
      FS_ACC(FP_M_ACC3EMPTY, ta6, ta7);
      if (new_step) mbar_wait(w_full(2), wpar);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t ah = pitch12 + ((sbase + OFF_HHI) >> 4), al = pitch12 + ((sbase + OFF_HLO) >> 4);
        const uint64_t b0 = dense + ((sbase + OFF_W3) >> 4);
        const uint32_t d = tmem_base + 256u;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3, dx = tap - dy * 3;
          const uint32_t aoff = (uint32_t)((dy * PITCH + dx) * 8);
          const uint64_t bt = b0 + (uint32_t)tap * NB3 * 8u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16(d, ah + aoff + 2u * k, bt + 2u * k, idesc3m, (tap | k) ? 1u : 0u);
            if (split) umma_f16(d + (uint32_t)p.N3, al + aoff + 2u * k, bt + 2u * k, idesc3c, 1u);
          }
        }
        umma_commit(acc3_full);
        if (last_of_step) umma_commit(w_empty(2));
      }
      __syncwarp();
      FS_T(ta8);
      FS_ACC(FP_M_ISSUE3, ta7, ta8);
    }
    FS_T(tm1);
    FS_ACC(FP_M_TOTAL, tm0, tm1);
    if (lane == 0) FS_FLUSH(FP_M_TOTAL, FP_M_WFULL);
  } else if (warp < 10) {
    // =============================================================== epilogues 1 and 2 (thread = linear tile pixel;
    // warps 2-5 drain M tile 0, warps 6-9 M tile 1)
    const int q4 = warp & 3;                       // TMEM lane quarter of this warp
    const int t = q4 * 32 + lane;                  // row within an M tile
    const int m = (warp - 2) >> 2;                 // M tile of this warp's group
    const int q = m * 128 + t;                     // linear pixel of the [18] x pitch-12 region
    const bool store = q < HPIX;
    const bool warp_live = (m * 128 + q4 * 32) < HPIX;   // (the last warp of M tile 1 holds padding rows only)
    float* s_epi = reinterpret_cast<float*>(gbase + OFF_TAB);
    uint8_t* rowh = gbase + OFF_HHI + q * 128;
    uint8_t* rowl = gbase + OFF_HLO + q * 128;
    const int rr = q / PITCH, cc = q - rr * PITCH;
    int cur_step = -1;
    const float* pre = nullptr;
    int pre_ld = 0;
    FS_T(te0);
    for (int seq = 0, item; (item = item_at(seq)) >= 0; ++seq) {
      const int step = item / p.n_tiles, tile = item - step * p.n_tiles;
      const int b = tile / per_img, r = tile % per_img;
      const int y0 = (r / p.tiles_x) * TH, x0 = (r % p.tiles_x) * TW;
      const uint32_t par = (uint32_t)seq & 1u;
      FS_T(tb0);
      if (step != cur_step) {
        cur_step = step;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int et = threadIdx.x - 64;
        {   // bias slots hold bias * scale: the epilogue is one FMA per channel, (v + b) * s = fma(v, s, b * s)
          const float* tb = p.tabs + (size_t)step * TAB_FLOATS;
          const float tv = __ldg(tb + et);
          s_epi[et] = (et & 64) ? tv : tv * __ldg(tb + et + 64);
        }
        pre = ldg_ptr(&(p.steps + step)->pre);
        pre_ld = __ldg(&(p.steps + step)->pre_ld);
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      FS_T(tb1);
      FS_ACC(FP_E_TABLES, tb0, tb1);
      // region pixel (rr, cc) -> image pixel (y0 - 1 + rr, x0 - 1 + cc)
      const int gy = y0 - 1 + rr, gx = x0 - 1 + cc;
      const bool inimg = store && cc < 10 && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
      const float* prep = (pre != nullptr && inimg) ? pre + (size_t)((b * p.H + gy) * p.W + gx) * pre_ld : nullptr;
      // the W_u * u addend of conv1 (conditional steps): the first half of the row is fetched BEFORE the accumulator
      // wait, the second half while the first is processed (thread = pixel: 256 contiguous bytes per thread)
      float4 pa[8];
      if (prep) {
#pragma unroll
        for (int j = 0; j < 8; ++j) pa[j] = __ldg(reinterpret_cast<const float4*>(prep) + j);
      }
#pragma unroll 1
      for (int stage = 0; stage < 2; ++stage) {
        FS_T(tb2);
        mbar_wait(stage == 0 ? acc1_full(m) : acc2_full(m), par);
        FS_T(tb3);
        FS_ACC(stage == 0 ? FP_E_ACC1 : FP_E_ACC2, tb2, tb3);
        tc_fence_after();
        const float* s_bias = s_epi + stage * 128;
        const float* s_scale = s_bias + 64;
        const bool keep = stage == 0 ? true : inimg;       // h2 outside the image is conv3's zero padding
        const bool add = stage == 0 && prep != nullptr;
        if (warp_live) {
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {                // 32 channels at a time: ONE TMEM round trip
            float v[32];
            {
              const uint32_t tcol = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(m * 128 + half * 32);
              uint32_t rv[32], rl[32];
              tmem_ld32_nowait(tcol, rv);
              if (split) tmem_ld32_nowait(tcol + 64u, rl);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int j = 0; j < 32; ++j)
                v[j] = split ? fmaf(__uint_as_float(rl[j]), 1.0f / 2048.0f, __uint_as_float(rv[j])) : __uint_as_float(rv[j]);
            }
            if (add) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[4 * j] += pa[j].x; v[4 * j + 1] += pa[j].y; v[4 * j + 2] += pa[j].z; v[4 * j + 3] += pa[j].w;
              }
              if (half == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) pa[j] = __ldg(reinterpret_cast<const float4*>(prep) + 8 + j);
              }
            }
            if (keep) {
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8) {                  // one 16-byte chunk (8 channels) per plane at a time
                const float4* sb4 = reinterpret_cast<const float4*>(s_bias + half * 32 + c8 * 8);
                const float4* ss4 = reinterpret_cast<const float4*>(s_scale + half * 32 + c8 * 8);
                const float4 b0 = sb4[0], b1 = sb4[1], s0 = ss4[0], s1 = ss4[1];
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                uint32_t hi[4], lo4[4];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                  const float a = fmaxf(fmaf(v[c8 * 8 + j], sc[j], bb[j]), 0.f);
                  const float c = fmaxf(fmaf(v[c8 * 8 + j + 1], sc[j + 1], bb[j + 1]), 0.f);
                  const uint32_t hh = cvt_f16x2_sat(a, c);     // saturates at +-65504 (fp16 range guard)
                  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hh));
                  const __half2 ll = __floats2half2_rn((a - hf.x) * 2048.0f, (c - hf.y) * 2048.0f);
                  hi[j >> 1] = hh;
                  lo4[j >> 1] = *reinterpret_cast<const uint32_t*>(&ll);
                }
                if (store) {
                  // 128B swizzle: 16-byte chunk index XOR (row & 7)
                  const int co = (((half * 4 + c8) ^ (q & 7))) * 16;
                  *reinterpret_cast<uint4*>(rowh + co) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                  if (split) *reinterpret_cast<uint4*>(rowl + co) = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
                }
              }
            } else if (store) {                                 // h2 outside the image: conv3's zero padding
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8) {
                const int co = (((half * 4 + c8) ^ (q & 7))) * 16;
                *reinterpret_cast<uint4*>(rowh + co) = make_uint4(0u, 0u, 0u, 0u);
                if (split) *reinterpret_cast<uint4*>(rowl + co) = make_uint4(0u, 0u, 0u, 0u);
              }
            }
          }
        }
        fence_async_smem();          // generic-proxy stores -> visible to the tensor core (async proxy)
        tc_fence_before();
        mbar_arrive(stage == 0 ? a2_ready(m) : a3_ready);
        FS_T(tb4);
        FS_ACC(stage == 0 ? FP_E_BODY1 : FP_E_BODY2, tb3, tb4);
      }
    }
    FS_T(te1);
    FS_ACC(FP_E_TOTAL, te0, te1);
    if (threadIdx.x == 64) FS_FLUSH(FP_E_TOTAL, FP_E_TABLES);
  } else {
    // =============================================================== tail (thread = output pixel)
    const int q4 = warp & 3;
    const int t = q4 * 32 + lane;
    const int oy = t >> 3, ox = t & 7;
    float* s_tab = reinterpret_cast<float*>(gbase + OFF_TAB) + 256;   // bias3 32 | scale3 32 | W | sc | b
    float* s_w = s_tab + 64;
    float* s_sc = s_w + MAXC * MAXC;
    float* s_b = s_sc + MAXC;
    int cur_step = -1, has_w = 0, has_next = 0;
    const int C = p.C;
    FS_T(tt0);
    for (int seq = 0, item; (item = item_at(seq)) >= 0; ++seq) {
      const int step = item / p.n_tiles, tile = item - step * p.n_tiles;
      const int b = tile / per_img, r = tile % per_img;
      const int gy = (r / p.tiles_x) * TH + oy, gx = (r % p.tiles_x) * TW + ox;
      const bool in = gy < p.H && gx < p.W;
      const uint32_t par = (uint32_t)seq & 1u;
      if (step != cur_step) {
        cur_step = step;
        asm volatile("bar.sync 2, 128;" ::: "memory");
        const int et = threadIdx.x - 320;
        const float* src = p.tabs + (size_t)step * TAB_FLOATS + 256;
        for (int i = et; i < 64 + TAIL_FLOATS; i += 128) s_tab[i] = __ldg(src + i);
        has_w = __ldg(&(p.steps + step)->has_w);
        has_next = __ldg(&(p.steps + step)->has_next);
        asm volatile("bar.sync 2, 128;" ::: "memory");
      }
      // the producer acquired this item's inputs (dependency counters + fence) before conv1 could run; z of this tile
      // was last written by this tile's previous step, which is ordered before that acquire.  z is fetched now, long
      // before the accumulator is ready
      FS_T(tc0);
      {
        uint32_t seen;
        do {
          asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(seen) : "r"(dep_seq) : "memory");
        } while (seen <= (uint32_t)seq);
      }
      FS_T(tc1);
      FS_ACC(FP_T_DEPSEQ, tc0, tc1);
      const uint32_t pix = in ? (uint32_t)((b * p.H + gy) * p.W + gx) : 0u;
      float* zp = p.z + (size_t)pix * p.z_ld;
      // z of this pixel: 16-byte accesses when the view allows it (a warp's scalar access to 32 pixel rows costs one
      // sector per lane and channel: measured, the tail was bound by exactly that -- 48 scattered stores per pixel)
      const bool vec = p.z_vec != 0;
      float zq[MAXC];
      if (vec) {
#pragma unroll
        for (int i = 0; i < MAXC; i += 4) {
          float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (in && i < C) t4 = __ldcg(reinterpret_cast<const float4*>(zp + i));
          zq[i] = t4.x; zq[i + 1] = t4.y; zq[i + 2] = t4.z; zq[i + 3] = t4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < MAXC; ++i) zq[i] = (in && i < C) ? __ldcg(zp + i) : 0.f;
      }
      FS_T(tc2);
      mbar_wait(acc3_full, par);
      FS_T(tc3);
      FS_ACC(FP_T_ACC3, tc2, tc3);
      tc_fence_after();
      float h[32];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        if (hh * 16 < p.N3) {
          float v[16];
          const uint32_t tcol = tmem_base + ((uint32_t)(q4 * 32) << 16) + 256u + (uint32_t)(hh * 16);
          tmem_ld16(tcol, v);
          if (split) {
            float lo[16];
            tmem_ld16(tcol + (uint32_t)p.N3, lo);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaf(lo[j], 1.0f / 2048.0f, v[j]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) h[hh * 16 + j] = (v[j] + s_tab[hh * 16 + j]) * s_tab[32 + hh * 16 + j];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) h[hh * 16 + j] = 0.f;
        }
      }
      tc_fence_before();
      mbar_arrive(acc3_empty);
      float lsum = 0.f;
      if (in) {
        __half* z16p = has_next ? p.z16[(step + 1) & 1] + (size_t)pix * 32 : nullptr;
        const bool fw = p.forward != 0, hw = has_w != 0, hn = has_next != 0;
        switch (C) {
          case 6: lsum = tail_pixel<6>(h, zq, zp, fw, hw, hn, s_w, s_sc, s_b, z16p, split, vec); break;
          case 12: lsum = tail_pixel<12>(h, zq, zp, fw, hw, hn, s_w, s_sc, s_b, z16p, split, vec); break;
          case 21: lsum = tail_pixel<21>(h, zq, zp, fw, hw, hn, s_w, s_sc, s_b, z16p, split, vec); break;
          case 24: lsum = tail_pixel<24>(h, zq, zp, fw, hw, hn, s_w, s_sc, s_b, z16p, split, vec); break;
          default: break;   // (the host only creates plans for these channel counts)
        }
      }
      if (p.forward && p.logdet) {   // per-image log-det: the tile lies in ONE image -> warp shuffle + one atomic per warp
        double s = (double)lsum;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) atomicAdd(p.logdet + b, s);
      }
      // publish: every tail thread's stores happen before the barrier; one thread releases them at gpu scope
      FS_T(tc4);
      FS_ACC(FP_T_BODY, tc3, tc4);
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (t == 0) red_release_add(p.done + tile, 1);
      FS_T(tc5);
      FS_ACC(FP_T_PUBLISH, tc4, tc5);
    }
    FS_T(tt1);
    FS_ACC(FP_T_TOTAL, tt0, tt1);
    if (t == 0) {
      FS_FLUSH(FP_T_TOTAL, FP_T_PUBLISH);
#ifdef HCF_TC_PROF_BUILD
      if (prof_on && blockIdx.x == 0) atomicAdd((unsigned long long*)p.prof + FP_LAUNCHES, 1ull);
#endif
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// z[:, :n_pass] (fp32 NHWC view) -> staging rows [hi 16 | lo 16] fp16 (the first step's conv1 operand)
__global__ void stage_z1_kernel(const float* __restrict__ z, int z_ld, int n_pass, long long npix, __half* __restrict__ z16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * 16) return;
  const long long pix = i >> 4;
  const int c = (int)(i & 15);
  float v = c < n_pass ? z[pix * z_ld + c] : 0.f;
  v = fminf(fmaxf(v, -65504.0f), 65504.0f);
  const __half hh = __float2half_rn(v);
  z16[pix * 32 + c] = hh;
  z16[pix * 32 + 16 + c] = __float2half_rn((v - __half2float(hh)) * 2048.0f);
}

}  // namespace fs
}  // namespace hcf

// ================================================================================================ host side / C ABI
struct hcf_flowstep_plan {
  hcf::fs::Maps2 maps;
  hcf::fs::Params p;
  hcf::fs::StepDesc* d_steps;
  float* d_tabs;
  std::vector<hcf_flowstep>* src;     // the caller's per-step pointers (for refresh)
  long long* d_prof;
  dim3 grid;
};

// conv1's z1 part: w [64][n_pass][3][3] fp32 (host) -> 3 tap blocks x rows [hi 64 ; lo 64] x 64 fp16; tap t sits in
// block t / 4 at channel slot (t % 4) * 16 (K = 16 per tap: the z1 staging rows carry 16 channels), pre-swizzled
extern "C" int64_t hcf_flowstep_w1_bytes(void) { return hcf::fs::W1_BYTES; }
extern "C" int hcf_flowstep_pack_w1(const float* w, int32_t n_pass, void* image) {
  using namespace hcf;
  HCF_REQUIRE(w && image && n_pass >= 1 && n_pass <= 16, "flowstep_pack_w1: bad args");
  memset(image, 0, fs::W1_BYTES);
  __half* img = reinterpret_cast<__half*>(image);
  for (int tap = 0; tap < 9; ++tap)
    for (int part = 0; part < 2; ++part)
      for (int n = 0; n < 64; ++n)
        for (int c = 0; c < n_pass; ++c) {
          const float v = w[((size_t)n * n_pass + c) * 9 + tap];
          const __half hi = __float2half_rn(v);
          const __half val = part == 0 ? hi : __float2half_rn((v - __half2float(hi)) * 2048.0f);
          const int row = part * 64 + n;
          const int j = (tap & 3) * 16 + c;                 // channel slot within the 64-wide row
          const int chunk = (j / 8) ^ (row & 7);
          img[((size_t)(tap >> 2) * 128 + row) * 64 + chunk * 8 + (j & 7)] = val;
        }
  return 0;
}

extern "C" int hcf_flowstep_stage_z1(const float* z, int32_t z_ld, int32_t n_pass, int64_t npix, void* z16, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(z && z16 && n_pass >= 1 && n_pass <= 16 && npix > 0, "flowstep_stage_z1: bad args");
  const long long n = npix * 16;
  fs::stage_z1_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(z, z_ld, n_pass, npix,
                                                                                   reinterpret_cast<__half*>(z16));
  return finish_launch("hcf_flowstep_stage_z1");
}

extern "C" int hcf_flowstep_chain_refresh(hcf_flowstep_plan* pl, void* stream);
extern "C" void hcf_flowstep_chain_destroy(hcf_flowstep_plan* pl);

extern "C" int hcf_flowstep_chain_create(const hcf_flowstep_chain_args* a, hcf_flowstep_plan** out) {
  using namespace hcf;
  HCF_REQUIRE(out != nullptr, "flowstep_chain: null out");
  *out = nullptr;
  HCF_REQUIRE(a && a->steps && a->n_steps >= 1 && a->B >= 1 && a->H >= 1 && a->W >= 1, "flowstep_chain: bad args");
  if (!(a->C == 6 || a->C == 12 || a->C == 21 || a->C == 24) || a->n_pass != a->C / 2 || 2 * (a->C - a->n_pass) > 32) {
    set_error("flowstep_chain: C = %d / n_pass = %d not supported (C in {6,12,21,24}, n_pass = C / 2)", a->C, a->n_pass);
    return HCF_ENOTSUP;
  }
  HCF_REQUIRE(a->z && a->z_ld >= a->C && a->z16_a && a->z16_b && aligned16(a->z16_a) && aligned16(a->z16_b) && a->done,
              "flowstep_chain: z / staging / counters");
  HCF_REQUIRE((uint64_t)a->B * a->H * a->W * (uint64_t)(a->z_ld > 64 ? a->z_ld : 64) < (1ull << 32),
              "flowstep_chain: buffer too large for 32-bit element offsets");
  for (int i = 0; i < a->n_steps; ++i) {
    const hcf_flowstep& s = a->steps[i];
    HCF_REQUIRE(s.w1 && s.w2 && s.w3 && aligned16(s.w1) && aligned16(s.w2) && aligned16(s.w3) && s.bias1 && s.scale1 &&
                    s.bias2 && s.scale2 && s.bias3 && s.scale3 && s.an_scale && s.an_bias,
                "flowstep_chain: step %d: missing weights / tables", i);
    HCF_REQUIRE(!s.pre || (aligned16(s.pre) && s.pre_ld % 4 == 0 && s.pre_ld >= 64), "flowstep_chain: step %d: pre view", i);
  }
  tc::EncodeTiledFn enc = tc::get_encode();
  HCF_REQUIRE(enc != nullptr, "flowstep_chain: cuTensorMapEncodeTiled entry point not found");
  hcf_flowstep_plan* pl = new hcf_flowstep_plan();
  memset(pl, 0, sizeof(*pl));
  fs::Params& p = pl->p;
  p.B = a->B; p.H = a->H; p.W = a->W;
  p.tiles_x = ceil_div(a->W, tc::TW); p.tiles_y = ceil_div(a->H, tc::TH);
  p.n_tiles = p.tiles_x * p.tiles_y * a->B;
  p.n_steps = a->n_steps;
  p.n_items = p.n_tiles * a->n_steps;
  p.C = a->C; p.n_pass = a->n_pass;
  p.N3 = tc::n_for(2 * (a->C - a->n_pass));
  p.split = a->split ? 1 : 0;
  p.forward = a->forward ? 1 : 0;
  p.w3_bytes = 9 * p.N3 * (p.split ? 2 : 1) * 128;
  p.z = a->z; p.z_ld = a->z_ld;
  p.z_vec = (aligned16(a->z) && a->z_ld % 4 == 0 && (a->C + 3) / 4 * 4 <= a->z_ld) ? 1 : 0;
  p.z16[0] = reinterpret_cast<__half*>(a->z16_a);
  p.z16[1] = reinterpret_cast<__half*>(a->z16_b);
  p.done = a->done;
  p.logdet = a->logdet;
  const cuuint32_t box[4] = {64, (cuuint32_t)fs::PITCH, (cuuint32_t)fs::Z1_ROWS, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 2; ++i) {
    const cuuint64_t dims[4] = {32, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->B};
    const cuuint64_t strides[3] = {64, (cuuint64_t)64 * a->W, (cuuint64_t)64 * a->W * a->H};
    CUresult r = enc(&pl->maps.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, p.z16[i], dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete pl;
      set_error("flowstep_chain: cuTensorMapEncodeTiled failed with %d", (int)r);
      return HCF_EINVAL;
    }
  }
  pl->src = new std::vector<hcf_flowstep>(a->steps, a->steps + a->n_steps);
  std::vector<fs::StepDesc> sd(a->n_steps);
  for (int i = 0; i < a->n_steps; ++i) {
    const hcf_flowstep& s = a->steps[i];
    sd[i].w1 = reinterpret_cast<const __half*>(s.w1);
    sd[i].w2 = reinterpret_cast<const __half*>(s.w2);
    sd[i].w3 = reinterpret_cast<const __half*>(s.w3);
    sd[i].pre = s.pre; sd[i].pre_ld = s.pre_ld;
    if (p.forward) {   // the tail applies the NEXT step's ActNorm + W
      sd[i].has_next = i + 1 < a->n_steps ? 1 : 0;
      sd[i].has_w = (i + 1 < a->n_steps && a->steps[i + 1].w) ? 1 : 0;
    } else {
      sd[i].has_next = i + 1 < a->n_steps ? 1 : 0;
      sd[i].has_w = s.w ? 1 : 0;
    }
  }
  cudaError_t e = cudaMalloc(&pl->d_steps, sizeof(fs::StepDesc) * a->n_steps);
  if (e == cudaSuccess) e = cudaMemcpy(pl->d_steps, sd.data(), sizeof(fs::StepDesc) * a->n_steps, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&pl->d_tabs, sizeof(float) * fs::TAB_FLOATS * a->n_steps);
  if (e == cudaSuccess) e = cudaMemset(pl->d_tabs, 0, sizeof(float) * fs::TAB_FLOATS * a->n_steps);
  if (e != cudaSuccess) {
    set_error("flowstep_chain: table upload: %s", cudaGetErrorString(e));
    hcf_flowstep_chain_destroy(pl);
    return (int)e;
  }
  p.steps = pl->d_steps;
  p.tabs = pl->d_tabs;
  if (hcf_flowstep_chain_refresh(pl, nullptr) != 0) {
    hcf_flowstep_chain_destroy(pl);
    return HCF_EINVAL;
  }
  if (getenv("HCF_TC_PROF")) {
    if (cudaMalloc(&pl->d_prof, sizeof(long long) * fs::FP_N) == cudaSuccess) cudaMemset(pl->d_prof, 0, sizeof(long long) * fs::FP_N);
    p.prof = pl->d_prof;
  }
  const int sms = tc::num_sms();
  pl->grid = dim3((unsigned)(p.n_tiles < sms ? p.n_tiles : sms));
  e = cudaFuncSetAttribute(reinterpret_cast<const void*>(fs::flowstep_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize,
                           fs::SMEM_BYTES);
  int per_sm = 0;
  if (e == cudaSuccess)
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinterpret_cast<const void*>(fs::flowstep_kernel),
                                                      fs::THREADS, fs::SMEM_BYTES);
  if (e != cudaSuccess || (long)per_sm * sms < (long)pl->grid.x) {
    set_error("flowstep_chain: %d CTAs cannot be co-resident (%d per SM): %s", (int)pl->grid.x, per_sm,
              cudaGetErrorString(e));
    hcf_flowstep_chain_destroy(pl);
    return HCF_ENOTSUP;
  }
  *out = pl;
  return 0;
}

// (re)gathers the per-step tables from the caller's device vectors: bias / scale of the three convs, and the tail's
// C x C matrix + ActNorm vectors (inverse: this step's W^-1, exp(-logs), bias; forward: the NEXT step's W, exp(logs), bias)
extern "C" int hcf_flowstep_chain_refresh(hcf_flowstep_plan* pl, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(pl && pl->d_tabs && pl->src, "flowstep_refresh: bad plan");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = pl->p.n_steps, C = pl->p.C, N3 = pl->p.N3;
  auto cp = [&](float* dst, const float* src, int count) {
    return src ? cudaMemcpyAsync(dst, src, sizeof(float) * count, cudaMemcpyDeviceToDevice, st) : cudaSuccess;
  };
  for (int i = 0; i < n; ++i) {
    const hcf_flowstep& s = (*pl->src)[i];
    float* t = pl->d_tabs + (size_t)i * fs::TAB_FLOATS;
    cudaError_t e = cp(t, s.bias1, 64);
    if (e == cudaSuccess) e = cp(t + 64, s.scale1, 64);
    if (e == cudaSuccess) e = cp(t + 128, s.bias2, 64);
    if (e == cudaSuccess) e = cp(t + 192, s.scale2, 64);
    if (e == cudaSuccess) e = cp(t + 256, s.bias3, N3);
    if (e == cudaSuccess) e = cp(t + 288, s.scale3, N3);
    const hcf_flowstep* tl = pl->p.forward ? (i + 1 < n ? &(*pl->src)[i + 1] : nullptr) : &s;
    if (tl) {
      if (e == cudaSuccess && tl->w) e = cp(t + 320, tl->w, C * C);
      if (e == cudaSuccess) e = cp(t + 320 + fs::MAXC * fs::MAXC, tl->an_scale, C);
      if (e == cudaSuccess) e = cp(t + 320 + fs::MAXC * fs::MAXC + fs::MAXC, tl->an_bias, C);
    }
    if (e != cudaSuccess) {
      set_error("flowstep_refresh: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  return 0;
}

extern "C" int hcf_flowstep_chain_set_status(hcf_flowstep_plan* pl, int32_t* status) {
  using namespace hcf;
  HCF_REQUIRE(pl != nullptr, "flowstep_set_status: null plan");
  pl->p.status = status;
  return 0;
}

// `done` must be zeroed before every run (hcf_flowstep_chain_args.done); the z1 staging buffer A must hold the fp16
// image of z[:, :n_pass] (hcf_flowstep_stage_z1); forward chains expect the first step's ActNorm + W already applied.
extern "C" int hcf_flowstep_chain_run(const hcf_flowstep_plan* pl, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(pl != nullptr, "flowstep_run: null plan");
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = pl->grid;
  cfg.blockDim = dim3((unsigned)fs::THREADS);
  cfg.dynamicSmemBytes = fs::SMEM_BYTES;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // the CTAs wait on each other's tiles: must be co-resident
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, fs::flowstep_kernel, pl->maps, pl->p);
  if (e != cudaSuccess) {
    set_error("hcf_flowstep_chain_run: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return finish_launch("hcf_flowstep_chain_run");
}

extern "C" void hcf_flowstep_chain_destroy(hcf_flowstep_plan* pl) {
  if (!pl) return;
  if (pl->d_prof) {   // HCF_TC_PROF=1: average cycles per work item, by role and wait class
    long long h[hcf::fs::FP_N];
    cudaDeviceSynchronize();
    if (cudaMemcpy(h, pl->d_prof, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
      static const char* names[hcf::fs::FP_N] = {"P.total", "P.weights", "P.deps", "P.z1_empty", "M.total", "M.z1_full", "M.issue1",
                                                 "M.a2_ready", "M.issue2", "M.a3_ready", "M.acc3_empty", "M.issue3", "M.w_full",
                                                 "E.total", "E.acc1", "E.body1", "E.acc2", "E.body2", "E.tables", "T.total",
                                                 "T.dep_seq", "T.acc3", "T.body", "T.publish", "launches"};
      const double items = (double)pl->p.n_items / pl->grid.x;
      const double launches = h[hcf::fs::FP_LAUNCHES] > 0 ? (double)h[hcf::fs::FP_LAUNCHES] : 1.0;
      fprintf(stderr, "[hcf prof] flowstep chain steps=%d tiles=%d C=%d fwd=%d items/CTA=%.1f launches=%.0f; cycles per item:",
              pl->p.n_steps, pl->p.n_tiles, pl->p.C, pl->p.forward, items, launches);
      for (int i = 0; i < hcf::fs::FP_LAUNCHES; ++i)
        fprintf(stderr, " %s=%.0f", names[i], (double)h[i] / pl->grid.x / launches / items);
      fprintf(stderr, "\n");
    }
    cudaFree(pl->d_prof);
  }
  if (pl->d_steps) cudaFree(pl->d_steps);
  if (pl->d_tabs) cudaFree(pl->d_tabs);
  delete pl->src;
  delete pl;
}
