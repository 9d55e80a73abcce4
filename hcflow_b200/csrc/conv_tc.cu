// Host side of the tcgen05 convolution: shared-memory ring sizing, TMA tensor maps (shared by the channel-slice views
// of one buffer), per-layer tables of a chained launch, weight-image packing, and the C ABI entry points
// (include/hcflow_b200.h).  The kernel and its device helpers are in conv_tc_kernel.cuh.
#include "conv_tc_kernel.cuh"
#include "conv_ws_kernel.cuh"

namespace hcf {
namespace tc {

// ------------------------------------------------------------------ host side
// ring depths and B slot granularity that fit in shared memory; false if nothing fits
static bool pick_rings(int mt, int passes, int ks, int NB, int extra, int* sa, int* sb, int* slab_taps, size_t* smem) {
  const int a_stage = a_part(mt, ks) * (passes == 3 ? 2 : 1);
  const int budget = SMEM_LIMIT - 1024 - TAIL_BYTES - extra;
  const int tap = NB * ROW_BYTES;
  const int taps = ks * ks;
  auto done = [&](int a, int b, int st) {
    *sa = a; *sb = b; *slab_taps = st;
    *smem = 1024 + (size_t)a * a_stage + (size_t)b * st * tap + TAIL_BYTES + extra;
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

typedef void (*KernelFn)(const Maps, const Params);

static KernelFn pick_kernel(int mt, int passes, int ks, bool f16, bool step, bool direct) {
  if (f16 && !step && ks == 3 && direct)   // a layer of the chain qualifies for the direct hi-plane stores
    return passes == 3 ? conv_tc_kernel<1, 3, 3, true, false, true> : conv_tc_kernel<1, 1, 3, true, false, true>;
  if (step) {   // fused FlowStep layers: 3x3 chains, MT = 1 (chain_create enforces both)
    if (f16) return passes == 3 ? conv_tc_kernel<1, 3, 3, true, true> : conv_tc_kernel<1, 1, 3, true, true>;
    return passes == 3 ? conv_tc_kernel<1, 3, 3, false, true> : conv_tc_kernel<1, 1, 3, false, true>;
  }
  if (f16) {
    if (ks == 1) return passes == 3 ? conv_tc_kernel<1, 3, 1, true> : conv_tc_kernel<1, 1, 1, true>;
    return passes == 3 ? conv_tc_kernel<1, 3, 3, true> : conv_tc_kernel<1, 1, 3, true>;
  }
  if (ks == 1) return passes == 3 ? conv_tc_kernel<1, 3, 1, false> : conv_tc_kernel<1, 1, 1, false>;
  if (passes == 3) return conv_tc_kernel<1, 3, 3, false>;
  return mt == 2 ? conv_tc_kernel<2, 1, 3, false> : conv_tc_kernel<1, 1, 3, false>;
}

}  // namespace tc
}  // namespace hcf

extern "C" int hcf_conv_tc_plan_refresh(hcf_conv_tc_plan* pl, void* stream);
extern "C" void hcf_conv_tc_plan_destroy(hcf_conv_tc_plan* p);

struct hcf_conv_tc_plan {
  hcf::tc::Maps maps;
  hcf::tc::Params p;
  hcf::tc::KernelFn fn;
  hcf::tc::LayerDesc* d_layers;
  void* d_ws;                           // weight-stationary schedule: per-layer chunk tables (null: per-item schedule)
  float* d_epi;
  std::vector<const float*>* epi_src;   // per layer: bias, scale device pointers and N (for hcf_conv_tc_plan_refresh)
  std::vector<int>* epi_n;
  long long* d_prof;
  mutable long runs;
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

// fp16 operand variant: the same, with 16-byte alignment counted in fp16 elements (the hi / lo planes
// mirror the fp32 buffer's geometry: same ld, same channel offset)
extern "C" int hcf_conv_tc16_supported(const hcf_conv_args* a) {
  if (!hcf_conv_tc_supported(a)) return 0;
  for (int i = 0; i < a->nseg; ++i)
    if (a->seg[i].ld % 8 != 0) return 0;
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

// fp16 weight image: kin padded per segment to multiples of 64.  split_kin = number of LEADING input channels
// (multiple of 64) packed as [hi ; lo] row blocks for the split; 0 = one pass everywhere, kin = split everywhere
extern "C" int64_t hcf_conv_tc16_weight_bytes(int32_t kin, int32_t cout, int32_t ks, int32_t split_kin) {
  if (kin % 64 != 0 || cout < 1 || cout > 128 || (ks != 1 && ks != 3) || split_kin < 0 || split_kin > kin ||
      split_kin % 64 != 0)
    return 0;
  return (int64_t)((kin + split_kin) / 64) * ks * ks * hcf::tc::n_for(cout) * 128;
}

// w: [cout][kin][ks][ks] fp32 (host) -> per 64-channel chunk [dy][dx][rows][64 fp16]; rows = [hi N ; lo' N] for the
// chunks below split_kin (hi = fp16(w), lo' = fp16((w - hi) * 2048)), hi N otherwise; 128-byte rows pre-swizzled
// like the activations TMA writes
extern "C" int hcf_conv_tc16_pack_weights(const float* w, int32_t kin, int32_t cout, int32_t ks, int32_t split_kin,
                                          void* image) {
  using namespace hcf;
  HCF_REQUIRE(w && image && hcf_conv_tc16_weight_bytes(kin, cout, ks, split_kin) > 0, "tc16_pack: bad args");
  const int N = tc::n_for(cout), KC = kin / 64, SKC = split_kin / 64;
  memset(image, 0, (size_t)hcf_conv_tc16_weight_bytes(kin, cout, ks, split_kin));
  __half* img = reinterpret_cast<__half*>(image);
  size_t base = 0;   // in fp16 elements
  for (int kc = 0; kc < KC; ++kc) {
    const int parts = kc < SKC ? 2 : 1, NB = N * parts;
    for (int dy = 0; dy < ks; ++dy)
      for (int dx = 0; dx < ks; ++dx)
        for (int part = 0; part < parts; ++part)
          for (int n = 0; n < cout; ++n)
            for (int j = 0; j < 64; ++j) {
              const float v = w[(((size_t)n * kin + kc * 64 + j) * ks + dy) * ks + dx];
              const __half hi = __float2half_rn(v);
              const __half val = part == 0 ? hi : __float2half_rn((v - __half2float(hi)) * 2048.0f);
              const int row = part * N + n;
              const int chunk = (j / 8) ^ (row & 7);   // 16-byte chunk = 8 fp16
              img[base + ((size_t)(dy * ks + dx) * NB + row) * 64 + chunk * 8 + (j & 7)] = val;
            }
    base += (size_t)ks * ks * NB * 64;
  }
  return 0;
}

namespace hcf {
namespace tc {

// maps a view pointer of an fp32 buffer to the same element of its fp16 hi / lo plane
static bool shadow_of(const hcf_shadow16* sh, int n_sh, const float* ptr, __half** hi, __half** lo) {
  for (int i = 0; i < n_sh; ++i) {
    const char* base = reinterpret_cast<const char*>(sh[i].f32);
    const char* q = reinterpret_cast<const char*>(ptr);
    if (q >= base && q < base + sh[i].bytes) {
      const size_t el = (size_t)(q - base) / 4;
      *hi = reinterpret_cast<__half*>(sh[i].hi) + el;
      *lo = reinterpret_cast<__half*>(sh[i].lo) + el;
      return true;
    }
  }
  return false;
}

static size_t ws_smem_bytes() {
  return 1024 + (size_t)WS_SA * a_part(1, 3) + (size_t)WS_SB * WS_SLOT + TAIL_BYTES + 3 * WS_TAB_BYTES;
}

// Chunk tables and image grouping of the weight-stationary schedule; false = the chain keeps the per-item schedule
// (accumulators wider than a TMEM slot, raw partial-sum outputs, a single image, or not asked for).
// OPT-IN (HCF_TC_WS=1): parity-green, but measured slower than the per-item schedule on the RRDB chains of configs[1]
// (80x80: 6.0 vs 5.2 ms, 40x40: 2.6 vs 2.0 ms; profiles/r02_ws_*): weights are fetched once per three tiles and the MMA
// issue time per tile drops 19 %, but a pass of three tiles that complete together needs two epilogue rounds before it
// can be published, and that latency chain (not the tensor pipe) then sets the pace -- DESIGN.md section 8.
static bool ws_plan(const std::vector<LayerDesc>& layers, int n, const Params& p, int sms, std::vector<WsLayer>* out,
                    int* n_phases, int* ipp, int* grid) {
  const char* want = getenv("HCF_TC_WS");
  if (!want || atoi(want) == 0) return false;
  if (p.B < 2 || p.nb_max > WS_BIG_COLS) return false;
  const int per_img = p.tiles_x * p.tiles_y;
  int k = WS_GMAX * sms / per_img;          // images per group such that a CTA owns at most WS_GMAX tiles of it
  if (k > p.B / 2) k = p.B / 2;             // at least two groups: a group's next layer waits on nothing recent
  if (k < 1) return false;
  const int np = (p.B + k - 1) / k;
  k = (p.B + np - 1) / np;
  if (const char* env = getenv("HCF_WS_IPP")) {   // tuning: images per group
    const int v = atoi(env);
    if (v >= 1 && v <= p.B / 2) k = v;
  }
  *n_phases = (p.B + k - 1) / k;
  *ipp = k;
  *grid = k * per_img < sms ? k * per_img : sms;
  if ((k * per_img + *grid - 1) / *grid > WS_GMAX) return false;
  out->resize(n);
  for (int i = 0; i < n; ++i) {
    const LayerDesc& L = layers[i];
    WsLayer& W = (*out)[i];
    memset(&W, 0, sizeof(W));
    if (L.raw2 || L.step_z) return false;
    const int split = L.parts == 2 ? L.split_kc : 0;
    if (L.parts == 2 && split < 1) return false;
    const int kcs = L.kchunks, total = kcs + (split < kcs ? split : kcs);
    if (total > WS_MAX_CHUNKS || L.N * L.parts > WS_BIG_COLS) return false;
    W.n_chunks = total; W.taps = L.taps; W.tap0 = L.tap0; W.nb = L.N * L.parts;
    W.wimg = (unsigned long long)reinterpret_cast<uintptr_t>(L.wimg);
    const uint32_t tap_n = (uint32_t)L.N * ROW_BYTES;
    int c = 0;
    auto add = [&](int kc, int kind) {   // kind 0: one pass, 1: hi plane x [B_hi ; B_lo], 2: lo plane x B_hi
      const int seg = kc < L.seg_end[0] ? 0 : (kc < L.seg_end[1] ? 1 : 2);
      const int kl = kc - (seg == 0 ? 0 : L.seg_end[seg - 1]);
      WsChunk& ch = W.ch[c++];
      ch.map = kind == 2 ? L.map_lo[seg] : L.map_idx[seg];
      ch.cch = kl * KCH16 + L.seg_coff[seg];
      const int kmax = (kc == kcs - 1) ? L.seg_last_k[2]
                                      : ((kc == L.seg_end[0] - 1) ? L.seg_last_k[0] : ((kc == L.seg_end[1] - 1) ? L.seg_last_k[1] : 4));
      ch.rows = kind == 1 ? 2 * L.N : L.N;
      ch.dcol = kind == 2 ? L.N : 0;
      ch.idesc = (1u << 4) | (((uint32_t)ch.rows >> 3) << 17) | ((128u >> 4) << 24);   // fp16 operands, fp32 accumulate
      const int sk = kc < split ? kc : split;
      ch.b_off = (uint32_t)(((size_t)sk * 2u + (size_t)(kc - sk)) * (uint32_t)L.taps * tap_n);
      ch.b_tap_src = kc < split ? 2u * tap_n : tap_n;
      const int per = WS_SLOT / (ch.rows * ROW_BYTES);
      int slabs = 1, tps = L.taps;
      if (per < L.taps) { slabs = (L.taps + per - 1) / per; tps = (L.taps + slabs - 1) / slabs; }
      ch.kst = kmax | (tps << 8) | (slabs << 16);
    };
    for (int kc = 0; kc < split && kc < kcs; ++kc) add(kc, 1);
    for (int kc = 0; kc < split && kc < kcs; ++kc) add(kc, 2);
    for (int kc = split; kc < kcs; ++kc) add(kc, 0);
  }
  return true;
}

static int chain_create(const hcf_conv_args* args, const void* const* wtc, const int32_t* layer_passes,
                        const int32_t* layer_split, const int32_t* out_flags, int32_t n, int32_t* done_flags, bool f16, const hcf_shadow16* shadows,
                        int32_t n_shadows, const hcf_seg16* seg16, hcf_conv_tc_plan** out) {
  HCF_REQUIRE(out != nullptr, "tc_chain: null out");
  *out = nullptr;
  HCF_REQUIRE(args && wtc && layer_passes && n >= 1, "tc_chain: bad args");
  HCF_REQUIRE(n == 1 || done_flags != nullptr, "tc_chain: a chain needs the done-flag array");
  HCF_REQUIRE(!f16 || (shadows && n_shadows >= 1 && out_flags), "tc_chain: fp16 chain needs the hi/lo planes and output flags");
  int passes = 1;   // kernel variant: 3 as soon as one layer uses the split
  for (int i = 0; i < n; ++i) {
    HCF_REQUIRE(layer_passes[i] == 1 || layer_passes[i] == 3, "tc_chain: conv %d: passes %d", i, layer_passes[i]);
    if (layer_passes[i] == 3) passes = 3;
  }
  int ks = 1;   // kernel variant: 3 as soon as one conv is 3x3; 1x1 convs of such a chain use the centre tap only
  for (int i = 0; i < n; ++i) ks = args[i].ks > ks ? args[i].ks : ks;
  const int kch = f16 ? KCH16 : KCH;
  for (int i = 0; i < n; ++i) {
    int rc = validate_conv_args(&args[i]);
    if (rc) return rc;
    if (!(f16 ? (seg16 ? hcf_conv_tc_supported(&args[i]) : hcf_conv_tc16_supported(&args[i]))
              : hcf_conv_tc_supported(&args[i]))) {
      set_error("tc_chain: conv %d: unsupported shape", i);
      return HCF_ENOTSUP;
    }
    HCF_REQUIRE(wtc[i] && aligned16(wtc[i]), "tc_chain: conv %d: weight image alignment", i);
    HCF_REQUIRE((args[i].ks == ks || args[i].ks == 1) && args[i].B == args[0].B && args[i].H == args[0].H &&
                    args[i].W == args[0].W, "tc_chain: conv %d: all convs of a chain share [B,H,W]; ks is 1 or the chain's", i);
    HCF_REQUIRE(!args[i].pre || (!args[i].res1 && !args[i].res2 && args[i].pre_ld >= args[i].cout),
                "tc_chain: conv %d: pre excludes res1 / res2", i);
    {   // the epilogue addresses outputs and residuals with 32-bit element offsets
      const uint64_t npix = (uint64_t)args[i].B * args[i].H * args[i].W;
      int ldm = args[i].out_ld;
      if (args[i].out2 && args[i].out2_ld > ldm) ldm = args[i].out2_ld;
      if (args[i].res1 && args[i].res1_ld > ldm) ldm = args[i].res1_ld;
      if (args[i].res2 && args[i].res2_ld > ldm) ldm = args[i].res2_ld;
      if (args[i].pre && args[i].pre_ld > ldm) ldm = args[i].pre_ld;
      if (args[i].raw2 && args[i].raw2_ld > ldm) ldm = args[i].raw2_ld;
      HCF_REQUIRE(npix * (uint64_t)ldm < (1ull << 32), "tc_chain: conv %d: buffer too large for 32-bit element offsets", i);
    }
  }
  EncodeTiledFn enc = get_encode();
  HCF_REQUIRE(enc != nullptr, "tc_chain: cuTensorMapEncodeTiled entry point not found");
  const hcf_conv_args* a0 = &args[0];
  hcf_conv_tc_plan* pl = new hcf_conv_tc_plan();
  memset(pl, 0, sizeof(*pl));
  Params& p = pl->p;
  p.B = a0->B; p.H = a0->H; p.W = a0->W;
  const int sms = num_sms();
  int nmax = 0, nbmax = 0;
  for (int i = 0; i < n; ++i) {
    const int nn = n_for(args[i].cout), nb = nn * (layer_passes[i] == 3 ? 2 : 1);
    nmax = nmax > nn ? nmax : nn;
    nbmax = nbmax > nb ? nbmax : nb;
  }
  p.nb_max = nbmax;
  // sub-tiles per work item: 2 halves the weight traffic per pixel but quantises worse on small images
  int mt = 1;
  const char* env = getenv("HCF_TC_MT");
  if (!f16 && passes == 1 && ks == 3 && n == 1) {
    const long items1 = (long)a0->B * ceil_div(a0->H, 16) * ceil_div(a0->W, 8);
    const long items2 = (long)a0->B * ceil_div(a0->H, 32) * ceil_div(a0->W, 8);
    const double t1 = (double)ceil_div((int)items1, sms) * (a_part(1, 3) + 9.0 * nmax * 128);
    const double t2 = (double)ceil_div((int)items2, sms) * (a_part(2, 3) + 9.0 * nmax * 128);
    mt = (t2 < 0.80 * t1) ? 2 : 1;
    if (env && (env[0] == '1' || env[0] == '2')) mt = env[0] - '0';
  }
  {
    const char* dbg = getenv("HCF_TC_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  int slot_taps = 0;
  int tail_extra = 0;
  for (int i = 0; i < n; ++i)
    if (args[i].step) tail_extra = 2 * STEP_TAB_BYTES;
  p.step_tab = tail_extra ? 1 : 0;
  if (tail_extra && mt != 1) mt = 1;
  if (!pick_rings(mt, passes, ks, p.nb_max, tail_extra, &p.sa, &p.sb, &slot_taps, &pl->smem_bytes)) {
    delete pl;
    set_error("tc_chain: tile does not fit in shared memory");
    return HCF_ENOTSUP;
  }
  if (const char* rings = getenv(n > 1 ? "HCF_TC_RINGS" : "HCF_TC_RINGS_SINGLE")) {   // tuning: "sa,sb,slab_taps"
    int ra = 0, rb = 0, rs = 0;
    if (sscanf(rings, "%d,%d,%d", &ra, &rb, &rs) == 3 && ra >= 1 && ra <= 4 && rb >= 2 && rb <= 18 &&
        (rs == 1 || rs == ks || rs == ks * ks)) {
      const size_t need = 1024 + (size_t)ra * a_part(mt, ks) * (passes == 3 ? 2 : 1) +
                          (size_t)rb * rs * p.nb_max * ROW_BYTES + TAIL_BYTES + tail_extra;
      if (need <= (size_t)SMEM_LIMIT) {
        p.sa = ra; p.sb = rb; slot_taps = rs;
        pl->smem_bytes = need;
      }
    }
  }
  p.slot_bytes = slot_taps * p.nb_max * ROW_BYTES;
  p.tiles_x = ceil_div(a0->W, TW); p.tiles_y = ceil_div(a0->H, TH * mt);
  p.n_tiles = p.tiles_x * p.tiles_y * a0->B;
  p.n_layers = n;
  p.n_items = p.n_tiles * n;
  p.done = n > 1 ? done_flags : nullptr;
  if (n > 1 && getenv("HCF_TC_PROF")) {
    if (cudaMalloc(&pl->d_prof, sizeof(long long) * PROF_N) == cudaSuccess) cudaMemset(pl->d_prof, 0, sizeof(long long) * PROF_N);
    p.prof = pl->d_prof;
  }

  // ---- tensor maps, de-duplicated.  A segment whose channel count is a multiple of the chunk width never
  // reads beyond its last chunk, so such segments of one buffer share a map of the widest extent seen.
  // fp16 chains treat every multiple of 32 that way: the half-empty last chunk then reads finite stale values
  // of the same buffer against zero weight rows (buffers are zero-initialised and only ever hold conv outputs).
  // Channel-slice views of ONE buffer (same row pitch, pointers less than one pixel row apart) share a map whose
  // base is the lowest of them; a segment addresses it with its channel offset.  Ragged views keep their own map.
  struct MapKey { const char* ptr; int ld; int C; bool ragged; };
  std::vector<MapKey> keys;
  const int esz_i = f16 ? 2 : 4;
  auto is_ragged = [&](int C) { return (C % (f16 ? 32 : kch)) != 0; };
  auto map_note = [&](const void* vptr, int ld, int C) {   // pass 1: cluster bases
    const char* ptr = reinterpret_cast<const char*>(vptr);
    if (is_ragged(C)) return;
    for (size_t k = 0; k < keys.size(); ++k) {
      if (keys[k].ragged || keys[k].ld != ld) continue;
      const long long d = ptr - keys[k].ptr;
      if (d > -(long long)ld * esz_i && d < (long long)ld * esz_i) {
        if (d < 0) { keys[k].C += (int)(-d / esz_i); keys[k].ptr = ptr; }
        const int need = (int)((ptr - keys[k].ptr) / esz_i) + C;
        if (need > keys[k].C) keys[k].C = need;
        return;
      }
    }
    keys.push_back({ptr, ld, C, false});
  };
  auto map_for = [&](const void* vptr, int ld, int C, int* coff) {   // pass 2: key index + channel offset
    const char* ptr = reinterpret_cast<const char*>(vptr);
    *coff = 0;
    if (is_ragged(C)) {
      for (size_t k = 0; k < keys.size(); ++k)
        if (keys[k].ragged && keys[k].ptr == ptr && keys[k].ld == ld && keys[k].C == C) return (int)k;
      keys.push_back({ptr, ld, C, true});
      return (int)keys.size() - 1;
    }
    for (size_t k = 0; k < keys.size(); ++k) {
      if (keys[k].ragged || keys[k].ld != ld) continue;
      const long long d = ptr - keys[k].ptr;
      if (d >= 0 && d < (long long)ld * esz_i) {
        *coff = (int)(d / esz_i);
        return (int)k;
      }
    }
    return -1;   // unreachable after pass 1
  };
  auto seg_planes = [&](int i, int s_, __half** hi, __half** lo, int* ld16) -> int {   // fp16: the planes of a segment
    const hcf_seg& sg = args[i].seg[s_];
    *ld16 = sg.ld;
    const hcf_seg16* ov = seg16 ? &seg16[i * 3 + s_] : nullptr;
    if (ov && ov->hi) {
      *hi = reinterpret_cast<__half*>(const_cast<void*>(ov->hi));
      *lo = reinterpret_cast<__half*>(const_cast<void*>(ov->lo));
      *ld16 = ov->ld;
      return 0;
    }
    return shadow_of(shadows, n_shadows, sg.ptr, hi, lo) ? 0 : -1;
  };
  for (int i = 0; i < n; ++i)   // pass 1
    for (int s_ = 0; s_ < args[i].nseg; ++s_) {
      const hcf_seg& sg = args[i].seg[s_];
      if (f16) {
        __half *hi = nullptr, *lo = nullptr;
        int ld16 = 0;
        if (seg_planes(i, s_, &hi, &lo, &ld16) == 0) {
          map_note(hi, ld16, sg.C);
          if (layer_passes[i] == 3 && lo) map_note(lo, ld16, sg.C);
        }
      } else {
        map_note(sg.ptr, sg.ld, sg.C);
      }
    }
  std::vector<LayerDesc> layers(n);
  for (int i = 0; i < n; ++i) {
    const hcf_conv_args& a = args[i];
    LayerDesc& L = layers[i];
    memset(&L, 0, sizeof(L));
    L.nseg = a.nseg;
    L.parts = layer_passes[i] == 3 ? 2 : 1;
    int kc = 0;
    for (int s = 0; s < 3; ++s) {
      if (s < a.nseg) {
        const hcf_seg& sg = a.seg[s];
        if (f16) {
          __half *hi = nullptr, *lo = nullptr;
          int ld16 = sg.ld;
          if (seg_planes(i, s, &hi, &lo, &ld16) != 0) {
            delete pl;
            set_error("tc_chain: conv %d segment %d: no fp16 planes registered for this buffer", i, s);
            return HCF_EINVAL;
          }
          if (!aligned16(hi) || ld16 % 8 != 0 || (L.parts == 2 && (!lo || !aligned16(lo)))) {
            delete pl;
            set_error("tc_chain: conv %d segment %d: fp16 view breaks TMA's 16-byte rules", i, s);
            return HCF_ENOTSUP;
          }
          int coff_lo = 0;
          L.map_idx[s] = map_for(hi, ld16, sg.C, &L.seg_coff[s]);
          L.map_lo[s] = L.parts == 2 ? map_for(lo, ld16, sg.C, &coff_lo) : 0;   // (same geometry: same offset)
        } else {
          L.map_idx[s] = map_for(sg.ptr, sg.ld, sg.C, &L.seg_coff[s]);
        }
        kc += (sg.C + kch - 1) / kch;
        const int rem = sg.C % kch == 0 ? kch : sg.C % kch;     // channels in the segment's last chunk
        L.seg_last_k[s] = (rem + kch / 4 - 1) / (kch / 4);
      } else {
        L.seg_last_k[s] = 4;
      }
      L.seg_end[s] = s < a.nseg ? kc : (1 << 30);
    }
    L.seg_end[a.nseg - 1] = 1 << 30;
    L.seg_last_k[2] = L.seg_last_k[a.nseg - 1];   // slot 2 = the FINAL segment's value (matched by kc == kchunks - 1)
    L.kchunks = kc;
    L.split_kc = L.parts == 2 ? kc : 0;
    if (f16 && L.parts == 2 && layer_split && layer_split[i] >= 0) {
      if (layer_split[i] % kch != 0 || layer_split[i] > kc * kch) {
        delete pl;
        set_error("tc_chain: conv %d: split range %d is not a multiple of %d channels", i, layer_split[i], kch);
        return HCF_EINVAL;
      }
      L.split_kc = layer_split[i] / kch;
    }
    L.N = n_for(a.cout);
    if (a.ks == ks) {   // largest tap count (whole chunk, one dy row, one tap) of this layer that fits a ring slot
      const int tap = L.N * L.parts * ROW_BYTES;
      L.slab_taps = (ks * ks * tap <= p.slot_bytes) ? ks * ks : ((ks * tap <= p.slot_bytes) ? ks : 1);
      L.taps = ks * ks;
      L.tap0 = 0;
    } else {            // 1x1 conv inside a 3x3 chain: same halo tile, centre tap
      L.slab_taps = 1;
      L.taps = 1;
      L.tap0 = (ks * ks) / 2;
    }
    L.cout = a.cout;
    L.act = a.act;
    L.wimg = reinterpret_cast<const float*>(wtc[i]); L.bias = a.bias; L.scale = a.scale;
    L.out = a.out; L.out_ld = a.out_ld; L.out2 = a.out2; L.out2_ld = a.out2_ld;
    L.res1 = a.res1; L.res1_ld = a.res1_ld; L.alpha1 = a.alpha1;
    L.res2 = a.res2; L.res2_ld = a.res2_ld; L.alpha2 = a.alpha2;
    L.pre = a.pre; L.pre_ld = a.pre_ld;
    L.raw2 = a.raw2; L.raw2_ld = a.raw2_ld;
    if (a.raw2 && !(a.cout == 64 && !a.res1 && !a.res2 && !a.out2 && !a.step && a.raw2_ld % 4 == 0 && aligned16(a.raw2) &&
                    a.raw2_ld >= 32)) {
      delete pl;
      set_error("tc_chain: conv %d: raw2 needs cout == 64, no residuals / second output, a 16-byte aligned view", i);
      return HCF_EINVAL;
    }
    if (a.step) {
      const hcf_conv_step& st = *a.step;
      if (!(st.z && st.C >= 2 && st.C <= STEP_MAXC && st.n_pass >= 1 && st.n_pass < st.C && st.z_ld >= st.C &&
            a.cout == 2 * (st.C - st.n_pass) && L.N <= 32 && st.an_scale && st.an_bias && !a.res1 && !a.res2 && !a.pre &&
            !a.out2 && (!st.z16_hi || st.z16_ld >= st.n_pass))) {
        delete pl;
        set_error("tc_chain: conv %d: fused FlowStep needs cout == 2 * (C - n_pass) <= 32, C <= %d, no residuals", i, STEP_MAXC);
        return HCF_EINVAL;
      }
      L.step_z = st.z; L.step_z_ld = st.z_ld; L.step_C = st.C; L.step_npass = st.n_pass;
      L.step_w = st.w; L.step_sc = st.an_scale; L.step_b = st.an_bias;
      L.step_z16 = reinterpret_cast<__half*>(st.z16_hi); L.step_z16_ld = st.z16_ld;
      L.step_z16_lo = reinterpret_cast<__half*>(st.z16_lo);
      L.out = nullptr; L.out2 = nullptr;
    }
    bool ov = aligned16(a.out) && a.out_ld % 4 == 0;
    if (a.out2) ov = ov && aligned16(a.out2) && a.out2_ld % 4 == 0;
    if (a.res1) ov = ov && aligned16(a.res1) && a.res1_ld % 4 == 0;
    if (a.res2) ov = ov && aligned16(a.res2) && a.res2_ld % 4 == 0;
    if (a.pre) ov = ov && aligned16(a.pre) && a.pre_ld % 4 == 0;
    L.out_vec = ov ? 1 : 0;
    if (f16) {
      const int fl = out_flags[i];
      __half *hi = nullptr, *lo = nullptr;
      if ((fl & (HCF_OUT_HI | HCF_OUT_LO)) && !a.step) {
        if (!shadow_of(shadows, n_shadows, a.out, &hi, &lo)) {
          delete pl;
          set_error("tc_chain: conv %d: no fp16 planes registered for the output buffer", i);
          return HCF_EINVAL;
        }
        L.out_hi = hi;
        L.out_lo = (fl & HCF_OUT_LO) ? lo : nullptr;
        if (a.out2) {
          if (!shadow_of(shadows, n_shadows, a.out2, &hi, &lo)) {
            delete pl;
            set_error("tc_chain: conv %d: no fp16 planes registered for the second output buffer", i);
            return HCF_EINVAL;
          }
          L.out2_hi = hi;
          L.out2_lo = (fl & HCF_OUT_LO) ? lo : nullptr;
        }
      }
      if (!(fl & HCF_OUT_F32) || a.step) { L.out = nullptr; L.out2 = nullptr; }
    }
  }
  if ((int)keys.size() > MAX_MAPS) {
    delete pl;
    set_error("tc_chain: %d distinct input views (max %d)", (int)keys.size(), MAX_MAPS);
    return HCF_ENOTSUP;
  }
  const cuuint32_t box[4] = {(cuuint32_t)kch, (cuuint32_t)halo_w(ks), (cuuint32_t)halo_rows(mt, ks), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const cuuint64_t esz = f16 ? 2 : 4;
  for (int i = 0; i < MAX_MAPS; ++i) {
    const MapKey& k = keys[i < (int)keys.size() ? i : 0];   // unused maps alias map 0 (never dereferenced)
    const cuuint64_t dims[4] = {(cuuint64_t)k.C, (cuuint64_t)a0->W, (cuuint64_t)a0->H, (cuuint64_t)a0->B};
    const cuuint64_t ld_b = (cuuint64_t)k.ld * esz;
    const cuuint64_t strides[3] = {ld_b, ld_b * a0->W, ld_b * a0->W * a0->H};
    CUresult r = enc(&pl->maps.m[i], f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                     const_cast<char*>(k.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete pl;
      set_error("tc_chain: cuTensorMapEncodeTiled failed with %d (view %d, C %d, ld %d)", (int)r, i, k.C, k.ld);
      return HCF_EINVAL;
    }
  }
  // inline epilogue constants: [layer][bias 128 | scale 128], gathered on the device from the padded per-layer vectors
  pl->epi_src = new std::vector<const float*>();
  pl->epi_n = new std::vector<int>();
  for (int i = 0; i < n; ++i) {
    pl->epi_src->push_back(args[i].bias);
    pl->epi_src->push_back(args[i].scale);
    pl->epi_n->push_back(layers[i].N);   // bias / scale are padded to >= N entries
  }
  cudaError_t e = cudaMalloc(&pl->d_epi, sizeof(float) * 256 * n);
  if (e == cudaSuccess) {
    std::vector<float> ident((size_t)256 * n);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < 128; ++j) { ident[256 * i + j] = 0.f; ident[256 * i + 128 + j] = 1.f; }
    e = cudaMemcpy(pl->d_epi, ident.data(), sizeof(float) * 256 * n, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && hcf_conv_tc_plan_refresh(pl, nullptr) != 0) e = cudaErrorUnknown;
  p.epi = pl->d_epi;
  if (e == cudaSuccess) e = cudaMalloc(&pl->d_layers, sizeof(LayerDesc) * n);
  if (e == cudaSuccess)
    e = cudaMemcpy(pl->d_layers, layers.data(), sizeof(LayerDesc) * n, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("tc_chain: layer table upload: %s", cudaGetErrorString(e));
    hcf_conv_tc_plan_destroy(pl);
    return (int)e;
  }
  p.layers = pl->d_layers;
  pl->grid = dim3((unsigned)(p.n_tiles < sms ? p.n_tiles : sms));
  p.tpc = 0;
  // ---- weight-stationary schedule (conv_ws_kernel.cuh) for fp16 3x3 chains whose accumulators fit its TMEM slots
  std::vector<WsLayer> ws;
  int ws_grid = 0;
  const bool use_ws = f16 && n > 1 && ks == 3 && mt == 1 && !tail_extra &&
                      ws_plan(layers, n, p, sms, &ws, &p.n_phases, &p.ipp, &ws_grid);
  if (use_ws) {
    e = cudaMalloc(&pl->d_ws, sizeof(WsLayer) * n);
    if (e == cudaSuccess) e = cudaMemcpy(pl->d_ws, ws.data(), sizeof(WsLayer) * n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      set_error("tc_chain: chunk table upload: %s", cudaGetErrorString(e));
      hcf_conv_tc_plan_destroy(pl);
      return (int)e;
    }
    p.ws_layers = pl->d_ws;
    pl->grid = dim3((unsigned)ws_grid);
    pl->smem_bytes = ws_smem_bytes();
  }
  if (!use_ws && n > 1 && p.n_tiles > sms && p.n_tiles <= 3 * sms) {
    // few tiles per CTA and layer (40x40 level: 240 tiles on 148 SMs): instead of rotating the items over the CTAs,
    // each CTA can own tpc tiles for all layers when the tile count divides into a grid that still covers >= 3/4 of
    // the SMs.  Opt-in (HCF_TC_STATIC=1): measured on the same box, 120 owning CTAs are not faster than 148 rotating
    // ones (encoder chain 2.29 vs 2.24 ms, FlowStep chains equal) -- profiles/r01b_static_ownership_ab.log
    const char* env_s = getenv("HCF_TC_STATIC");
    const bool want = env_s ? atoi(env_s) != 0 : false;
    for (int t = 2; t <= 3 && want && p.tpc == 0; ++t)
      if (p.n_tiles % t == 0 && p.n_tiles / t <= sms && 4 * (p.n_tiles / t) >= 3 * sms) {
        p.tpc = t;
        pl->grid = dim3((unsigned)(p.n_tiles / t));
      }
  }
  pl->threads = f16 ? 384 : (passes == 3 ? 320 : 192);
  if (tail_extra && ks != 3) {
    hcf_conv_tc_plan_destroy(pl);
    set_error("tc_chain: fused FlowStep layers need a 3x3 chain");
    return HCF_ENOTSUP;
  }
  bool any_direct = false;   // (mirrors the kernel's per-layer test: hi plane only, no residual / addend, one pass)
  for (int i = 0; i < n && f16; ++i) {
    const LayerDesc& L = layers[i];
    if (L.out_hi && !L.out_lo && !L.out && !L.out2 && !L.out2_hi && !L.out2_lo && !L.res1 && !L.res2 && !L.pre && !L.raw2 &&
        !L.step_z && L.parts == 1 && L.out_vec && L.cout % 32 == 0 && L.out_ld % 8 == 0)
      any_direct = true;
  }
  pl->fn = use_ws ? conv_ws_kernel : pick_kernel(mt, passes, ks, f16, tail_extra != 0, any_direct);
  e = cudaFuncSetAttribute(reinterpret_cast<const void*>(pl->fn), cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SMEM_LIMIT);
  if (e != cudaSuccess) {
    set_error("tc_chain: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    hcf_conv_tc_plan_destroy(pl);
    return (int)e;
  }
  if (n > 1) {   // a chain needs its whole grid resident at once (the launch is cooperative): check it now
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinterpret_cast<const void*>(pl->fn), pl->threads,
                                                      pl->smem_bytes);
    if (e != cudaSuccess || (long)per_sm * sms < (long)pl->grid.x) {
      set_error("tc_chain: %d CTAs of %d threads / %zu B shared memory cannot be co-resident (%d per SM x %d SMs)",
                (int)pl->grid.x, pl->threads, pl->smem_bytes, per_sm, sms);
      hcf_conv_tc_plan_destroy(pl);
      return HCF_ENOTSUP;
    }
  }
  *out = pl;
  return 0;
}

}  // namespace tc
}  // namespace hcf

// A chain of n convolutions on the same [B,H,W] grid with the same kernel size, each TC-eligible,
// executed by one persistent launch.  Conv i may read anything convs < i wrote (dependencies are
// tracked per 3x3 tile neighbourhood, which also covers write-after-read).  `done_flags`
// (device, B*ceil(H/16)*ceil(W/8) int32) must be zeroed before every run; NULL is allowed for n == 1.
extern "C" int hcf_conv_chain_create(const hcf_conv_args* args, const float* const* wtc, const int32_t* layer_passes,
                                     int32_t n, int32_t* done_flags, hcf_conv_tc_plan** out) {
  return hcf::tc::chain_create(args, reinterpret_cast<const void* const*>(wtc), layer_passes, nullptr, nullptr, n,
                               done_flags, false, nullptr, 0, nullptr, out);
}

// The same chain on fp16 operands (see include/hcflow_b200.h).
extern "C" int hcf_conv_chain16_create(const hcf_conv_args* args, const void* const* w16, const int32_t* layer_passes,
                                       const int32_t* layer_split, const int32_t* out_flags, int32_t n, int32_t* done_flags,
                                       const hcf_shadow16* shadows, int32_t n_shadows, const hcf_seg16* seg16,
                                       hcf_conv_tc_plan** out) {
  return hcf::tc::chain_create(args, w16, layer_passes, layer_split, out_flags, n, done_flags, true, shadows, n_shadows,
                               seg16, out);
}

// Re-gathers the per-layer bias / scale vectors into the plan's inline table (they are read through it, not through
// the argument pointers): call after the arrays behind hcf_conv_args.bias / .scale were rewritten in place.
extern "C" int hcf_conv_tc_plan_refresh(hcf_conv_tc_plan* pl, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(pl && pl->d_epi && pl->epi_src && pl->epi_n, "tc_plan_refresh: bad plan");
  cudaStream_t st = (cudaStream_t)stream;
  for (size_t i = 0; i < pl->epi_n->size(); ++i) {
    const float* b = (*pl->epi_src)[2 * i];
    const float* sc = (*pl->epi_src)[2 * i + 1];
    const size_t bytes = sizeof(float) * (size_t)(*pl->epi_n)[i];
    cudaError_t e = cudaSuccess;
    if (b) e = cudaMemcpyAsync(pl->d_epi + 256 * i, b, bytes, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && sc) e = cudaMemcpyAsync(pl->d_epi + 256 * i + 128, sc, bytes, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) {
      set_error("tc_plan_refresh: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  return 0;
}

extern "C" int hcf_conv_tc_plan_create(const hcf_conv_args* a, const float* wtc, int32_t passes,
                                       hcf_conv_tc_plan** out) {
  return hcf_conv_chain_create(a, &wtc, &passes, 1, nullptr, out);
}

extern "C" int32_t hcf_conv_tc_plan_layers(const hcf_conv_tc_plan* pl) { return pl ? pl->p.n_layers : 0; }

// Sticky device status word of a plan (NULL = none): bit 0 (HCF_STATUS_F16_OVERFLOW) an fp16 operand plane saturated,
// bit 1 (HCF_STATUS_DEP_TIMEOUT) a dependency wait of a chained launch timed out.  The host reads / clears it.
extern "C" int hcf_conv_tc_plan_set_status(hcf_conv_tc_plan* pl, int32_t* status) {
  using namespace hcf;
  HCF_REQUIRE(pl != nullptr, "tc_plan_set_status: null plan");
  pl->p.status = status;
  return 0;
}

extern "C" int hcf_conv_tc_run(const hcf_conv_tc_plan* pl, void* stream) {
  using namespace hcf;
  HCF_REQUIRE(pl != nullptr, "tc_run: null plan");
  ++pl->runs;
  if (pl->p.n_layers > 1) {
    // A chained launch's CTAs wait on each other's tiles: they must all be resident at once.  A cooperative launch
    // makes the driver guarantee that (it fails with cudaErrorCooperativeLaunchTooLarge instead of deadlocking when
    // the grid cannot be co-resident, and is not started next to a kernel that holds the SMs it needs).
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = pl->grid;
    cfg.blockDim = dim3((unsigned)pl->threads);
    cfg.dynamicSmemBytes = pl->smem_bytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, pl->fn, pl->maps, pl->p);
    if (e != cudaSuccess) {
      set_error("hcf_conv_tc_run (cooperative, %d CTAs): %s", (int)pl->grid.x, cudaGetErrorString(e));
      return (int)e;
    }
    return finish_launch("hcf_conv_tc_run");
  }
  pl->fn<<<pl->grid, pl->threads, pl->smem_bytes, (cudaStream_t)stream>>>(pl->maps, pl->p);
  return finish_launch("hcf_conv_tc_run");
}

extern "C" void hcf_conv_tc_plan_destroy(hcf_conv_tc_plan* p) {
  if (!p) return;
  if (p->d_prof) {   // HCF_TC_PROF=1: average cycles per CTA and launch, by role and wait class
    long long h[hcf::tc::PROF_N];
    cudaDeviceSynchronize();
    if (cudaMemcpy(h, p->d_prof, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
      static const char* names[hcf::tc::PROF_N] = {"P.total", "P.deps", "P.emptyA", "P.emptyB", "M.total", "M.tmem_empty",
                                                   "M.fullA", "M.fullB", "M.convA", "E.total", "E.tmem_full", "E.body",
                                                   "E.publish", "E.layer", "E.row", "E.coal", "M.issue", "launches"};
      const double items = (double)p->p.n_items / p->grid.x;
      const double launches = h[hcf::tc::PROF_LAUNCHES] > 0 ? (double)h[hcf::tc::PROF_LAUNCHES] : 1.0;
      fprintf(stderr, "[hcf prof] chain layers=%d tiles=%d items/CTA=%.1f launches=%.0f; cycles per item:", p->p.n_layers,
              p->p.n_tiles, items, launches);
      for (int i = 0; i < hcf::tc::PROF_LAUNCHES; ++i)
        fprintf(stderr, " %s=%.0f", names[i], (double)h[i] / p->grid.x / launches / items);
      fprintf(stderr, "\n");
    }
    cudaFree(p->d_prof);
  }
  if (p->d_layers) cudaFree(p->d_layers);
  if (p->d_ws) cudaFree(p->d_ws);
  if (p->d_epi) cudaFree(p->d_epi);
  delete p->epi_src;
  delete p->epi_n;
  delete p;
}
