// Micro-benchmark of tcgen05.mma issue/throughput on one SM (every SM runs the same loop):
// cycles per MMA for M=128, N in {32..256}, kind tf32 / bf16, A from shared memory or TMEM,
// accumulating into 1, 2 or 4 alternating TMEM accumulators.  Operand contents are irrelevant
// (shared memory is zero-filled); only the timing matters.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench tools/mma_bench.cu
//   ./tools/mma_bench            -> JSON lines
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <int KIND>   // 0 tf32, 1 bf16
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct Result { long long cycles; };
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred;
}

// ASRC 0: A from smem (aligned tile), 1: A from smem through a halo-shifted descriptor (pitch 1280), 2: A from TMEM
template <int KIND, int ASRC, int N, int NACC>
__global__ void __launch_bounds__(128, 1) mma_loop(int iters, Result* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  // zero the operand area: A 4 stages x 24 KB, B 64 KB
  for (int i = threadIdx.x; i < (160 * 1024) / 16; i += blockDim.x)
    reinterpret_cast<float4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {   // whole warp follows the loop, one elected lane issues (keeps descriptors in uniform registers)
    const uint32_t a_base = base, b_base = base + 96 * 1024;
    const uint32_t afmt = KIND == 0 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (afmt << 7) | (afmt << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a_tmpl = make_desc(0, ASRC == 1 ? 1280u : 1024u);
    const uint64_t b_tmpl = make_desc(0, 1024u);
    // accumulators at columns 0, N, 2N ...; TMEM A operand (ASRC 2) lives in the last 32 columns
    const uint32_t a_tm = tmem + 480u;
    const uint64_t a0 = a_tmpl + (a_base >> 4), b0 = b_tmpl + (b_base >> 4);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {   // rep 0 = warm-up
      t0 = clock64();
      for (int i = 0; i < iters; i += 8) {
       if (elect_one()) {
        // all offsets below are compile-time constants: the issuing thread executes ~1 instruction per MMA
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t d = tmem + (uint32_t)((u % NACC) * N);
          constexpr int dummy = 0; (void)dummy;
          const int tap = (u >> 2), k = u & 3;
          const uint64_t bd = b0 + (uint32_t)(((tap % 2) * N * 128) >> 4) + 2u * k;
          if (ASRC == 2) {
            umma_ts<KIND>(d, a_tm + 8u * k, bd, idesc, 1u);
          } else {
            const uint32_t off = ASRC == 1 ? (uint32_t)((tap * 10 + 1 + tap) * 128) : (uint32_t)(tap * 16384);
            umma_ss<KIND>(d, a0 + (off >> 4) + 2u * k, bd, idesc, 1u);
          }
        }
       }
       __syncwarp();
      }
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      __syncwarp();
      mbar_wait(smem_u32(&bar), rep & 1);
      t1 = clock64();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out->cycles = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int KIND, int ASRC, int N, int NACC>
static void run(Result* d_out) {
  const int nacc = NACC;
  const int iters = 4096;
  cudaFuncSetAttribute(mma_loop<KIND, ASRC, N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  mma_loop<KIND, ASRC, N, NACC><<<148, 128, 200 * 1024>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  Result r{};
  cudaMemcpy(&r, d_out, sizeof(r), cudaMemcpyDeviceToHost);
  const double cyc = (double)r.cycles / iters;
  const double kelems = KIND == 0 ? 8.0 : 16.0;
  printf("{\"kind\": \"%s\", \"a_src\": \"%s\", \"N\": %d, \"nacc\": %d, \"cycles_per_mma\": %.1f, "
         "\"mac_per_clk\": %.0f, \"err\": \"%s\"}\n",
         KIND == 0 ? "tf32" : "bf16", ASRC == 0 ? "smem" : (ASRC == 1 ? "smem_halo" : "tmem"), N, nacc, cyc,
         128.0 * N * kelems / cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
  fflush(stdout);
}

int main() {
  Result* d_out;
  cudaMalloc(&d_out, sizeof(Result));
#define ALL(N, NACC) run<0, 0, N, NACC>(d_out); run<0, 1, N, NACC>(d_out); run<0, 2, N, NACC>(d_out); \
  run<1, 0, N, NACC>(d_out); run<1, 2, N, NACC>(d_out);
  ALL(32, 1) ALL(32, 2) ALL(32, 4) ALL(64, 1) ALL(64, 2) ALL(64, 4) ALL(96, 1) ALL(96, 2) ALL(96, 4)
  ALL(128, 1) ALL(128, 2) ALL(160, 1) ALL(160, 2) ALL(192, 1) ALL(192, 2) ALL(256, 1)
  return 0;
}
