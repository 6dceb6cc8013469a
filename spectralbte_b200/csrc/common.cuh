// spectralbte_b200/csrc/common.cuh
// Shared device helpers for the sm_100a kernels: complex arithmetic on double2, mbarrier + TMA bulk
// copy wrappers (inline PTX), deterministic block reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sbte {

// ---------------------------------------------------------------- complex helpers
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void cmac(double2& acc, double w, double2 p) {
  acc.x = fma(w, p.x, acc.x);
  acc.y = fma(w, p.y, acc.y);
}

// ---------------------------------------------------------------- symmetrised xi_x enumeration
// For f == g the summand W[zeta][xi] f^[xi] f^[sigma(xi)] pairs up under the involution
// sigma_zeta(xi) = wrap(zeta + N/2 - xi): with Ws[zeta][xi] = W[zeta][xi] + W[zeta][sigma(xi)] only one xi of
// each pair has to be visited.  Pairing is decided on the x component: with a = (zeta_x + N/2) mod N the
// representatives are xi_x in [0, a/2] and [a+1, (a+N)/2] (self-paired planes xi_x = sigma_x(xi_x) keep
// their original weights and are visited whole).  nrep = N/2 + 1 (a even) or N/2 (a odd).
__host__ __device__ __forceinline__ int sym_nrep(int N, int zx) {
  const int a = (zx + N / 2) % N;
  return a / 2 + 1 + (a + N) / 2 - a;
}
__host__ __device__ __forceinline__ int sym_rep(int N, int zx, int c) {   // c-th representative xi_x
  const int a = (zx + N / 2) % N;
  const int h = a / 2;
  return (c <= h) ? c : c + (a - h);
}

// Everything below is inline PTX (or warp intrinsics).  tests/emul/cuda_emul.h -- test infrastructure that runs the
// batched kernels' control flow on the host -- defines SBTE_HOST_EMUL and supplies same-named host stand-ins.
#ifndef SBTE_HOST_EMUL
#define SBTE_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#define SBTE_SETMAXNREG_DEC(n) asm volatile("setmaxnreg.dec.sync.aligned.u32 " #n ";")
#define SBTE_SETMAXNREG_INC(n) asm volatile("setmaxnreg.inc.sync.aligned.u32 " #n ";")

// ---------------------------------------------------------------- shared-memory addressing
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier (transaction barrier)
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a lost TMA completion becomes a trap (launch error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}

// ---------------------------------------------------------------- TMA: 1-D bulk copy global -> shared
// dst/src 16-byte aligned, bytes a multiple of 16; completion is signalled on `bar` (complete_tx).
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// TMA: 2-D tiled tensor copy global -> shared through a CUtensorMap (coordinates innermost first)
__device__ __forceinline__ void tma_tensor2d_g2s(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------- programmatic dependent launch
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in
// the stream is still running: everything before pdl_wait() must not touch the predecessor's output.
// pdl_launch_dependents() lets the successor's CTAs be scheduled early.  Both are no-ops otherwise.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- streaming global loads
// 128-bit read-only load that does not allocate in L1 (the weight stream is touched exactly once)
__device__ __forceinline__ double2 ldg_stream_f64x2(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// ---------------------------------------------------------------- fire-and-forget accumulation in L2
// p[0] += v without waiting for the old value (SASS RED.E.ADD.F64): an IEEE double add performed at the L2 slice.  Used
// where exactly ONE thread ever updates the location, in program order, so the result is the same left fold -- bit for
// bit -- as a load / add / store by that thread would give, without its round trip.
__device__ __forceinline__ void red_add_f64(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// ---------------------------------------------------------------- hand-over of a running sum between two CTAs of a launch
// One word per (producer CTA, warp): the producer stores its data, fences, and publishes 1; the consumer waits for the 1
// (bounded in wall time: a lost publication traps instead of hanging the GPU), reads the data past L1 and re-arms the word with 0 for
// the next launch in the stream.  Used by the batched convolution where a stream-K cut falls inside a xi_x chunk.
__device__ __forceinline__ void carry_publish(int* flag) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
}
__device__ __forceinline__ void carry_await(const int* flag) {
  int v = 0;
  unsigned spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v != 0) break;
    __nanosleep(64);
    // The producer publishes at the very start of its range and is dispatched before its consumer, so this wait is
    // short unless the device is shared and the producer's CTA has to queue behind somebody else's kernel: two minutes
    // of wall time (the halo waits' default bound) before a lost publication becomes a trap.
    if ((++spins & 1023u) == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 120000000000ULL) __trap();
    }
  }
}
__device__ __forceinline__ void carry_rearm(int* flag) { *reinterpret_cast<volatile int*>(flag) = 0; }
__device__ __forceinline__ double2 carry_load(const double2* p) { return __ldcg(p); }

// ---------------------------------------------------------------- deterministic block reduction
// Sums `NV` doubles per thread over the whole block in a fixed tree order; result valid in all
// threads. `scratch` needs NV * 32 doubles. Block size must be a multiple of 32, <= 1024.
template <int NV>
__device__ __forceinline__ void block_reduce_sum(double (&v)[NV], double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
  __syncthreads();  // protect scratch from a previous use
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) scratch[i * 32 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; i++) {
    double x = (lane < nwarp) ? scratch[i * 32 + lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    v[i] = x;
  }
}

#endif  // SBTE_HOST_EMUL

}  // namespace sbte
