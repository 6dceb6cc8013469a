// tests/emul/cuda_emul.h -- TEST INFRASTRUCTURE ONLY.  A small host shim that lets g++ compile the batched
// convolution kernels of spectralbte_b200/csrc (plain C++ once the CUDA keywords expand to nothing) and run one CTA
// at a time with one OS thread per CUDA thread: mbarriers (arrival + transaction counts, phase parity), TMA bulk
// and 2-D tensor copies (synchronous memcpy + complete_tx), __syncthreads / __syncwarp, dynamic shared memory.
// It checks what cannot be checked by arithmetic alone before a kernel has run on a GPU: the barrier protocol
// (a wrong count hangs -> bounded wait aborts), the stream-K walk, tile switches, flushes, TMA coordinates and the
// shared-memory layout.  It says nothing about performance.
#pragma once
#define SBTE_HOST_EMUL 1
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

// CUDA keywords the host compiler has not been told about (cuda_runtime.h defines most of them away already)
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#ifndef __grid_constant__
#define __grid_constant__
#endif
#ifndef __global__
#define __global__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __host__
#define __host__
#endif
#ifndef __forceinline__
#define __forceinline__ inline __attribute__((always_inline))
#endif

namespace emul {

// ---- per-thread CUDA built-ins
struct Barrier {   // reusable barrier for __syncthreads / __syncwarp
  std::mutex m;
  std::condition_variable cv;
  int count = 0, waiting = 0;
  unsigned gen = 0;
  void init(int n) { count = n; waiting = 0; gen = 0; }
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned g = gen;
    if (++waiting == count) { waiting = 0; gen++; cv.notify_all(); return; }
    cv.wait(lk, [&] { return gen != g; });
  }
  void leave() {   // a thread that exits the kernel no longer takes part
    std::unique_lock<std::mutex> lk(m);
    count--;
    if (count > 0 && waiting == count) { waiting = 0; gen++; cv.notify_all(); }
  }
};

struct Cta {
  unsigned char* smem = nullptr;
  Barrier all;
  std::vector<Barrier> warps;
  std::atomic<bool> failed{false};
};

inline thread_local Cta* cta = nullptr;
inline thread_local int warp_id = 0;

// ---- mbarrier emulation: the 8 bytes the kernel reserves hold a pointer to this object
struct MBar {
  std::mutex m;
  int count = 0, pending = 0;
  long long tx = 0;
  unsigned phase = 0;   // number of completed phases
  void maybe_complete() {
    if (pending == 0 && tx == 0) { phase++; pending = count; }
  }
};
inline std::vector<MBar*>& all_bars() { static std::vector<MBar*> v; return v; }
inline std::mutex& bars_mutex() { static std::mutex m; return m; }
inline MBar* bar_of(uint64_t* p) { return reinterpret_cast<MBar*>(*p); }

// fake tensor map: what tma_tensor2d_g2s needs (stored in the first bytes of a CUtensorMap)
struct FakeTensorMap {
  const double* base;
  long long dim0, dim1;   // innermost first (elements)
  int box0, box1;
};

}  // namespace emul

static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { emul::cta->all.wait(); }
inline void __syncwarp() { emul::cta->warps[emul::warp_id].wait(); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
[[noreturn]] inline void __trap() {
  fprintf(stderr, "emul: __trap() in block %u thread %u (bounded wait expired: lost arrival / transaction)\n", blockIdx.x, threadIdx.x);
  fflush(stderr);
  _Exit(3);
}

#define SBTE_DYN_SMEM(name) unsigned char* name = emul::cta->smem
#define SBTE_SETMAXNREG_DEC(n) ((void)0)
#define SBTE_SETMAXNREG_INC(n) ((void)0)

namespace sbte {

inline void mbar_init(uint64_t* bar, uint32_t count) {
  emul::MBar* b = new emul::MBar();
  b->count = (int)count; b->pending = (int)count;
  { std::lock_guard<std::mutex> lk(emul::bars_mutex()); emul::all_bars().push_back(b); }
  *bar = reinterpret_cast<uint64_t>(b);
}
inline void mbar_fence_init() {}
inline void fence_proxy_async() {}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  emul::MBar* b = emul::bar_of(bar);
  std::lock_guard<std::mutex> lk(b->m);
  b->tx += bytes; b->pending--;
  if (b->pending < 0) { fprintf(stderr, "emul: too many arrivals on a barrier\n"); _Exit(4); }
  b->maybe_complete();
}
inline void mbar_arrive_cnt(uint64_t* bar, uint32_t count) {
  emul::MBar* b = emul::bar_of(bar);
  std::lock_guard<std::mutex> lk(b->m);
  b->pending -= (int)count;
  if (b->pending < 0) { fprintf(stderr, "emul: too many arrivals on a barrier\n"); _Exit(4); }
  b->maybe_complete();
}
inline void mbar_arrive(uint64_t* bar) { mbar_arrive_cnt(bar, 1); }
inline void mbar_complete_tx(uint64_t* bar, long long bytes) {
  emul::MBar* b = emul::bar_of(bar);
  std::lock_guard<std::mutex> lk(b->m);
  b->tx -= bytes;
  b->maybe_complete();
}
// try_wait.parity P succeeds once the phase of parity P has completed
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  emul::MBar* b = emul::bar_of(bar);
  std::lock_guard<std::mutex> lk(b->m);
  return (b->phase & 1u) != (parity & 1u);
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  unsigned long long spins = 0;
  const auto t0 = std::chrono::steady_clock::now();
  while (!mbar_try_wait(bar, parity)) {
    std::this_thread::yield();
    // a deadlocked protocol must fail the test, not hang it: no single wait of these small problems takes 30 s
    if ((++spins & 1023) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30)) __trap();
  }
}
inline void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  if ((reinterpret_cast<uintptr_t>(smem_dst) - reinterpret_cast<uintptr_t>(emul::cta->smem)) % 16 != 0 || bytes % 16 != 0 ||
      reinterpret_cast<uintptr_t>(gsrc) % 16 != 0) {
    fprintf(stderr, "emul: misaligned bulk copy\n"); _Exit(5);
  }
  memcpy(smem_dst, gsrc, bytes);
  mbar_complete_tx(bar, bytes);
}
inline void tma_tensor2d_g2s(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  const emul::FakeTensorMap* t = reinterpret_cast<const emul::FakeTensorMap*>(tmap);
  if ((reinterpret_cast<uintptr_t>(smem_dst) - reinterpret_cast<uintptr_t>(emul::cta->smem)) % 128 != 0) {
    fprintf(stderr, "emul: tensor copy destination not 128-byte aligned\n"); _Exit(5);
  }
  double* dst = reinterpret_cast<double*>(smem_dst);
  for (int r = 0; r < t->box1; r++)
    for (int c = 0; c < t->box0; c++) {
      const long long gr = (long long)c1 + r, gc = (long long)c0 + c;
      dst[(size_t)r * t->box0 + c] = (gr < t->dim1 && gc < t->dim0 && gr >= 0 && gc >= 0) ? t->base[gr * t->dim0 + gc] : 0.0;
    }
  mbar_complete_tx(bar, (long long)t->box0 * t->box1 * 8);
}
inline double2 ldg_stream_f64x2(const double* p) { return make_double2(p[0], p[1]); }
inline void red_add_f64(double* p, double v) { *p += v; }   // single writer per location (see common.cuh)
// hand-over of a running sum between two CTAs (common.cuh carry_*): CTAs run one after the other here, in launch order, so
// a consumer always finds the word published; anything else is a protocol error (wrong slot, missing publication)
inline void carry_publish(int* flag) { __atomic_store_n(flag, 1, __ATOMIC_RELEASE); }
inline void carry_await(const int* flag) {
  if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) == 0) {
    fprintf(stderr, "emul: block %u warp %d waits for a running sum nobody published\n", blockIdx.x, emul::warp_id);
    fflush(stderr);
    _Exit(6);
  }
}
inline void carry_rearm(int* flag) { __atomic_store_n(flag, 0, __ATOMIC_RELAXED); }
inline double2 carry_load(const double2* p) { return *p; }
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}

}  // namespace sbte

namespace emul {

// Runs one CTA of `threads` CUDA threads: body(tid) is the kernel call.
template <typename F>
void run_cta(unsigned block, unsigned nblocks, int threads, size_t smem_bytes, F body) {
  Cta c;
  std::vector<unsigned char> smem(smem_bytes + 256, 0xCD);   // poisoned: unwritten shared memory shows up as garbage
  c.smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem.data()) + 127) & ~uintptr_t(127));
  c.all.init(threads);
  c.warps = std::vector<Barrier>(threads / 32);
  for (auto& w : c.warps) w.init(32);
  std::vector<std::thread> ts;
  ts.reserve(threads);
  for (int tid = 0; tid < threads; tid++) {
    ts.emplace_back([&, tid] {
      cta = &c;
      warp_id = tid / 32;
      threadIdx = {(unsigned)tid, 0, 0};
      blockIdx = {block, 0, 0};
      blockDim = dim3(threads, 1, 1);
      gridDim = dim3(nblocks, 1, 1);
      body(tid);
      c.warps[tid / 32].leave();
      c.all.leave();
    });
  }
  for (auto& t : ts) t.join();
  std::lock_guard<std::mutex> lk(bars_mutex());
  for (MBar* b : all_bars()) delete b;
  all_bars().clear();
}

}  // namespace emul
