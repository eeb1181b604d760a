// mbarrier + TMA (cp.async.bulk.tensor) PTX wrappers and the host-side tensor-map cache shared by
// the TMA kernels (k_fv_tma.cu, k_remap_slab.cu).  sm_100a only.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace b2s {
namespace impl {

// ---- PTX wrappers (mbarrier + TMA) ------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---- halo gate (csrc/halo_ctx.cu, k_halo.cu k_halo_exchange) -----------------------------------
// gate[b], b < kGateSlots: raised (st.release.gpu) by the exchange kernel when every halo cell of sub-domain b has
// been written; gate[kGateSlots]: CTAs of the consuming stencil that have finished; gate[kGateSlots + 1]: status
// (1 = a wait gave up).  A gated stencil walks its items in sub-domain order; its producer lane acquires gate[b]
// before the first TMA load of sub-domain b.  The halo cells were written through the generic proxy by another
// kernel and are read through the async proxy, hence the proxy fence.  Bounded spin: a protocol error becomes a
// status word (b2s_halo_status), not a hung GPU.
static constexpr int kGateSlots = 64;
static constexpr int kGateTraceOffset = 72;  // int offset from the gate words to the uint64 timeline slots (diagnostics)

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// timeline of the overlapped step, written by a handful of threads (b2s_halo_trace): [0] exchange start, [1] exchange
// end, [2] gate 0 opened, [3] first stencil CTA started, [4] stencil CTA 0 got gate 0, [5] last stencil CTA finished
__device__ __forceinline__ void gate_trace(int* gate, int slot) {
  reinterpret_cast<unsigned long long*>(gate + kGateTraceOffset)[slot] = global_timer_ns();
}

__device__ __forceinline__ void gate_acquire(int* gate, int b) {
  int v;
  const long long t0 = clock64();
  for (;;) {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(gate + b) : "memory");
    if (v != 0) break;
    if (clock64() - t0 > 4000000000LL) {
      atomicExch(gate + kGateSlots + 1, 1);
      break;
    }
    __nanosleep(64);
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");
}
// called by one thread per CTA when the CTA is done: the last CTA lowers the gates for the next exchange
__device__ __forceinline__ void gate_release(int* gate, int nb, int nctas) {
  __threadfence();
  if (atomicAdd(gate + kGateSlots, 1) == nctas - 1) {
    gate[kGateSlots] = 0;
    for (int b = 0; b < nb; ++b) gate[b] = 0;
    gate_trace(gate, 5);
    __threadfence();
  }
}

// cp.async (LDGSTS): one naturally aligned element global -> shared without a register round trip
template <int BYTES>
__device__ __forceinline__ void cp_async(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- host side: tensor maps (tma_host.cu) -----------------------------------------------------

// A field as TMA sees it: 16-byte aligned base `p - off` (off elements), extents in elements.
template <typename T>
struct TmaField {
  const T* base;
  int off;
  bool ok;
};

template <typename T>
inline TmaField<T> tma_field(const T* first, int64_t sj, int64_t sk, int64_t sb, int nk, int nb) {
  constexpr int V = 16 / sizeof(T);
  TmaField<T> f;
  const uintptr_t a = reinterpret_cast<uintptr_t>(first);
  f.off = static_cast<int>((a % 16) / sizeof(T));
  f.base = first - f.off;
  f.ok = (a % sizeof(T) == 0) && sj > 0 && sj % V == 0 && (nk == 1 || (sk > 0 && sk % V == 0)) &&
         (nb == 1 || (sb > 0 && sb % V == 0));
  return f;
}

// 4-D (i, j, k, b) tiled tensor map over element strides (1, sj, sk, sb) with extents e0..e3 and
// box b0 x b1 x b2 x 1; cached by (base, strides, extents, box).  False = the driver refused it.
bool make_tensor_map(CUtensorMap* out, const void* base, int elem_size, int64_t sj, int64_t sk, int64_t sb, int e0,
                     int e1, int e2, int e3, int b0, int b1, int b2);

template <typename T>
inline bool make_map(CUtensorMap* out, const T* base, int64_t sj, int64_t sk, int64_t sb, int e0, int e1, int e2,
                     int e3, int b0, int b1, int b2 = 1) {
  return make_tensor_map(out, base, (int)sizeof(T), sj, sk, sb, e0, e1, e2, e3, b0, b1, b2);
}

}  // namespace impl
}  // namespace b2s
