// Device side of the library-owned halo exchange (csrc/halo_ctx.cu): the strip copy with the neighbour handshake inside,
// as a __device__ function so that it can run as its own kernel (k_halo.cu k_halo_exchange) or as the first phase of a
// stencil kernel (k_fv_tma.cu / k_fv_stream.cu, b2s_halo_fv_tp2d: halo update + transport in ONE launch).
#pragma once
#include "common.cuh"

namespace b2s {
namespace impl {

// Everything the exchange needs, by value in kernel parameters.  links == nullptr: no exchange.
struct HaloXchg {
  const int64_t* links;       // [nlinks, 12], sorted by destination sub-domain
  const int64_t* peer_flags;  // [world] address of every rank's int32 flag array as mapped here
  const int* b_total;         // [64] work units (link, level) per destination sub-domain
  int* state;                 // [0] epoch, [1] blocks done, [2] status, [32 + b] units done, [128 + b] gate b, [200..] timeline
  void* dst;                  // this rank's field
  unsigned long long peers;   // bit r set: some link reads rank r's field (the ranks whose announcement is awaited)
  int nlinks, nk, my_rank, world, gated;
  int single_wait;            // != 0: every block polls the peers' flags itself (the first form of the handshake wait)
};

static constexpr int kExchangeWords = 12;
static constexpr long long kSyncTimeoutCycles = 4000000000LL;
static constexpr int kReadyWord = 3;  // epoch for which block 0 has seen every awaited announcement (kHandshake 1)
static constexpr int kDoneWord = 32;
static constexpr int kGateWord = 128;
static constexpr int kTraceWord = 200;  // uint64 timeline slots (tma.cuh gate_trace): [0] start, [1] end, [2] gate 0 opened

__device__ __forceinline__ void trace_ns(int* state, int slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  reinterpret_cast<unsigned long long*>(state + kTraceWord)[slot] = t;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_relaxed_sys(const int* p) {
  int v;
  asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// acquire / release fences (not the sequentially consistent __threadfence*: nothing here needs SC ordering)
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__device__ __forceinline__ void strip_decode(int t, int nd, int np, int64_t ssd, int& d, int& p) {
  if (ssd == 1 || ssd == -1) {  // depth runs along i on the source side
    d = t % nd;
    p = t / nd;
  } else {
    p = t % np;
    d = t / np;
  }
}

// Wait until every rank in `mask` has announced `epoch` (flag array at `offset`: 0 = announcements, 64 = deliveries).
// Called by every thread of the block; contains block barriers.
//   * block 0: one thread per awaited rank polls this GPU's own flag array with system-scope relaxed loads, one
//     system-scope acquire fence each, then thread 0 publishes the epoch in state[ready_word] (release, GPU scope);
//   * every other block: ONE thread polls that word (a GPU-scope load that hits this GPU's L2) and fences at GPU scope.
// The first version had every block poll the flags and fence at system scope: up to seven MEMBAR.SYS per block, hundreds
// of blocks hammering the very lines the peers' NVLink writes must land in.  The announcements happen-before block 0's
// fence, which happens-before its release, which the other blocks acquire: their loads of peer memory are ordered
// after the peers' writes by causality.  `single` (a plan option) keeps the per-block form for A/B runs.
__device__ __forceinline__ void await_flags(const int64_t* peer_flags, int my_rank, int world, unsigned long long mask, int offset,
                                            int epoch, int* state, int ready_word, bool single) {
  const int nthreads = blockDim.x;
  if (single || blockIdx.x == 0) {
    for (int r = threadIdx.x; r < world; r += nthreads)
      if ((mask >> r) & 1ull) {
        const int* mine = reinterpret_cast<const int*>(static_cast<uintptr_t>(peer_flags[my_rank])) + offset + r;
        const long long t0 = clock64();
        while (ld_relaxed_sys(mine) < epoch) {
          if (clock64() - t0 > kSyncTimeoutCycles) {
            atomicExch(state + 2, 1);
            break;
          }
        }
        fence_acq_rel_sys();
      }
    __syncthreads();
    if (!single && threadIdx.x == 0) st_release_gpu(state + ready_word, epoch);
  } else {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      while (ld_relaxed_gpu(state + ready_word) < epoch) {
        if (clock64() - t0 > kSyncTimeoutCycles) {
          atomicExch(state + 2, 1);
          break;
        }
      }
      fence_acq_rel_gpu();
    }
    __syncthreads();
  }
}

// Called by EVERY thread of EVERY block of the grid (it contains block barriers).  A work unit is one (link, level)
// strip; block j takes units j, j + gridDim.x, ... in link order, so sub-domains complete -- and their gates open --
// one after the other.  Protocol: see k_halo.cu.
// s_scratch: one int of shared memory (the caller's, so that kernels with a dynamic TMA window keep it unpadded).
template <typename T>
__device__ __forceinline__ void halo_exchange_body_inl(const HaloXchg& X, int* s_scratch) {
  const int nthreads = blockDim.x;
  int* state = X.state;
  if (threadIdx.x == 0) *s_scratch = *reinterpret_cast<volatile int*>(state) + 1;
  __syncthreads();
  const int epoch = *s_scratch;
  if (X.world > 1) {
    // block 0 announces "my field is final for this epoch" to every peer (one thread per peer) ...
    if (blockIdx.x == 0)
      for (int r = threadIdx.x; r < X.world; r += nthreads)
        if (r != X.my_rank) st_release_sys(reinterpret_cast<int*>(static_cast<uintptr_t>(X.peer_flags[r])) + X.my_rank, epoch);
    // ... and every block waits until the announcements of all the ranks it may read have arrived (await_flags)
    await_flags(X.peer_flags, X.my_rank, X.world, X.peers, 0, epoch, state, kReadyWord, X.single_wait != 0);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_ns(state, 0);
  const int nk = X.nk, units = X.nlinks * nk;
  T* dst = static_cast<T*>(X.dst);
  int cur_b = -1, cur_n = 0;  // units of sub-domain cur_b this block has copied and not yet reported
  // report: every thread's stores of those units are done (block barrier), visible device-wide (fence), counted; the
  // block that completes a sub-domain's count opens its gate
  auto report = [&]() {
    if (!X.gated || cur_n == 0) return;
    __syncthreads();
    if (threadIdx.x == 0) {
      fence_acq_rel_gpu();
      if (atomicAdd(state + kDoneWord + cur_b, cur_n) + cur_n == X.b_total[cur_b]) {
        state[kDoneWord + cur_b] = 0;
        st_release_gpu(state + kGateWord + cur_b, 1);
        if (cur_b == 0) trace_ns(state, 2);
      }
    }
  };
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int link = u / nk, k0 = u - link * nk;
    const int64_t* L = X.links + (int64_t)link * kExchangeWords;
    const int dst_b = (int)(L[11] >> 16);
    if (dst_b != cur_b) {
      report();
      cur_b = dst_b, cur_n = 0;
    }
    const int nd = (int)L[8], np = (int)L[9], n = nd * np;
    const T* src = reinterpret_cast<const T*>(static_cast<uintptr_t>(L[10])) + (L[0] + k0 * L[3]);
    T* out = dst + (L[4] + k0 * L[7]);
    const int64_t ssd = L[1], ssp = L[2], dsd = L[5], dsp = L[6];
    for (int t0 = threadIdx.x; t0 < n; t0 += 4 * nthreads) {  // four independent loads in flight per thread
      T v[4];
      int d[4], p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = t0 + i * nthreads;
        if (t < n) {
          strip_decode(t, nd, np, ssd, d[i], p[i]);
          v[i] = src[d[i] * ssd + p[i] * ssp];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (t0 + i * nthreads < n) out[d[i] * dsd + p[i] * dsp] = v[i];
    }
    ++cur_n;
  }
  report();
  // the last block of the launch advances the epoch for the next launch / graph replay
  __syncthreads();
  if (threadIdx.x == 0) {
    fence_acq_rel_gpu();
    if (atomicAdd(state + 1, 1) == (int)gridDim.x - 1) {
      state[1] = 0;
      trace_ns(state, 1);
      fence_acq_rel_gpu();
      *reinterpret_cast<volatile int*>(state) = epoch;
    }
  }
}

// ---- version 2 of the exchange body (the standalone kernel k_halo_exchange2, k_halo.cu) ----------------------------
// Same protocol, same table, same gates; what changed is how the copy is fed:
//   * a work unit is (link, chunk of `ku` levels) and a thread issues up to EIGHT independent loads before its first
//     store -- at the 8-GPU sub-domain size the whole update (4 MB) is in flight after one batch, so its duration is
//     one (NVLink) round trip instead of one per level;
//   * the link table is staged in shared memory once per block (it was a dependent global load per unit);
//   * the neighbours' announcements are awaited only before the first unit that reads a PEER: the same-GPU strips
//     (sorted first inside each sub-domain) are copied while the announcements are still on their way.
// s_links: kMaxCachedLinks * kExchangeWords int64 of shared memory; s_epoch: one int of shared memory.
static constexpr int kMaxCachedLinks = 64;
static constexpr int kLoadsInFlight = 8;
static constexpr int kMaxLevelsPerUnit = 16;

template <typename T>
__device__ __forceinline__ void halo_exchange_body2(const HaloXchg& X, int ku, int* s_epoch, int64_t* s_links) {
  const int nthreads = blockDim.x;
  int* state = X.state;
  if (threadIdx.x == 0) *s_epoch = *reinterpret_cast<volatile int*>(state) + 1;
  const bool cached = X.nlinks <= kMaxCachedLinks;
  if (cached)
    for (int w = threadIdx.x; w < X.nlinks * kExchangeWords; w += nthreads) s_links[w] = X.links[w];
  __syncthreads();
  const int epoch = *s_epoch;
  const int64_t* table = cached ? s_links : X.links;
  // block 0 announces "my field is final for this epoch" to every peer (one thread per peer)
  if (X.world > 1 && blockIdx.x == 0)
    for (int r = threadIdx.x; r < X.world; r += nthreads)
      if (r != X.my_rank) st_release_sys(reinterpret_cast<int*>(static_cast<uintptr_t>(X.peer_flags[r])) + X.my_rank, epoch);
  bool awaited = X.world <= 1;
  // every awaited peer at once: one thread per peer polls this GPU's own flag array, one acquire fence at the end
  auto await_peers = [&]() {
    await_flags(X.peer_flags, X.my_rank, X.world, X.peers, 0, epoch, state, kReadyWord, X.single_wait != 0);
    awaited = true;
  };
  if (!awaited && blockIdx.x == 0) await_peers();  // block 0 relays the announcements to the other blocks: at once
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_ns(state, 0);
  const int nk = X.nk, nchunks = (nk + ku - 1) / ku, units = X.nlinks * nchunks;
  T* dst = static_cast<T*>(X.dst);
  int cur_b = -1, cur_n = 0;  // (link, level) strips of sub-domain cur_b this block has copied and not yet reported
  auto report = [&]() {
    if (!X.gated || cur_n == 0) return;
    __syncthreads();
    if (threadIdx.x == 0) {
      fence_acq_rel_gpu();
      if (atomicAdd(state + kDoneWord + cur_b, cur_n) + cur_n == X.b_total[cur_b]) {
        state[kDoneWord + cur_b] = 0;
        st_release_gpu(state + kGateWord + cur_b, 1);
        if (cur_b == 0) trace_ns(state, 2);
      }
    }
  };
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int link = u / nchunks, k0 = (u - link * nchunks) * ku;
    const int nlev = min(ku, nk - k0);
    const int64_t* L = table + (int64_t)link * kExchangeWords;
    const int dst_b = (int)(L[11] >> 16);
    if (dst_b != cur_b) {
      report();
      cur_b = dst_b, cur_n = 0;
    }
    if (!awaited && (L[11] & 0xffff) != 0) await_peers();  // block-uniform: L is the same for every thread
    const int nd = (int)L[8], np = (int)L[9], n = nd * np, total = n * nlev;
    // 32-bit strides: b2s_halo_plan has checked that every offset relative to the unit's first element fits (kNarrow)
    const int ssd = (int)L[1], ssp = (int)L[2], ssk = (int)L[3], dsd = (int)L[5], dsp = (int)L[6], dsk = (int)L[7];
    const T* src = reinterpret_cast<const T*>(static_cast<uintptr_t>(L[10])) + (L[0] + k0 * L[3]);
    T* out = dst + (L[4] + k0 * L[7]);
    for (int e0 = threadIdx.x; e0 < total; e0 += kLoadsInFlight * nthreads) {
      T v[kLoadsInFlight];
      int o[kLoadsInFlight];
#pragma unroll
      for (int i = 0; i < kLoadsInFlight; ++i) {
        const int e = e0 + i * nthreads;
        if (e < total) {
          const int kk = e / n, t = e - kk * n;
          int d, p;
          strip_decode(t, nd, np, ssd, d, p);
          v[i] = src[d * ssd + p * ssp + kk * ssk];
          o[i] = d * dsd + p * dsp + kk * dsk;
        }
      }
#pragma unroll
      for (int i = 0; i < kLoadsInFlight; ++i)
        if (e0 + i * nthreads < total) out[o[i]] = v[i];
    }
    cur_n += nlev;
  }
  report();
  if (!awaited) await_peers();  // a block without peer strips still may not let the kernel end before the neighbours announced
  // the last block of the launch advances the epoch for the next launch / graph replay
  __syncthreads();
  if (threadIdx.x == 0) {
    fence_acq_rel_gpu();
    if (atomicAdd(state + 1, 1) == (int)gridDim.x - 1) {
      state[1] = 0;
      trace_ns(state, 1);
      fence_acq_rel_gpu();
      *reinterpret_cast<volatile int*>(state) = epoch;
    }
  }
}

// ---- version 3: the MIXED exchange (k_halo_exchange3, k_halo.cu) -- same-GPU strips are pulled, strips that cross
// NVLink are PUSHED by the rank that owns the source -----------------------------------------------------------------
// Why: loads on a peer address are round trips, and the link keeps only so many of them in flight -- measured on two
// B200s the in-place pull of 4 MB of strips took ~30 us (~130 GB/s) with four or with eight loads in flight per thread
// (profiles/README.md, round 2).  Stores are posted: the owner reads its own strip (local) and writes it into the
// neighbour's halo; nothing waits for a round trip except the two flag hops.
// Protocol per epoch n (A and D are int32 flag arrays in peer-mapped memory, one slot per rank):
//   1. block 0 writes A[me] = n at every peer: "my field is final for epoch n AND my halos may be overwritten"
//      (the previous consumer of the halos ran earlier in stream order);
//   2. a unit that touches a peer (a pull FROM it or a push INTO it) first waits for A[peer] >= n;
//   3. push units of peer r are counted; the block that completes the count of r fences (sys) and writes D[me] = n at r:
//      "everything I owe you for epoch n has landed";
//   4. before it ends, every block waits for D[r] >= n from every rank r that pushes into this one.
// When the kernel has ended, the halos are complete (the stencil that follows in stream order may read them), and every
// neighbour has announced epoch n, i.e. has finished reading what it needed of epoch n-1 -- the same guarantee a
// ping-pong time loop got from the pull protocol.
// Row words: 0..9 as in the link table, [10] base address of the source buffer, [11] = (peer + 1, 0 for a same-GPU
// strip) | destination sub-domain << 16 | push << 40, [12] base address of the destination buffer, [13] unused.
struct HaloXchg3 {
  const int64_t* rows;        // [nrows, 14]
  const int64_t* peer_flags;  // [world] address of every rank's flag array as mapped here: A at [0, 64), D at [64, 128)
  const int* push_total;      // [world] (link, level) strips this rank pushes to each peer
  int* state;                 // [0] epoch, [1] blocks done, [2] status, [kPushWord + r] strips pushed to r so far
  unsigned long long wait_a;  // ranks whose announcement is awaited: sources of pulls across GPUs and targets of pushes
  unsigned long long wait_d;  // ranks that push into this one
  int nrows, nk, my_rank, world;
  int nrows1;                 // rows [0, nrows1) run before the deliveries are awaited, rows [nrows1, nrows) after (the
                              // unpacking of what the peers delivered into this rank's staging buffer)
};
static constexpr int kMixedWords = 14;
static constexpr int kPushWord = 320;
static constexpr int kDeliveredOffset = 64;  // D flags follow the A flags in the flag array

template <typename T>
__device__ __forceinline__ void halo_exchange_body3(const HaloXchg3& X, int ku, int* s_epoch, int64_t* s_rows) {
  const int nthreads = blockDim.x;
  int* state = X.state;
  if (threadIdx.x == 0) *s_epoch = *reinterpret_cast<volatile int*>(state) + 1;
  const bool cached = X.nrows <= kMaxCachedLinks;
  if (cached)
    for (int w = threadIdx.x; w < X.nrows * kMixedWords; w += nthreads) s_rows[w] = X.rows[w];
  __syncthreads();
  const int epoch = *s_epoch;
  const int64_t* table = cached ? s_rows : X.rows;
  if (X.world > 1 && blockIdx.x == 0)
    for (int r = threadIdx.x; r < X.world; r += nthreads)
      if (r != X.my_rank) st_release_sys(reinterpret_cast<int*>(static_cast<uintptr_t>(X.peer_flags[r])) + X.my_rank, epoch);
  // one thread per awaited rank polls this GPU's own flag array (A: offset 0, D: offset 64), one acquire fence each
  auto await = [&](unsigned long long mask, int offset) {
    await_flags(X.peer_flags, X.my_rank, X.world, mask, offset, epoch, state, offset == 0 ? kReadyWord : kReadyWord + 1, false);
  };
  bool announced = X.world <= 1 || X.wait_a == 0ull;
  if (!announced && blockIdx.x == 0) {  // block 0 relays the announcements to the other blocks: at once
    await(X.wait_a, 0);
    announced = true;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_ns(state, 0);
  const int nk = X.nk, nchunks = (nk + ku - 1) / ku;
  auto run_rows = [&](int row0, int row1) {
  const int units = (row1 - row0) * nchunks;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int row = row0 + u / nchunks, k0 = (u % nchunks) * ku;
    const int nlev = min(ku, nk - k0);
    const int64_t* L = table + (int64_t)row * kMixedWords;
    const int peer = (int)(L[11] & 0xffff) - 1;
    const bool push = ((L[11] >> 40) & 1) != 0;
    if (!announced && peer >= 0) {  // block-uniform
      await(X.wait_a, 0);
      announced = true;
    }
    const int nd = (int)L[8], np = (int)L[9], n = nd * np, total = n * nlev;
    const int ssd = (int)L[1], ssp = (int)L[2], ssk = (int)L[3], dsd = (int)L[5], dsp = (int)L[6], dsk = (int)L[7];
    const T* src = reinterpret_cast<const T*>(static_cast<uintptr_t>(L[10])) + (L[0] + k0 * L[3]);
    T* out = reinterpret_cast<T*>(static_cast<uintptr_t>(L[12])) + (L[4] + k0 * L[7]);
    // consecutive threads follow whichever of (d, p) is contiguous on the REMOTE side: the source of a pull, the
    // destination of a push
    const int order_stride = push ? dsd : ssd;
    for (int e0 = threadIdx.x; e0 < total; e0 += kLoadsInFlight * nthreads) {
      T v[kLoadsInFlight];
      int o[kLoadsInFlight];
#pragma unroll
      for (int i = 0; i < kLoadsInFlight; ++i) {
        const int e = e0 + i * nthreads;
        if (e < total) {
          const int kk = e / n, t = e - kk * n;
          int d, p;
          strip_decode(t, nd, np, order_stride, d, p);
          v[i] = src[d * ssd + p * ssp + kk * ssk];
          o[i] = d * dsd + p * dsp + kk * dsk;
        }
      }
#pragma unroll
      for (int i = 0; i < kLoadsInFlight; ++i)
        if (e0 + i * nthreads < total) out[o[i]] = v[i];
    }
    if (push) {
      __syncthreads();  // every thread's stores of this unit are issued ...
      if (threadIdx.x == 0) {
        fence_acq_rel_sys();  // ... and ordered, system-wide, before the count that may release the delivery flag
        if (atomicAdd(state + kPushWord + peer, nlev) + nlev == X.push_total[peer]) {
          state[kPushWord + peer] = 0;
          fence_acq_rel_sys();
          st_release_sys(reinterpret_cast<int*>(static_cast<uintptr_t>(X.peer_flags[peer])) + kDeliveredOffset + X.my_rank, epoch);
        }
      }
    }
  }
  };
  run_rows(0, X.nrows1);
  if (!announced) await(X.wait_a, 0);
  if (X.wait_d != 0ull) await(X.wait_d, kDeliveredOffset);
  run_rows(X.nrows1, X.nrows);  // same-GPU copies out of the staging buffer the peers have just filled
  // the last block of the launch advances the epoch for the next launch / graph replay
  __syncthreads();
  if (threadIdx.x == 0) {
    fence_acq_rel_gpu();
    if (atomicAdd(state + 1, 1) == (int)gridDim.x - 1) {
      state[1] = 0;
      trace_ns(state, 1);
      fence_acq_rel_gpu();
      *reinterpret_cast<volatile int*>(state) = epoch;
    }
  }
}

// The stencil kernels call the exchange as a REAL function: inlined, its register needs and code size changed the
// allocation of the consumers' hot loop (tile kernel 91 -> 79 registers, +8 % run time at 3 x 192 x 192 x 72 even with
// the exchange switched off, measured A/B against the round-1 library on one box).
template <typename T>
__device__ __noinline__ void halo_exchange_call(const HaloXchg* X, int* s_scratch) {
  halo_exchange_body_inl<T>(*X, s_scratch);
}

}  // namespace impl
}  // namespace b2s
