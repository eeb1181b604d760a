// K7 halo_move: pack / unpack / same-GPU neighbour copy of halo strips, one kernel for all three.
// Replaces the pack-unpack half of NDSL's HaloUpdater (not in /root/reference; SURVEY.md 2.1 K7,
// 8e).  A "link" is an affine strip copy described by 10 int64 (device array `links`):
//   [0] src_off [1] src_sd [2] src_sp [3] src_sk   element offset/strides into `src`
//   [4] dst_off [5] dst_sd [6] dst_sp [7] dst_sk   element offset/strides into `dst`
//   [8] nd  (halo depth, 3)   [9] np (edge length)
//   dst[dst_off + d*dst_sd + p*dst_sp + k*dst_sk] = src[src_off + d*src_sd + p*src_sp + k*src_sk]
// Negative strides express the index reversal / rotation across cubed-sphere tile edges.
//   pack:   src = field, dst = send buffer      unpack: src = recv buffer, dst = field
//   local:  src = dst = field (neighbouring sub-domains resident on the same GPU)
#include "halo_device.cuh"
#include "impl.cuh"

namespace b2s {
namespace impl {

static constexpr int kLinkWords = 10;

// grid = (strip chunks, k chunks, link): one thread per strip element and KU levels.  The faster-varying
// thread index follows whichever of (d, p) is contiguous on the SOURCE side, so reads coalesce for
// north/south strips (p along i) and stay sector-dense for west/east ones (d along i: 3 adjacent elements).
// KU = 1 is the default: measured on C384x72 (profiles/r01_halo_kernels.md) 11.9 us per update against 13.2
// (KU = 4) and 14.2 (KU = 8), fp32 the same as fp64 -- the copy is bound by the number of 32-byte sectors the
// west/east strips touch (3 elements per 3 KB row), not by bytes or by loads in flight.  With corner blocks
// in the table (24 extra 3x3 links) KU = 4 wins (13.7 vs 16.0 us): b2s_set_option("halo_levels", 1 | 4 | 8).
static constexpr int kKU = 1;

template <typename T, int KU>
__device__ __forceinline__ void copy_levels(const T* sp, int64_t ssk, T* dp, int64_t dsk, int nlev) {
  T v[KU];
  if (nlev >= KU) {
#pragma unroll
    for (int u = 0; u < KU; ++u) v[u] = sp[u * ssk];
#pragma unroll
    for (int u = 0; u < KU; ++u) dp[u * dsk] = v[u];
  } else {
#pragma unroll
    for (int u = 0; u < KU; ++u)
      if (u < nlev) v[u] = sp[u * ssk];
#pragma unroll
    for (int u = 0; u < KU; ++u)
      if (u < nlev) dp[u * dsk] = v[u];
  }
}

static int halo_levels() {
  const int ku = option("halo_levels", kKU);
  return ku == 4 || ku == 8 ? ku : kKU;
}

template <typename T, int KU>
__global__ void __launch_bounds__(256) k_halo_move(int nk, const int64_t* __restrict__ links, const T* src, T* dst) {
  const int64_t* L = links + (int64_t)blockIdx.z * kLinkWords;
  const int nd = (int)L[8], np = (int)L[9];
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= nd * np) return;
  const int64_t ssd = L[1], ssp = L[2], ssk = L[3], dsk = L[7];
  int d, p;
  strip_decode(t, nd, np, ssd, d, p);
  const int k0 = blockIdx.y * KU;
  copy_levels<T, KU>(src + (L[0] + d * ssd + p * ssp + k0 * ssk), ssk, dst + (L[4] + d * L[5] + p * L[6] + k0 * dsk), dsk, nk - k0);
}

template <typename T>
int halo_move(int nlinks, int nk, const int64_t* links, const T* src, T* dst, cudaStream_t s, int max_strip) {
  B2S_ARGCHECK(nlinks >= 0 && nk > 0, "halo_move: bad sizes nlinks=%d nk=%d", nlinks, nk);
  if (nlinks == 0) return B2S_OK;
  B2S_ARGCHECK(links && src && dst, "halo_move: null pointer");
  B2S_ARGCHECK(nk <= 65535 && nlinks <= 65535, "halo_move: grid too large (nk=%d, nlinks=%d)", nk, nlinks);
  const int ku = halo_levels();
  dim3 grid((max_strip + 255) / 256, (nk + ku - 1) / ku, nlinks);
  if (ku == 8) k_halo_move<T, 8><<<grid, 256, 0, s>>>(nk, links, src, dst);
  else if (ku == 4) k_halo_move<T, 4><<<grid, 256, 0, s>>>(nk, links, src, dst);
  else k_halo_move<T, 1><<<grid, 256, 0, s>>>(nk, links, src, dst);
  return check_launch("halo_move");
}

template <typename T>
int halo_move(int nlinks, int nk, int max_strip, const int64_t* links, const T* src, T* dst, cudaStream_t s) {
  B2S_ARGCHECK(max_strip > 0, "halo_move: max_strip must be the largest nd*np of the table, got %d", max_strip);
  return halo_move<T>(nlinks, nk, links, src, dst, s, max_strip);
}

// -------------------------------------------------------------------------------------------
// halo_pull: the whole halo update of a GPU as ONE kernel over NVLink peer memory.
// Same strip copies as halo_move, but every link carries the BASE ADDRESS of its source buffer in
// word [10]: the field of the GPU that owns the neighbouring sub-domain, mapped into this process
// (torch symmetric memory / CUDA IPC), or this GPU's own field for same-GPU neighbours.  Loads on a
// peer address travel over NVLink (SASS is an ordinary LDG; the address aperture routes it), so
// there is no pack buffer, no NCCL launch and no unpack: ordering against the peers' writes is a
// device-side barrier issued before this kernel (halo/p2p.py).
// -------------------------------------------------------------------------------------------
static constexpr int kPullWords = 11;

template <typename T, int KU>
__global__ void __launch_bounds__(256) k_halo_pull(int nk, const int64_t* __restrict__ links, T* dst, int words) {
  const int64_t* L = links + (int64_t)blockIdx.z * words;
  const int nd = (int)L[8], np = (int)L[9];
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= nd * np) return;
  const T* src = reinterpret_cast<const T*>(static_cast<uintptr_t>(L[10]));
  const int64_t ssd = L[1], ssp = L[2], ssk = L[3], dsk = L[7];
  int d, p;
  strip_decode(t, nd, np, ssd, d, p);
  const int k0 = blockIdx.y * KU;
  copy_levels<T, KU>(src + (L[0] + d * ssd + p * ssp + k0 * ssk), ssk, dst + (L[4] + d * L[5] + p * L[6] + k0 * dsk), dsk, nk - k0);
}

template <typename T>
int halo_pull(int nlinks, int nk, int max_strip, const int64_t* links, T* dst, cudaStream_t s) {
  B2S_ARGCHECK(nlinks >= 0 && nk > 0 && max_strip > 0, "halo_pull: bad sizes nlinks=%d nk=%d max_strip=%d", nlinks, nk, max_strip);
  if (nlinks == 0) return B2S_OK;
  B2S_ARGCHECK(links && dst, "halo_pull: null pointer");
  B2S_ARGCHECK(nk <= 65535 && nlinks <= 65535, "halo_pull: grid too large (nk=%d, nlinks=%d)", nk, nlinks);
  const int ku = halo_levels();
  dim3 grid((max_strip + 255) / 256, (nk + ku - 1) / ku, nlinks);
  if (ku == 8) k_halo_pull<T, 8><<<grid, 256, 0, s>>>(nk, links, dst, kPullWords);
  else if (ku == 4) k_halo_pull<T, 4><<<grid, 256, 0, s>>>(nk, links, dst, kPullWords);
  else k_halo_pull<T, 1><<<grid, 256, 0, s>>>(nk, links, dst, kPullWords);
  return check_launch("halo_pull");
}

// -------------------------------------------------------------------------------------------
// k_halo_exchange: the halo update of a b2s_halo context (csrc/halo_ctx.cu) as ONE launch -- halo_pull with the
// neighbour handshake inside, so there is no separate barrier kernel and no NCCL call on the path.
//   * every rank owns an int32 flag array flags[world] in peer-mapped memory; peer_flags[r] is the address of
//     rank r's array as mapped into this process;
//   * the step number (epoch) lives on the device (state[0]), so a CUDA graph can replay the launch: every block
//     reads it at entry, the last block to finish advances it (state[1] counts blocks);
//   * block 0 announces "my field is final for this epoch" to every peer: st.release.sys of the epoch into
//     flags[my_rank] of each peer (the field was written by earlier kernels in stream order);
//   * every block first waits until the announcements of all the ranks it may read have arrived -- one thread per awaited
//     peer polls this GPU's OWN flag array (relaxed loads, one acquire fence at the end) -- then pulls.  Announcements are monotonic, a rank can run at most one step ahead of
//     a neighbour, and -- adjacency being symmetric -- a neighbour's announcement of epoch n+1 also says it has
//     finished pulling epoch n from this rank, which is what a ping-pong time loop needs before overwriting;
//   * gated: links are sorted by destination sub-domain b and carry it in word [11]; the last block of the links
//     into sub-domain b raises gate[b] (st.release.gpu) -- blocks are dispatched in link order, so the gates open
//     one sub-domain after the other.  A gated stencil (fv_tp2d_gated) walks its items in the same order, acquires
//     gate[b] before its first load of sub-domain b and lowers the gates when it finishes: the exchange of
//     sub-domains 1.. overlaps the stencil of sub-domains 0.., with the stencil's DRAM-friendly item order untouched
//     (an interior-cells-first order was measured and costs 13-28 % of the stencil: profiles/r02_overlap.md);
//   * no wait is unbounded: after ~2 s of spinning a block records status 1 in state[2] and carries on
//     (wrong halos, reported by b2s_halo_status, instead of a hung GPU).
// Link words: 0..9 as halo_move, [10] base address of the source buffer, [11] = (owning rank + 1) | (b << 16) with
// owning rank -1 for this GPU and b the destination sub-domain (batch index).
// state: [0] epoch, [1] blocks done, [2] status, [kDoneWord + b] blocks done for sub-domain b, [kGateWord + b] gate b.
// -------------------------------------------------------------------------------------------
// PERSISTENT grid: gridDim.x = a few blocks per SM walk the (link, level) work units in link order.  A
// one-block-per-strip grid was measured first: thousands of 256-thread blocks fill every thread slot of the GPU, a
// stencil launched beside them cannot become resident until they drain, and the "overlapped" step was as long as the
// serial one (profiles/r02_overlap.md).  The body lives in halo_device.cuh: the fused step (b2s_halo_fv_tp2d) runs
// it as the first phase of the stencil kernel instead.
template <typename T>
__global__ void __launch_bounds__(256) k_halo_exchange(const HaloXchg X) {
  __shared__ int s_epoch;
  halo_exchange_body_inl<T>(X, &s_epoch);
}

// version 2 (halo_device.cuh halo_exchange_body2): ku levels per work unit, eight loads in flight per thread, link table in
// shared memory, announcements awaited only before the first peer strip.  b2s_set_option("halo_variant", 1) selects the
// first version for A/B runs.
template <typename T>
__global__ void __launch_bounds__(256, 4) k_halo_exchange2(const HaloXchg X, int ku) {
  __shared__ int s_epoch;
  __shared__ int64_t s_links[kMaxCachedLinks * kExchangeWords];
  halo_exchange_body2<T>(X, ku, &s_epoch, s_links);
}

// version 3 (halo_device.cuh halo_exchange_body3): same-GPU strips pulled, strips that cross NVLink pushed by their owner
template <typename T>
__global__ void __launch_bounds__(256, 4) k_halo_exchange3(const HaloXchg3 X, int ku) {
  __shared__ int s_epoch;
  __shared__ int64_t s_rows[kMaxCachedLinks * kMixedWords];
  halo_exchange_body3<T>(X, ku, &s_epoch, s_rows);
}

int halo_exchange3_launch(int elem_size, int max_strip, const HaloXchg3& X, cudaStream_t s) {
  B2S_ARGCHECK(X.world >= 1 && X.world <= 64 && X.my_rank >= 0 && X.my_rank < X.world, "halo_exchange: rank %d of %d", X.my_rank, X.world);
  B2S_ARGCHECK(X.peer_flags && X.state && X.push_total && X.nk > 0 && X.nrows >= 0 && (X.nrows == 0 || X.rows), "halo_exchange: bad mixed plan");
  int ku = option("halo_levels_per_unit", 0);
  if (ku <= 0 || ku > kMaxLevelsPerUnit) ku = 1;
  if (ku > X.nk) ku = X.nk;
  int per_sm = option("halo_blocks_per_sm", 0);
  if (per_sm <= 0 || per_sm > 4) per_sm = 4;
  const int64_t slots = (int64_t)sm_count() * per_sm;
  const int64_t units = (int64_t)X.nrows * ((X.nk + ku - 1) / ku);
  const int grid = (int)(units < 1 ? 1 : (units < slots ? units : slots));  // a rank without strips still announces and waits
  if (elem_size == 8)
    k_halo_exchange3<double><<<grid, 256, 0, s>>>(X, ku);
  else
    k_halo_exchange3<float><<<grid, 256, 0, s>>>(X, ku);
  return check_launch("halo_exchange");
}

// version 4 = the handshake as its OWN one-block kernel, then the plain flat-grid pull (k_halo_pull: one thread per strip
// element, no flags, no fences, no epoch logic): two launches, each as simple as it gets.  The kernel boundary orders
// the pull after the acquired announcements.  The single block also advances the epoch.
__global__ void __launch_bounds__(64) k_halo_handshake(int my_rank, int world, const int64_t* __restrict__ peer_flags, unsigned long long peers,
                                                       int* state) {
  const int epoch = *reinterpret_cast<volatile int*>(state) + 1;
  for (int r = threadIdx.x; r < world; r += blockDim.x)
    if (r != my_rank) st_release_sys(reinterpret_cast<int*>(static_cast<uintptr_t>(peer_flags[r])) + my_rank, epoch);
  for (int r = threadIdx.x; r < world; r += blockDim.x)
    if ((peers >> r) & 1ull) {
      const int* mine = reinterpret_cast<const int*>(static_cast<uintptr_t>(peer_flags[my_rank])) + r;
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
  if (threadIdx.x == 0) {
    trace_ns(state, 0);
    *reinterpret_cast<volatile int*>(state) = epoch;
  }
}

// an exchange without links still has to announce, advance the epoch and raise the gate
__global__ void k_halo_exchange_empty(int my_rank, int world, const int64_t* __restrict__ peer_flags, int* state, int nb, int gated) {
  const int epoch = *reinterpret_cast<volatile int*>(state) + 1;
  if ((int)threadIdx.x < world && (int)threadIdx.x != my_rank) {
    st_release_sys(reinterpret_cast<int*>(static_cast<uintptr_t>(peer_flags[threadIdx.x])) + my_rank, epoch);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile int*>(state) = epoch;
    if (gated)
      for (int b = 0; b < nb; ++b) st_release_gpu(state + kGateWord + b, 1);
  }
}

// see impl.cuh: the kernels of the exchange itself, loaded before the first one can spin
int halo_kernels_preload() {
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, k_halo_exchange<double>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_halo_exchange<float>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_halo_exchange2<double>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_halo_exchange2<float>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_halo_exchange3<double>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_halo_exchange3<float>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_halo_handshake);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_halo_pull<double, 1>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_halo_pull<float, 1>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_halo_exchange_empty);
  if (e != cudaSuccess) return set_error((int)e, "halo exchange kernels: %s", cudaGetErrorString(e));
  return B2S_OK;
}

int halo_exchange_launch(int elem_size, int nb, int max_strip, const HaloXchg& X, bool narrow, cudaStream_t s) {
  B2S_ARGCHECK(X.world >= 1 && X.world <= 64 && X.my_rank >= 0 && X.my_rank < X.world, "halo_exchange: rank %d of %d", X.my_rank, X.world);
  B2S_ARGCHECK(X.peer_flags && X.state, "halo_exchange: null pointer");
  if (X.nlinks == 0) {
    k_halo_exchange_empty<<<1, 64, 0, s>>>(X.my_rank, X.world, X.peer_flags, X.state, nb, X.gated);
    return check_launch("halo_exchange");
  }
  B2S_ARGCHECK(X.nk > 0 && X.links && X.dst && X.b_total, "halo_exchange: bad sizes nlinks=%d nk=%d", X.nlinks, X.nk);
  const int64_t units1 = (int64_t)X.nlinks * X.nk;  // (link, level) strips
  int per_sm = option("halo_blocks_per_sm", 0);
  // beside a gated stencil: version 1 -- measured on two GPUs (C384x72, round 2) the forked version 2 made the overlapped
  // step three times longer (0.9 ms against 0.31 ms); its 64 registers per thread do not fit twice per SM beside three
  // 160-thread stencil CTAs
  // auto: alone on the stream -> version 4, the one-block handshake kernel followed by the flat-grid pull (on 8 GPUs
  // 93.5 us per step against 96.4 / 98.3 us for the one-kernel versions 1 / 2, and against 113 - 115 us before the handshake
  // wait was relayed by block 0: profiles/README.md, round 2); beside a gated stencil -> version 1 (below)
  const int variant = option("halo_variant", 0);
  if ((variant == 4 || variant == 0) && !X.gated) {
    k_halo_handshake<<<1, 64, 0, s>>>(X.my_rank, X.world, X.peer_flags, X.peers, X.state);
    if (int rc = check_launch("halo_exchange(handshake)")) return rc;
    const int ku = halo_levels();
    dim3 grid((max_strip + 255) / 256, (X.nk + ku - 1) / ku, X.nlinks);
    if (elem_size == 8) {
      if (ku == 8) k_halo_pull<double, 8><<<grid, 256, 0, s>>>(X.nk, X.links, static_cast<double*>(X.dst), kExchangeWords);
      else if (ku == 4) k_halo_pull<double, 4><<<grid, 256, 0, s>>>(X.nk, X.links, static_cast<double*>(X.dst), kExchangeWords);
      else k_halo_pull<double, 1><<<grid, 256, 0, s>>>(X.nk, X.links, static_cast<double*>(X.dst), kExchangeWords);
    } else {
      if (ku == 8) k_halo_pull<float, 8><<<grid, 256, 0, s>>>(X.nk, X.links, static_cast<float*>(X.dst), kExchangeWords);
      else if (ku == 4) k_halo_pull<float, 4><<<grid, 256, 0, s>>>(X.nk, X.links, static_cast<float*>(X.dst), kExchangeWords);
      else k_halo_pull<float, 1><<<grid, 256, 0, s>>>(X.nk, X.links, static_cast<float*>(X.dst), kExchangeWords);
    }
    return check_launch("halo_exchange(pull)");
  }
  if (variant == 1 || (variant == 0 && X.gated) || !narrow) {
    // beside a gated stencil (forked exchange): 2 blocks x 256 threads x ~50 registers per SM leave room for four stencil
    // CTAs; alone on the GPU: every thread slot, the copy is bound by the number of (remote) loads in flight
    if (per_sm <= 0 || per_sm > 8) per_sm = X.gated ? 2 : 8;
    const int grid = (int)(units1 < (int64_t)sm_count() * per_sm ? units1 : (int64_t)sm_count() * per_sm);
    if (elem_size == 8)
      k_halo_exchange<double><<<grid, 256, 0, s>>>(X);
    else
      k_halo_exchange<float><<<grid, 256, 0, s>>>(X);
    return check_launch("halo_exchange");
  }
  // Version 2: a persistent grid walks the (link, chunk of ku levels) units in link order, so sub-domains complete -- and
  // their gates open -- one after the other.  ku = 1 level per unit: filling a unit up to ONE batch of eight loads per
  // thread (3 levels of a 3 x 192 strip) measured 2 % slower per step on 8 GPUs than one level per unit with three times
  // the units (profiles/README.md, round 2); b2s_set_option("halo_levels_per_unit", n) for A/B runs.  Blocks per SM: 2
  // beside a gated stencil (thousands of blocks would take every thread slot and keep the stencil from becoming
  // resident, profiles/r02_overlap.md), 4 (the register limit) alone on the stream.
  int ku = option("halo_levels_per_unit", 0);
  if (ku <= 0 || ku > kMaxLevelsPerUnit) ku = 1;
  if (ku > X.nk) ku = X.nk;
  if (per_sm <= 0 || per_sm > 4) per_sm = X.gated ? 2 : 4;
  const int64_t slots = (int64_t)sm_count() * per_sm;
  const int64_t units = (int64_t)X.nlinks * ((X.nk + ku - 1) / ku);
  const int grid = (int)(units < slots ? units : slots);
  if (elem_size == 8)
    k_halo_exchange2<double><<<grid, 256, 0, s>>>(X, ku);
  else
    k_halo_exchange2<float><<<grid, 256, 0, s>>>(X, ku);
  return check_launch("halo_exchange");
}

template int halo_pull<double>(int, int, int, const int64_t*, double*, cudaStream_t);
template int halo_pull<float>(int, int, int, const int64_t*, float*, cudaStream_t);

template int halo_move<double>(int, int, int, const int64_t*, const double*, double*, cudaStream_t);
template int halo_move<float>(int, int, int, const int64_t*, const float*, float*, cudaStream_t);

}  // namespace impl
}  // namespace b2s
