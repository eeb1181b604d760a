// K6b (slab variant): conservative remap with the SOURCE column block staged in shared memory.
// Spec: oracle/numpy_oracle.py remap_column (no source in /root/reference).
//
// Why: the thread-per-column kernel (k_vertical.cu k_remap_nested) is a chain of ~2 nk dependent
// HBM loads per column (every advance of the source pointer waits for the edge it just loaded);
// it runs at 60 % (fp64) / 47 % (fp32) of the HBM roofline on occupancy alone.  Here the dependent
// chain runs against shared memory and HBM only sees bulk, independent, coalesced traffic:
//   * CTA = one block of 32 adjacent columns of one (j, b) row; the block's source edges pe1
//     (nk1+1 levels) and values q1 (nk1 levels) are brought into shared memory as two slabs
//     [level][32 columns] -- by TWO TMA tile loads issued by one thread (cp.async.bulk.tensor.4d,
//     box 32 x 1 x levels x 1, completing on an mbarrier), or by per-element cp.async (LDGSTS) when a
//     field does not meet the TMA alignment rules.  2-6 CTAs per SM keep 70-200 KB of loads in
//     flight per SM with no registers holding them;
//   * the nk2 target levels are split into contiguous chunks, one per warp (lane = column, so
//     pe2 loads and q2 stores are coalesced 128/256-byte rows).  A chunk's target edges are loaded
//     into registers BEFORE the wait on the slab, so their latency hides behind the slab load;
//   * a warp finds its starting source layer with a binary search in the slab (first layer whose
//     lower edge lies below the chunk's first target edge), then marches exactly like the oracle:
//     same overlaps, same order of accumulation -> bit-identical results for monotone edges (the
//     precondition the specification states; for non-monotone edges the marching pointer of the
//     thread-per-column kernel and the search differ, remap_variant=1 keeps the former);
//   * lane = column means slab address = level*32 + lane: every lane hits its own bank whatever
//     level it is at -- no bank conflicts in the data-dependent reads.
// remap_delp (pe_prefix fused): the delp slab is loaded one row down and warp 0 turns it into pe1
// in place (sequential in k, the oracle's order of additions) before the other warps start.
// Algorithmic bytes/point: 24 R + 8 W, each element crossing HBM exactly once.
#include "impl.cuh"
#include "tma.cuh"

namespace b2s {
namespace impl {

namespace {

constexpr int kCols = 32;   // columns per CTA (lane = column)
constexpr int kWarps = 8;   // warps per CTA (one chunk of target levels each)
constexpr int kThreads = kCols * kWarps;

template <typename T>
struct SlabParams {
  int ni, nj, nk1, nk2, ntile_i;
  T ptop;
  F3<const T> e1;  // pe1 (nk1+1 levels) or delp (nk1 levels)
  F3<const T> q1, pe2;
  F3<T> q2;
};

// LOADER: 0 = cp.async per element, 1 = TMA tile loads
template <typename T, int CH, bool DELP, int LOADER, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_remap_slab(const __grid_constant__ CUtensorMap tm_e,
                                                               const __grid_constant__ CUtensorMap tm_q,
                                                               const SlabParams<T> P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nk1 = P.nk1, nk2 = P.nk2;
  T* E = reinterpret_cast<T*>(smem);            // [nk1+1][32] source edges
  T* Q = E + (size_t)(nk1 + 1) * kCols;         // [nk1][32] source values
  uint64_t* bar = reinterpret_cast<uint64_t*>(Q + (size_t)nk1 * kCols);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int ti = t % P.ntile_i;
  t /= P.ntile_i;
  const int j = t % P.nj;
  const int b = t / P.nj;
  const int i = ti * kCols + lane;
  const bool valid = i < P.ni;
  constexpr int E0 = DELP ? 1 : 0;              // first slab row the load fills
  const int ke = DELP ? nk1 : nk1 + 1;          // levels of the e1 field

  // ---- 1. start the slab loads ----
  if (LOADER == 1) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      fence_barrier_init();
      mbar_arrive_expect_tx(bar, (uint32_t)((ke + nk1) * kCols * sizeof(T)));
      tma_load_4d(E + E0 * kCols, &tm_e, bar, ti * kCols, j, 0, b);
      tma_load_4d(Q, &tm_q, bar, ti * kCols, j, 0, b);
    }
  } else {
    if (valid) {
      const T* ep = P.e1.at(i, j, 0, b);
      const T* qp = P.q1.at(i, j, 0, b);
      for (int k = warp; k < ke; k += kWarps) cp_async<sizeof(T)>(E + (k + E0) * kCols + lane, ep + (int64_t)k * P.e1.sk);
      for (int k = warp; k < nk1; k += kWarps) cp_async<sizeof(T)>(Q + k * kCols + lane, qp + (int64_t)k * P.q1.sk);
    }
    cp_async_commit();
  }

  // ---- 2. target edges of this warp's first chunk, in flight while the slab arrives ----
  const T* e2 = P.pe2.at(valid ? i : 0, j, 0, b);
  T* o2 = P.q2.at(valid ? i : 0, j, 0, b);
  int k2b = warp * CH;
  T tgt[CH + 1];
#pragma unroll
  for (int u = 0; u <= CH; ++u) tgt[u] = (valid && k2b + u <= nk2) ? __ldg(e2 + (int64_t)(k2b + u) * P.pe2.sk) : T(0);

  // ---- 3. slab complete ----
  if (LOADER == 1) {
    __syncthreads();  // the barrier word is initialised
    mbar_wait(bar, 0);
  } else {
    cp_async_wait_all();
    __syncthreads();
  }
  if (DELP) {
    // pe1[0] = ptop; pe1[k+1] = pe1[k] + delp[k]: warp 0, lane = column, 8 levels loaded ahead
    if (warp == 0) {
      T acc = P.ptop;
      E[lane] = acc;
      for (int kb = 0; kb < nk1; kb += 8) {
        T d[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) d[u] = kb + u < nk1 ? E[(kb + u + 1) * kCols + lane] : T(0);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (kb + u < nk1) {
            acc = acc + d[u];
            E[(kb + u + 1) * kCols + lane] = acc;
          }
      }
    }
    __syncthreads();
  }
  if (!valid) return;

  // ---- 4. chunks of target levels ----
  const T* El = E + lane;
  const T* Ql = Q + lane;
  for (;;) {
    if (k2b >= nk2) break;
    T lo = tgt[0];
    // first source layer whose lower edge lies below lo (or the last layer): where the oracle's
    // marching pointer stands when it reaches target level k2b
    int k1;
    {
      int a = 0, z = nk1 - 1;
      while (a < z) {
        const int m = (a + z) >> 1;
        if (El[(m + 1) * kCols] > lo) z = m; else a = m + 1;
      }
      k1 = a;
    }
    T top = El[k1 * kCols], bot = El[(k1 + 1) * kCols], qv = Ql[k1 * kCols];
#pragma unroll
    for (int u = 0; u < CH; ++u) {
      if (k2b + u < nk2) {
        const T hi = tgt[u + 1];
        while (k1 < nk1 - 1 && bot <= lo) {
          ++k1;
          top = bot;
          bot = El[(k1 + 1) * kCols];
          qv = Ql[k1 * kCols];
        }
        T acc = T(0);
        for (;;) {
          const T a = lo > top ? lo : top;
          const T c = hi < bot ? hi : bot;
          if (c > a) acc = acc + (c - a) * qv;
          if (bot >= hi || k1 == nk1 - 1) break;
          ++k1;
          top = bot;
          bot = El[(k1 + 1) * kCols];
          qv = Ql[k1 * kCols];
        }
        __stcs(o2 + (int64_t)(k2b + u) * P.q2.sk, acc / (hi - lo));
        lo = hi;
      }
    }
    k2b += kWarps * CH;
    if (k2b >= nk2) break;
#pragma unroll
    for (int u = 0; u <= CH; ++u) tgt[u] = (k2b + u <= nk2) ? __ldg(e2 + (int64_t)(k2b + u) * P.pe2.sk) : T(0);
  }
}

template <typename T>
size_t slab_bytes(int nk1) {
  return (size_t)(2 * nk1 + 1) * kCols * sizeof(T) + 16;
}

template <typename T, int CH, bool DELP, int LOADER, int MINB>
int launch_slab(const CUtensorMap& me, const CUtensorMap& mq, const SlabParams<T>& P, int64_t grid, cudaStream_t s,
                const char* what) {
  auto kern = k_remap_slab<T, CH, DELP, LOADER, MINB>;
  const size_t smem = slab_bytes<T>(P.nk1);
  static size_t configured = 0;  // largest dynamic shared size this instance was opted in for
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return set_error((int)e, "%s(slab): cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    configured = smem;
  }
  kern<<<(unsigned)grid, kThreads, smem, s>>>(me, mq, P);
  return check_launch(what);
}

}  // namespace

// variant: 2 = cp.async loader, 3 = TMA loader.  *applicable = false (and B2S_OK) when this shape
// or alignment is outside what the slab kernel covers; the caller then uses the nested kernel.
template <typename T, bool DELP>
int remap_slab(int variant, int ni, int nj, int nk1, int nk2, int nb, T ptop, F3<const T> e1, F3<const T> q1,
               F3<const T> pe2, F3<T> q2, cudaStream_t s, bool* applicable) {
  *applicable = false;
  const char* what = DELP ? "remap_delp" : "remap";
  // at least two CTAs per SM must fit, or the loads of one CTA cannot overlap the march of another
  if (2 * (slab_bytes<T>(nk1) + 1024) > (size_t)227 * 1024) return B2S_OK;
  const int ntile_i = (ni + kCols - 1) / kCols;
  const int64_t grid = (int64_t)ntile_i * nj * nb;
  if (grid > 0x7fffffffLL) return B2S_OK;
  SlabParams<T> P;
  P.ni = ni, P.nj = nj, P.nk1 = nk1, P.nk2 = nk2, P.ntile_i = ntile_i;
  P.ptop = ptop;
  P.e1 = e1, P.q1 = q1, P.pe2 = pe2, P.q2 = q2;
  const int ke = DELP ? nk1 : nk1 + 1;
  CUtensorMap me, mq;
  bool tma = variant == 3 && ke <= 256;
  if (tma) {
    const TmaField<T> fe = tma_field<T>(e1.p, e1.sj, e1.sk, e1.sb, ke, nb);
    const TmaField<T> fq = tma_field<T>(q1.p, q1.sj, q1.sk, q1.sb, nk1, nb);
    tma = fe.ok && fq.ok && fe.off == 0 && fq.off == 0 &&
          make_map<T>(&me, e1.p, e1.sj, e1.sk, e1.sb, ni, nj, ke, nb, kCols, 1, ke) &&
          make_map<T>(&mq, q1.p, q1.sj, q1.sk, q1.sb, ni, nj, nk1, nb, kCols, 1, nk1);
  }
  if (!tma) {
    if (variant == 3) return B2S_OK;  // TMA was asked for explicitly and does not apply
    memset(&me, 0, sizeof(me));
    memset(&mq, 0, sizeof(mq));
  }
  *applicable = true;
  // chunk length: 8 warps x CH levels cover the column in one round for nk2 <= 80 (CH = 10) or <= 144 (CH = 18)
  const bool small = nk2 <= kWarps * 10;
  if (tma) {
    return small ? launch_slab<T, 10, DELP, 1, 3>(me, mq, P, grid, s, what)
                 : launch_slab<T, 18, DELP, 1, 2>(me, mq, P, grid, s, what);
  }
  return small ? launch_slab<T, 10, DELP, 0, 3>(me, mq, P, grid, s, what)
               : launch_slab<T, 18, DELP, 0, 2>(me, mq, P, grid, s, what);
}

#define INSTANTIATE(T)                                                                                            \
  template int remap_slab<T, false>(int, int, int, int, int, int, T, F3<const T>, F3<const T>, F3<const T>, F3<T>, \
                                    cudaStream_t, bool*);                                                        \
  template int remap_slab<T, true>(int, int, int, int, int, int, T, F3<const T>, F3<const T>, F3<const T>, F3<T>,  \
                                   cudaStream_t, bool*);
INSTANTIATE(double)
INSTANTIATE(float)

}  // namespace impl
}  // namespace b2s
