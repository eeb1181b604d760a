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

template <typename T>
struct SlabParams {
  int ni, nj, nk1, nk2, ntile_i;
  int off_e, off_q;  // TMA loader: elements between the 16-byte aligned tensor base and compute column 0
  T ptop;
  F3<const T> e1;  // pe1 (nk1+1 levels) or delp (nk1 levels)
  F3<const T> q1, pe2;
  F3<T> q2;
};

// Geometry of a CTA: CG column groups of 32 columns (slab rows of 32*CG elements: 256 B for fp64 with
// CG = 1 and for fp32 with CG = 2), NW warps per column group, each marching a chunk of CH target levels.
template <int NW, int CG>
struct SlabGeom {
  static constexpr int COLS = 32 * CG;
  static constexpr int THREADS = 32 * NW * CG;
};

// Shared-memory layout: [pad][E slab: nk1+1 rows][Q slab: nk1 rows][2 mbarriers].  TMA destinations must
// be 128-byte aligned: the pad puts the first row the E load fills (row 1 for remap_delp) on a 128-byte
// boundary, the Q slab starts on the next one.  rowb = slab row pitch in bytes.
struct SlabLayout {
  int e_off, q_off, bar_off, total;
  __host__ __device__ SlabLayout(int nk1, int rowb, bool delp) {
    e_off = delp ? (128 - rowb % 128) % 128 : 0;
    q_off = (e_off + (nk1 + 1) * rowb + 127) / 128 * 128;
    bar_off = (q_off + nk1 * rowb + 15) / 16 * 16;
    total = bar_off + 16;
  }
};

// LOADER: 0 = cp.async per element, 1 = TMA tile loads, 2 = TMA tile loads of fields that start off a
// 16-byte boundary (interior windows of halo-padded storage): a TMA box must start 16-byte aligned in
// global memory, so the box starts at the aligned column at or before the block's first one, the slab
// rows are 16 bytes wider and every thread skips `off` leading elements.
template <typename T, int CH, int NW, int CG, bool DELP, int LOADER, int MINB>
__global__ void __launch_bounds__(32 * NW * CG, MINB) k_remap_slab(const __grid_constant__ CUtensorMap tm_e,
                                                                   const __grid_constant__ CUtensorMap tm_q,
                                                                   const SlabParams<T> P) {
  constexpr int BCOLS = SlabGeom<NW, CG>::COLS;                            // columns of the block
  constexpr int COLS = BCOLS + (LOADER == 2 ? 16 / (int)sizeof(T) : 0);  // slab row pitch in elements
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nk1 = P.nk1, nk2 = P.nk2;
  const SlabLayout L(nk1, COLS * (int)sizeof(T), DELP);
  T* E = reinterpret_cast<T*>(smem + L.e_off);              // [nk1+1][COLS] source edges
  T* Q = reinterpret_cast<T*>(smem + L.q_off);              // [nk1][COLS] source values
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);  // [0]: E slab, [1]: Q slab

  const int warp = (threadIdx.x >> 5) % NW;   // chunk index inside the column group
  const int col = (threadIdx.x >> 5) / NW * 32 + (threadIdx.x & 31);  // column of the block owned by this thread
  int t = blockIdx.x;
  const int ti = t % P.ntile_i;
  t /= P.ntile_i;
  const int j = t % P.nj;
  const int b = t / P.nj;
  const int i = ti * BCOLS + col;
  const bool valid = i < P.ni;
  constexpr int E0 = DELP ? 1 : 0;            // first slab row the load fills
  const int ke = DELP ? nk1 : nk1 + 1;        // levels of the e1 field

  // ---- 1. start the slab loads ----
  if (LOADER != 0) {
    if (threadIdx.x == 0) {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      fence_barrier_init();
      mbar_arrive_expect_tx(&bar[0], (uint32_t)(ke * COLS * sizeof(T)));
      tma_load_4d(E + E0 * COLS, &tm_e, &bar[0], ti * BCOLS, j, 0, b);
      mbar_arrive_expect_tx(&bar[1], (uint32_t)(nk1 * COLS * sizeof(T)));
      tma_load_4d(Q, &tm_q, &bar[1], ti * BCOLS, j, 0, b);
    }
  } else {
    if (valid) {
      const T* ep = P.e1.at(i, j, 0, b);
      const T* qp = P.q1.at(i, j, 0, b);
      for (int k = warp; k < ke; k += NW) cp_async<sizeof(T)>(E + (k + E0) * COLS + col, ep + (int64_t)k * P.e1.sk);
      for (int k = warp; k < nk1; k += NW) cp_async<sizeof(T)>(Q + k * COLS + col, qp + (int64_t)k * P.q1.sk);
    }
    cp_async_commit();
  }

  // ---- 2. target edges of this warp's first chunk, in flight while the slab arrives ----
  // (running pointers: one 64-bit add per level instead of a 64-bit multiply-add per access)
  int k2b = warp * CH;
  const int64_t sk2 = P.pe2.sk, sko = P.q2.sk;
  const T* e2 = P.pe2.at(valid ? i : 0, j, k2b, b);
  T* o2 = P.q2.at(valid ? i : 0, j, k2b, b);
  int nlev = valid ? min(CH, nk2 - k2b) : -1;  // target levels of this chunk (<= 0: none; < 0: no edge is loaded either)
  T tgt[CH + 1];
  {
    const T* p = e2;
#pragma unroll
    for (int u = 0; u <= CH; ++u) {
      tgt[u] = u <= nlev ? __ldg(p) : T(0);
      p += sk2;
    }
  }

  // ---- 3. slab complete ----
  if (LOADER != 0) {
    __syncthreads();  // the barrier words are initialised
    mbar_wait(&bar[0], 0);
  } else {
    cp_async_wait_all();
    __syncthreads();
  }
  if (DELP) {
    // pe1[0] = ptop; pe1[k+1] = pe1[k] + delp[k], in place, sequential in k (the oracle's order of
    // additions): ONE warp per column group, lane = column, the next 8 levels loaded while the current 8
    // are added; the q1 slab is still arriving meanwhile.  The warp that does it rotates with the block
    // index, so the extra instructions spread over the four schedulers of the SM.
    if (warp == (int)(blockIdx.x % NW)) {
      T* Ec = E + col + (LOADER == 2 ? P.off_e : 0);
      T acc = P.ptop;
      Ec[0] = acc;
      Ec += COLS;  // Ec[k * COLS] = delp[k] -> pe1[k + 1]
      T d[8], dn[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) d[u] = u < nk1 ? Ec[u * COLS] : T(0);
      int kb = 0;
      for (; kb + 16 <= nk1; kb += 8) {  // this batch and the next are complete: no predicates
#pragma unroll
        for (int u = 0; u < 8; ++u) dn[u] = Ec[(8 + u) * COLS];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          acc = acc + d[u];
          Ec[u * COLS] = acc;
          d[u] = dn[u];
        }
        Ec += 8 * COLS;
      }
      for (; kb < nk1; kb += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) dn[u] = kb + 8 + u < nk1 ? Ec[(8 + u) * COLS] : T(0);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (kb + u < nk1) {
            acc = acc + d[u];
            Ec[u * COLS] = acc;
          }
          d[u] = dn[u];
        }
        Ec += 8 * COLS;
      }
    }
    __syncthreads();
  }
  if (LOADER != 0) mbar_wait(&bar[1], 0);

  // ---- 4. chunks of target levels ----
  // Per target layer the oracle first skips the source layers that end at or above the layer's upper
  // edge (`while bot <= lo`) and then accumulates overlaps until a source layer reaches the lower edge.
  // A skipped layer has no overlap (min(hi,bot) <= lo), so running it through the accumulation loop adds
  // nothing and advances the pointer just the same: one loop does both, with identical results.
  const T* El = E + col + (LOADER == 2 ? P.off_e : 0);
  const T* Ql = Q + col + (LOADER == 2 ? P.off_q : 0);
  const int klast = nk1 - 1;
  while (nlev > 0) {
    T lo = tgt[0];
    // first source layer whose lower edge lies below lo (or the last layer): where the oracle's
    // marching pointer stands when it reaches target level k2b
    int k1;
    {
      int a = 0, z = klast;
      while (a < z) {
        const int m = (a + z) >> 1;
        if (El[(m + 1) * COLS] > lo) z = m; else a = m + 1;
      }
      k1 = a;
    }
    const T* Ek = El + k1 * COLS;  // Ek[0] = top edge of the layer in hand
    const T* Qk = Ql + k1 * COLS;
    T top = Ek[0], bot = Ek[COLS], qv = Qk[0];
    T* op = o2;
#pragma unroll
    for (int u = 0; u < CH; ++u) {
      if (u < nlev) {
        const T hi = tgt[u + 1];
        T acc = T(0);
        for (;;) {
          const T a = lo > top ? lo : top;
          const T c = hi < bot ? hi : bot;
          if (c > a) acc = acc + (c - a) * qv;
          if (bot >= hi || k1 == klast) break;
          ++k1;
          Ek += COLS;
          Qk += COLS;
          top = bot;
          bot = Ek[COLS];
          qv = Qk[0];
        }
        __stcs(op, acc / (hi - lo));
        op += sko;
        lo = hi;
      }
    }
    k2b += NW * CH;
    nlev = min(CH, nk2 - k2b);
    if (nlev <= 0) break;
    e2 += (int64_t)(NW * CH) * sk2;
    o2 += (int64_t)(NW * CH) * sko;
    {
      const T* p = e2;
#pragma unroll
      for (int u = 0; u <= CH; ++u) {
        tgt[u] = u <= nlev ? __ldg(p) : T(0);
        p += sk2;
      }
    }
  }
}

// cols = slab row pitch in elements (block columns, + 16 bytes for the shifted TMA loader)
size_t slab_bytes(int nk1, int cols, size_t elem, bool delp = true) {
  return (size_t)SlabLayout(nk1, cols * (int)elem, delp).total;
}

template <typename T, int CH, int NW, int CG, bool DELP, int LOADER, int MINB>
int launch_slab(const CUtensorMap& me, const CUtensorMap& mq, const SlabParams<T>& P, int nj, int nb, cudaStream_t s,
                const char* what) {
  using G = SlabGeom<NW, CG>;
  auto kern = k_remap_slab<T, CH, NW, CG, DELP, LOADER, MINB>;
  const size_t smem = slab_bytes(P.nk1, G::COLS + (LOADER == 2 ? 16 / (int)sizeof(T) : 0), sizeof(T), DELP);
  static size_t configured = 0;  // largest dynamic shared size this instance was opted in for
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return set_error((int)e, "%s(slab): cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    configured = smem;
  }
  const int64_t grid = (int64_t)P.ntile_i * nj * nb;
  kern<<<(unsigned)grid, G::THREADS, smem, s>>>(me, mq, P);
  return check_launch(what);
}

}  // namespace

// variant: 2 = cp.async loader, 3 = TMA loader.  *applicable = false (and B2S_OK) when this shape
// or alignment is outside what the slab kernel covers; the caller then uses the nested kernel.
// Geometry (b2s_set_option "remap_nw" = 8 | 16 warps per column group, "remap_cg" = 1 | 2 column
// groups; 0 = automatic, from the sweeps in profiles/): the TMA loader comes in 8x1, 16x1 and 8x2,
// the cp.async loader (the path for fields TMA cannot address) in 8x1.
template <typename T, bool DELP>
int remap_slab(int variant, int ni, int nj, int nk1, int nk2, int nb, T ptop, F3<const T> e1, F3<const T> q1,
               F3<const T> pe2, F3<T> q2, cudaStream_t s, bool* applicable) {
  *applicable = false;
  const char* what = DELP ? "remap_delp" : "remap";
  int nw = option("remap_nw", 0), cg = option("remap_cg", 0);
  if (variant != 3) nw = 8, cg = 1;
  if (cg == 0) cg = 1;  // 32-column CTAs win for both precisions (profiles/r01_remap_geometry.json)
  if (cg == 2 && ni <= 32) cg = 1;
  if (nw == 0) nw = 8;
  if (cg == 2) nw = 8;
  constexpr int V = 16 / (int)sizeof(T);
  // at least two CTAs per SM must fit, or the loads of one CTA cannot overlap the march of another
  if (2 * (slab_bytes(nk1, 32 * cg + V, sizeof(T)) + 1024) > (size_t)227 * 1024) {
    if (cg == 2 && 2 * (slab_bytes(nk1, 32 + V, sizeof(T)) + 1024) <= (size_t)227 * 1024) cg = 1;
    else return B2S_OK;
  }
  const int ntile_i = (ni + 32 * cg - 1) / (32 * cg);
  if ((int64_t)ntile_i * nj * nb > 0x7fffffffLL) return B2S_OK;
  SlabParams<T> P;
  P.ni = ni, P.nj = nj, P.nk1 = nk1, P.nk2 = nk2, P.ntile_i = ntile_i;
  P.ptop = ptop;
  P.off_e = P.off_q = 0;
  P.e1 = e1, P.q1 = q1, P.pe2 = pe2, P.q2 = q2;
  const int ke = DELP ? nk1 : nk1 + 1;
  CUtensorMap me, mq;
  bool tma = variant == 3 && ke <= 256;
  bool shifted = false;
  if (tma) {
    const TmaField<T> fe = tma_field<T>(e1.p, e1.sj, e1.sk, e1.sb, ke, nb);
    const TmaField<T> fq = tma_field<T>(q1.p, q1.sj, q1.sk, q1.sb, nk1, nb);
    shifted = fe.off != 0 || fq.off != 0;
    if (shifted) cg = 1, nw = 8;  // the shifted loader comes in one geometry
    const int box = 32 * cg + (shifted ? V : 0);
    P.off_e = fe.off, P.off_q = fq.off;
    P.ntile_i = (ni + 32 * cg - 1) / (32 * cg);
    tma = fe.ok && fq.ok &&
          make_map<T>(&me, fe.base, e1.sj, e1.sk, e1.sb, ni + fe.off, nj, ke, nb, box, 1, ke) &&
          make_map<T>(&mq, fq.base, q1.sj, q1.sk, q1.sb, ni + fq.off, nj, nk1, nb, box, 1, nk1);
  }
  if (!tma) {
    if (variant == 3) return B2S_OK;  // TMA was asked for explicitly and does not apply
    memset(&me, 0, sizeof(me));
    memset(&mq, 0, sizeof(mq));
  }
  *applicable = true;
  // chunk length: NW warps x CH levels cover the column in one round for nk2 <= 9 NW (CH = 9), else CH = 18
  // (8 warps: nk2 <= 144 in one round; 16 warps always march 9 levels per round)
  const bool small = nk2 <= nw * 9 || nw == 16;
#define B2S_SLAB(CH, NW, CG, LOADER, MINB) launch_slab<T, CH, NW, CG, DELP, LOADER, MINB>(me, mq, P, nj, nb, s, what)
  if (!tma) return small ? B2S_SLAB(9, 8, 1, 0, 4) : B2S_SLAB(18, 8, 1, 0, 3);
  if (shifted) return small ? B2S_SLAB(9, 8, 1, 2, 3) : B2S_SLAB(18, 8, 1, 2, 3);
  if (cg == 2) return small ? B2S_SLAB(9, 8, 2, 1, 2) : B2S_SLAB(18, 8, 2, 1, sizeof(T) == 4 ? 2 : 1);
  if (nw == 16) return B2S_SLAB(9, 16, 1, 1, 2);
  return small ? B2S_SLAB(9, 8, 1, 1, 4) : B2S_SLAB(18, 8, 1, 1, 3);
#undef B2S_SLAB
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
