// K6a-c: vertical column scans of the dycore.  No source in /root/reference;
// the specification is oracle/numpy_oracle.py (pe_prefix, remap_column, tridiag); field names
// from src/tcn/py_ftn_interface/example_def_dycore.yaml:52-58, the tridiagonal solve is named at
// geos_documentation/moist/GF.drawio:502.
//
// K is never split (SURVEY.md 5 "Long-context": columns are the unit of work): one thread per
// column (W adjacent columns where alignment allows), the carry in registers, levels loaded
// U ahead of use.  Sequential-in-k arithmetic in the oracle's order -> bitwise reproducible.
#include "impl.cuh"
#include "vec.cuh"

namespace b2s {
namespace impl {

static constexpr int kBlock = 128;

// -------------------------------------------------------------------------------------------
// K6a pe_prefix: pe[0] = ptop; pe[k+1] = pe[k] + delp[k].  Bytes/point: 8 R + 8 W.
// -------------------------------------------------------------------------------------------
template <typename T, int W, int U>
__global__ void __launch_bounds__(kBlock) k_pe_prefix(int niw, int nj, int nk, int ncols, T ptop, F3<const T> delp,
                                                      F3<T> pe) {
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (c >= ncols) return;
  const Col cc = decompose_column(c, niw, nj);
  const int i = cc.i * W;
  const T* dp = delp.at(i, cc.j, 0, cc.b);
  T* pp = pe.at(i, cc.j, 0, cc.b);
  Vec<T, W> acc;
#pragma unroll
  for (int w = 0; w < W; ++w) acc.v[w] = ptop;
  VecIO<T, W>::st(pp, acc);
  for (int kb = 0; kb < nk; kb += U) {
    Vec<T, W> x[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (kb + u < nk) x[u] = VecIO<T, W>::ld(dp + (int64_t)(kb + u) * delp.sk);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (kb + u < nk) {
#pragma unroll
        for (int w = 0; w < W; ++w) acc.v[w] = acc.v[w] + x[u].v[w];
        VecIO<T, W>::st(pp + (int64_t)(kb + u + 1) * pe.sk, acc);
      }
  }
}

template <typename T>
int pe_prefix(int ni, int nj, int nk, int nb, T ptop, F3<const T> delp, F3<T> pe, cudaStream_t s) {
  B2S_ARGCHECK(ni > 0 && nj > 0 && nk > 0 && nb > 0, "pe_prefix: empty domain %dx%dx%dx%d", ni, nj, nk, nb);
  B2S_ARGCHECK(delp.p && pe.p, "pe_prefix: null field");
  constexpr int WMAX = MaxWidth<T>::value;
  const bool wide = ni % WMAX == 0 && WidthProbe(WMAX, sizeof(T)).field(delp).field(pe).ok &&
                    (int64_t)(ni / WMAX) * nj * nb >= (int64_t)sm_count() * 1024;
  if (wide) {
    const int ncols = (ni / WMAX) * nj * nb;
    k_pe_prefix<T, WMAX, 8><<<(ncols + kBlock - 1) / kBlock, kBlock, 0, s>>>(ni / WMAX, nj, nk, ncols, ptop, delp, pe);
  } else {
    const int ncols = ni * nj * nb;
    k_pe_prefix<T, 1, 12><<<(ncols + kBlock - 1) / kBlock, kBlock, 0, s>>>(ni, nj, nk, ncols, ptop, delp, pe);
  }
  return check_launch("pe_prefix");
}

// -------------------------------------------------------------------------------------------
// K6b remap: conservative piecewise-constant remap of q1 (source edges pe1, nk1 layers) onto
// target edges pe2 (nk2 layers).  FORWARD over target layers, a source pointer marching
// monotonically (while-loop + variable-K reads: all three dsl_patterns features at once).
// The source layer in hand (top, bot, q) stays in registers and is refilled only when the
// pointer advances, so every element of pe1/q1/pe2 is requested once.
// Bytes/point: 24 R + 8 W.
// -------------------------------------------------------------------------------------------
// Latency-bound by nature: every load depends on the comparison of the previous one (ncu source page:
// 39 % of warp samples wait for the source edge the preceding advance loaded, 30 % for the target edge
// loaded at the top of the level).  What hides that latency here is plain occupancy: ~45 registers,
// 10-11 CTAs (40+ warps) per SM.  Every restructuring tried on the device LOST to this kernel
// (3.5-3.8 ms on C720x137 fp64; profiles/README.md "remap"): a merged-edge uniform loop (4.7 ms),
// register look-ahead queues of depth 1/2/4/8 (3.5-6.0 ms), per-thread shared-memory rings refilled at
// lane-specific times (5.2-16 ms), rings refilled at uniform points per thread (5.4-6.8 ms) or anchored
// at the slowest lane of the warp (4.8-5.2 ms): each adds instructions and registers (fewer resident
// warps) faster than it removes stalls.
template <typename T, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) k_remap_nested(int ni, int nj, int nk1, int nk2, int ncols,
                                                         F3<const T> pe1, F3<const T> q1, F3<const T> pe2,
                                                         F3<T> q2) {
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (c >= ncols) return;
  const Col cc = decompose_column(c, ni, nj);
  // (running pointers bumped per advance instead of the index arithmetic below measured 9 % SLOWER on
  //  C720x137 -- 3.81 ms vs 3.49 ms, three runs each -- so the indices stay)
  const T* e1 = pe1.at(cc.i, cc.j, 0, cc.b);
  const T* s1 = q1.at(cc.i, cc.j, 0, cc.b);
  const T* e2 = pe2.at(cc.i, cc.j, 0, cc.b);
  T* o2 = q2.at(cc.i, cc.j, 0, cc.b);
  int k1 = 0;
  T top = __ldg(e1), bot = __ldg(e1 + pe1.sk), qv = __ldg(s1);
  T lo = __ldg(e2);
  for (int k2 = 0; k2 < nk2; ++k2) {
    const T hi = __ldg(e2 + (int64_t)(k2 + 1) * pe2.sk);
    while (k1 < nk1 - 1 && bot <= lo) {
      ++k1;
      top = bot;
      bot = __ldg(e1 + (int64_t)(k1 + 1) * pe1.sk);
      qv = __ldg(s1 + (int64_t)k1 * q1.sk);
    }
    T acc = T(0);
    for (;;) {
      const T a = lo > top ? lo : top;
      const T b = hi < bot ? hi : bot;
      if (b > a) acc = acc + (b - a) * qv;
      if (bot >= hi || k1 == nk1 - 1) break;
      ++k1;
      top = bot;
      bot = __ldg(e1 + (int64_t)(k1 + 1) * pe1.sk);
      qv = __ldg(s1 + (int64_t)k1 * q1.sk);
    }
    __stcs(o2 + (int64_t)k2 * q2.sk, acc / (hi - lo));
    lo = hi;
  }
}

// slab variant (k_remap_slab.cu): source column block staged in shared memory
template <typename T, bool DELP>
int remap_slab(int variant, int ni, int nj, int nk1, int nk2, int nb, T ptop, F3<const T> e1, F3<const T> q1,
               F3<const T> pe2, F3<T> q2, cudaStream_t s, bool* applicable);

// b2s_set_option("remap_variant", v): 0 auto (slab with TMA loads where the fields allow it, else slab
// with cp.async loads, else nested), 1 nested (thread per column), 2 slab + cp.async, 3 slab + TMA.
template <typename T, bool DELP>
int remap_try_slab(const char* what, int ni, int nj, int nk1, int nk2, int nb, T ptop, F3<const T> e1, F3<const T> q1,
                   F3<const T> pe2, F3<T> q2, cudaStream_t s, bool* done) {
  *done = false;
  const int variant = option("remap_variant", 0);
  if (variant == 1) return B2S_OK;
  int rc = B2S_OK;
  if (variant == 0 || variant == 3) {
    rc = remap_slab<T, DELP>(3, ni, nj, nk1, nk2, nb, ptop, e1, q1, pe2, q2, s, done);
    if (*done) return rc;
    if (variant == 3)
      return set_error(B2S_EUNSUPPORTED, "%s: remap_variant=3 forced but the fields do not meet the TMA rules", what);
  }
  rc = remap_slab<T, DELP>(2, ni, nj, nk1, nk2, nb, ptop, e1, q1, pe2, q2, s, done);
  if (!*done && variant == 2)
    return set_error(B2S_EUNSUPPORTED, "%s: remap_variant=2 forced but %d source levels do not fit shared memory", what, nk1);
  return rc;
}

template <typename T>
int remap(int ni, int nj, int nk1, int nk2, int nb, F3<const T> pe1, F3<const T> q1, F3<const T> pe2, F3<T> q2,
          cudaStream_t s) {
  B2S_ARGCHECK(ni > 0 && nj > 0 && nk1 > 0 && nk2 > 0 && nb > 0, "remap: empty domain %dx%dx(%d->%d)x%d", ni, nj, nk1,
               nk2, nb);
  B2S_ARGCHECK(pe1.p && q1.p && pe2.p && q2.p, "remap: null field");
  {
    bool done = false;
    const int rc = remap_try_slab<T, false>("remap", ni, nj, nk1, nk2, nb, T(0), pe1, q1, pe2, q2, s, &done);
    if (done || rc != B2S_OK) return rc;
  }
  const int ncols = ni * nj * nb;
  const int grid = (ncols + kBlock - 1) / kBlock;
  k_remap_nested<T, 12><<<grid, kBlock, 0, s>>>(ni, nj, nk1, nk2, ncols, pe1, q1, pe2, q2);
  return check_launch("remap");
}

// -------------------------------------------------------------------------------------------
// K6a+b fused, remap_delp: the source edges of the remap ARE the prefix sum of delp
// (pe1[k+1] = pe1[k] + delp[k], pe1[0] = ptop), so the marching pointer builds them as it goes:
// same additions in the same order as pe_prefix, same overlaps as remap -> bit-identical to running
// the two kernels, without writing pe1 (8 B/pt) or reading it back (8 B/pt), and one launch fewer.
// Bytes/point: 24 R (delp, q1, pe2) + 8 W; the unfused pair moves 48.
// -------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kBlock, 12) k_remap_delp(int ni, int nj, int nk1, int nk2, int ncols, T ptop,
                                                           F3<const T> delp, F3<const T> q1, F3<const T> pe2,
                                                           F3<T> q2) {
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (c >= ncols) return;
  const Col cc = decompose_column(c, ni, nj);
  const T* d1 = delp.at(cc.i, cc.j, 0, cc.b);
  const T* s1 = q1.at(cc.i, cc.j, 0, cc.b);
  const T* e2 = pe2.at(cc.i, cc.j, 0, cc.b);
  T* o2 = q2.at(cc.i, cc.j, 0, cc.b);
  int k1 = 0;
  T top = ptop, bot = ptop + __ldg(d1), qv = __ldg(s1);
  T lo = __ldg(e2);
  for (int k2 = 0; k2 < nk2; ++k2) {
    const T hi = __ldg(e2 + (int64_t)(k2 + 1) * pe2.sk);
    while (k1 < nk1 - 1 && bot <= lo) {
      ++k1;
      top = bot;
      bot = top + __ldg(d1 + (int64_t)k1 * delp.sk);
      qv = __ldg(s1 + (int64_t)k1 * q1.sk);
    }
    T acc = T(0);
    for (;;) {
      const T a = lo > top ? lo : top;
      const T b = hi < bot ? hi : bot;
      if (b > a) acc = acc + (b - a) * qv;
      if (bot >= hi || k1 == nk1 - 1) break;
      ++k1;
      top = bot;
      bot = top + __ldg(d1 + (int64_t)k1 * delp.sk);
      qv = __ldg(s1 + (int64_t)k1 * q1.sk);
    }
    __stcs(o2 + (int64_t)k2 * q2.sk, acc / (hi - lo));
    lo = hi;
  }
}

template <typename T>
int remap_delp(int ni, int nj, int nk1, int nk2, int nb, T ptop, F3<const T> delp, F3<const T> q1, F3<const T> pe2,
               F3<T> q2, cudaStream_t s) {
  B2S_ARGCHECK(ni > 0 && nj > 0 && nk1 > 0 && nk2 > 0 && nb > 0, "remap_delp: empty domain %dx%dx(%d->%d)x%d", ni, nj,
               nk1, nk2, nb);
  B2S_ARGCHECK(delp.p && q1.p && pe2.p && q2.p, "remap_delp: null field");
  {
    bool done = false;
    const int rc = remap_try_slab<T, true>("remap_delp", ni, nj, nk1, nk2, nb, ptop, delp, q1, pe2, q2, s, &done);
    if (done || rc != B2S_OK) return rc;
  }
  const int ncols = ni * nj * nb;
  k_remap_delp<T><<<(ncols + kBlock - 1) / kBlock, kBlock, 0, s>>>(ni, nj, nk1, nk2, ncols, ptop, delp, q1, pe2, q2);
  return check_launch("remap_delp");
}

// -------------------------------------------------------------------------------------------
// K6c tridiag (Thomas): FORWARD eliminate  m = b - a w[-1]; w = c/m; x = (d - a x[-1])/m
//                       BACKWARD substitute x = x - w x[+1]
// w (the modified super-diagonal) is spilled to the caller's scratch field, x doubles as the
// forward right-hand side (SURVEY.md 7 "K-carry register pressure").
// Bytes/point: 32 R + 16 W forward, 16 R + 8 W backward = 72.
// -------------------------------------------------------------------------------------------
template <typename T, int U>
__global__ void __launch_bounds__(kBlock) k_tridiag(int ni, int nj, int nk, int ncols, F3<const T> a, F3<const T> b,
                                                    F3<const T> c, F3<const T> d, F3<T> w, F3<T> x) {
  const int col = blockIdx.x * kBlock + threadIdx.x;
  if (col >= ncols) return;
  const Col cc = decompose_column(col, ni, nj);
  const T* ap = a.at(cc.i, cc.j, 0, cc.b);
  const T* bp = b.at(cc.i, cc.j, 0, cc.b);
  const T* cp = c.at(cc.i, cc.j, 0, cc.b);
  const T* dp = d.at(cc.i, cc.j, 0, cc.b);
  T* wp = w.at(cc.i, cc.j, 0, cc.b);
  T* xp = x.at(cc.i, cc.j, 0, cc.b);
  T wprev = T(0), xprev = T(0);
  for (int kb = 0; kb < nk; kb += U) {
    T va[U], vb[U], vc[U], vd[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (kb + u < nk) {
        va[u] = __ldcs(ap + (int64_t)(kb + u) * a.sk);
        vb[u] = __ldcs(bp + (int64_t)(kb + u) * b.sk);
        vc[u] = __ldcs(cp + (int64_t)(kb + u) * c.sk);
        vd[u] = __ldcs(dp + (int64_t)(kb + u) * d.sk);
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (kb + u < nk) {
        const int k = kb + u;
        if (k == 0) {
          wprev = vc[u] / vb[u];
          xprev = vd[u] / vb[u];
        } else {
          const T m = vb[u] - va[u] * wprev;
          xprev = (vd[u] - va[u] * xprev) / m;
          wprev = vc[u] / m;
        }
        wp[(int64_t)k * w.sk] = wprev;
        xp[(int64_t)k * x.sk] = xprev;
      }
  }
  // xprev == x[nk-1]
  for (int kb = nk - 1; kb > 0; kb -= U) {
    T vw[U], vx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = kb - 1 - u;
      if (k >= 0) {
        vw[u] = wp[(int64_t)k * w.sk];
        vx[u] = xp[(int64_t)k * x.sk];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = kb - 1 - u;
      if (k >= 0) {
        xprev = vx[u] - vw[u] * xprev;
        __stcs(xp + (int64_t)k * x.sk, xprev);
      }
    }
  }
}

template <typename T>
int tridiag(int ni, int nj, int nk, int nb, F3<const T> a, F3<const T> b, F3<const T> c, F3<const T> d, F3<T> w,
            F3<T> x, cudaStream_t s) {
  B2S_ARGCHECK(ni > 0 && nj > 0 && nk > 0 && nb > 0, "tridiag: empty domain %dx%dx%dx%d", ni, nj, nk, nb);
  B2S_ARGCHECK(a.p && b.p && c.p && d.p && w.p && x.p, "tridiag: null field");
  const int ncols = ni * nj * nb;
  k_tridiag<T, 4><<<(ncols + kBlock - 1) / kBlock, kBlock, 0, s>>>(ni, nj, nk, ncols, a, b, c, d, w, x);
  return check_launch("tridiag");
}

#define INSTANTIATE(T)                                                                                        \
  template int pe_prefix<T>(int, int, int, int, T, F3<const T>, F3<T>, cudaStream_t);                         \
  template int remap<T>(int, int, int, int, int, F3<const T>, F3<const T>, F3<const T>, F3<T>, cudaStream_t); \
  template int remap_delp<T>(int, int, int, int, int, T, F3<const T>, F3<const T>, F3<const T>, F3<T>, cudaStream_t); \
  template int tridiag<T>(int, int, int, int, F3<const T>, F3<const T>, F3<const T>, F3<const T>, F3<T>, F3<T>, cudaStream_t);
INSTANTIATE(double)
INSTANTIATE(float)

}  // namespace impl
}  // namespace b2s
