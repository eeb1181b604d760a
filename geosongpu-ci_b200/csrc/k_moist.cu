// K4a-c: GEOS moist-physics-style column stencils.  There is no source for
// these in /root/reference; the specification is oracle/numpy_oracle.py (find_klcl,
// saturation_adjust, cloud_top), named after geos_documentation/moist/GFDL_1M.drawio:76,111 and
// the Fortran quoted at dsl_patterns/WIP__hybrid_index_2dout.py:10-15.
//
// Column searches: one thread per column, U levels loaded ahead, the `found` predicate in a
// register; a warp stops reading as soon as __all_sync says every lane is done, so loads stay
// coalesced (no per-lane early return) and the untouched part of the column is never read.
#include "impl.cuh"
#include "vec.cuh"

namespace b2s {
namespace impl {

static constexpr int kBlock = 128;

// -------------------------------------------------------------------------------------------
// K4a find_klcl: BACKWARD, first level (from the surface k = nk-1 upward) with PLmb <= PLCL.
//   KLCL = k (else -1), PLmb_at_KLCL = PLmb[k] (else untouched).
// Algorithmic bytes/point: 8 R + 24/nk (full column; the early exit reads less).
// -------------------------------------------------------------------------------------------
template <typename T, typename I, int U>
__global__ void __launch_bounds__(kBlock) k_find_klcl(int ni, int nj, int nk, int ncols, F3<const T> pl,
                                                      F2<const T> plcl, F2<T> pat, F2<I> klcl) {
  const int c = blockIdx.x * kBlock + threadIdx.x;
  const bool valid = c < ncols;
  const Col cc = decompose_column(valid ? c : 0, ni, nj);
  const T* pp = pl.at(cc.i, cc.j, 0, cc.b);
  const T lim = valid ? __ldg(plcl.at(cc.i, cc.j, cc.b)) : T(0);
  int found = valid ? -1 : 0;
  T pfound = T(0);
  for (int kb = nk; kb > 0; kb -= U) {
    T x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = kb - 1 - u;
      if (k >= 0 && found < 0) x[u] = __ldcs(pp + (int64_t)k * pl.sk);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = kb - 1 - u;
      if (k >= 0 && found < 0 && x[u] <= lim) {
        found = k;
        pfound = x[u];
      }
    }
    if (__all_sync(0xffffffffu, found >= 0)) break;
  }
  if (valid) {
    *klcl.at(cc.i, cc.j, cc.b) = static_cast<I>(found);
    if (found >= 0) *pat.at(cc.i, cc.j, cc.b) = pfound;
  }
}

template <typename T>
int find_klcl(int ni, int nj, int nk, int nb, F3<const T> PLmb, F2<const T> PLCL, F2<T> PLmb_at_KLCL,
              F2<typename IndexOf<T>::type> KLCL, cudaStream_t s) {
  using I = typename IndexOf<T>::type;
  B2S_ARGCHECK(ni > 0 && nj > 0 && nk > 0 && nb > 0, "find_klcl: empty domain %dx%dx%dx%d", ni, nj, nk, nb);
  B2S_ARGCHECK(PLmb.p && PLCL.p && PLmb_at_KLCL.p && KLCL.p, "find_klcl: null field");
  const int ncols = ni * nj * nb;
  k_find_klcl<T, I, 8><<<(ncols + kBlock - 1) / kBlock, kBlock, 0, s>>>(ni, nj, nk, ncols, PLmb, PLCL, PLmb_at_KLCL, KLCL);
  return check_launch("find_klcl");
}

// -------------------------------------------------------------------------------------------
// K4c cloud_top: FORWARD, smallest k with ql > ql_min, -1 for a clear column.
// Algorithmic bytes/point: 8 R + 8/nk.
// -------------------------------------------------------------------------------------------
template <typename T, typename I, int U>
__global__ void __launch_bounds__(kBlock) k_cloud_top(int ni, int nj, int nk, int ncols, T ql_min, F3<const T> ql,
                                                      F2<I> ktop) {
  const int c = blockIdx.x * kBlock + threadIdx.x;
  const bool valid = c < ncols;
  const Col cc = decompose_column(valid ? c : 0, ni, nj);
  const T* qp = ql.at(cc.i, cc.j, 0, cc.b);
  int found = valid ? -1 : 0;
  for (int kb = 0; kb < nk; kb += U) {
    T x[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (kb + u < nk && found < 0) x[u] = __ldcs(qp + (int64_t)(kb + u) * ql.sk);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (kb + u < nk && found < 0 && x[u] > ql_min) found = kb + u;
    if (__all_sync(0xffffffffu, found >= 0)) break;
  }
  if (valid) *ktop.at(cc.i, cc.j, cc.b) = static_cast<I>(found);
}

template <typename T>
int cloud_top(int ni, int nj, int nk, int nb, T ql_min, F3<const T> ql, F2<typename IndexOf<T>::type> ktop,
              cudaStream_t s) {
  using I = typename IndexOf<T>::type;
  B2S_ARGCHECK(ni > 0 && nj > 0 && nk > 0 && nb > 0, "cloud_top: empty domain %dx%dx%dx%d", ni, nj, nk, nb);
  B2S_ARGCHECK(ql.p && ktop.p, "cloud_top: null field");
  const int ncols = ni * nj * nb;
  k_cloud_top<T, I, 8><<<(ncols + kBlock - 1) / kBlock, kBlock, 0, s>>>(ni, nj, nk, ncols, ql_min, ql, ktop);
  return check_launch("cloud_top");
}

// -------------------------------------------------------------------------------------------
// K4b saturation_adjust: PARALLEL pointwise, two fixed Newton steps, in place on T, q, ql.
// Algorithmic bytes/point: 32 R + 24 W = 56.  The reciprocals of (T - 29.65) and (p - (1-eps) es)
// are formed once per step and reused.
// -------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T exp_(T x);
template <>
__device__ __forceinline__ double exp_<double>(double x) {
  return exp(x);
}
template <>
__device__ __forceinline__ float exp_<float>(float x) {
  return expf(x);
}

template <typename T>
__device__ __forceinline__ void sat_adjust_point(T& t, T& qv, T& l, const T pp) {
  const T eps = T(0.622), lcp = T(2.5e6 / 1004.0), one = T(1.0);
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const T rtm = one / (t - T(29.65));
    const T es = T(611.2) * exp_<T>(T(17.67) * (t - T(273.15)) * rtm);
    const T rden = one / (pp - (one - eps) * es);
    const T qs = eps * es * rden;
    const T des = es * T(17.67 * 243.5) * rtm * rtm;
    const T dqs = eps * pp * des * rden * rden;
    T dq = (qv - qs) / (one + lcp * dqs);
    dq = dq > -l ? dq : -l;
    t += lcp * dq;
    qv -= dq;
    l += dq;
  }
}

// fp64: the library exp() and the three IEEE divisions per Newton step cost ~400 issued instructions
// per point (ncu/SASS: 242 DP operations plus ~330 moves that materialise 64-bit literals, plus the
// slow-path checks of every division), which paces the kernel at ~90 Gpts/s = 78 % of the HBM
// roofline: it was ISSUE-bound, not HBM-bound.  This version keeps every constant in constant
// memory (DFMA/DMUL read c[bank][offset] operands directly, no moves), forms reciprocals from
// MUFU.RCP64H + two Newton steps (<= 1 ulp, no slow path: the operands are O(1)..O(1e5) physical
// values) and evaluates exp() as 2^k * P13(r), |r| <= ln2/2 (Taylor remainder 4e-18).  Total error
// vs. the oracle's exp/divide is a few ulp, far inside the 1e-12 the specification asks for.
struct SatConst {
  double eps, one_m_eps, lcp, c2965, c27315, c1767, c6112, cdes;
  double l2e, ln2hi, ln2lo, magic;
  double p[14];  // 1/n!, n = 0..13
};
__constant__ SatConst kSat = {
    0.622, 1.0 - 0.622, 2.5e6 / 1004.0, 29.65, 273.15, 17.67, 611.2, 17.67 * 243.5,
    1.4426950408889634074, 6.93147180369123816490e-01, 1.90821492927058770002e-10, 6755399441055744.0,
    {1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880, 1.0 / 3628800,
     1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0}};

__device__ __forceinline__ double rcp_newton(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));  // MUFU.RCP64H: ~20 good bits
  double e = __fma_rn(-x, r, 1.0);
  r = __fma_rn(r, e, r);
  e = __fma_rn(-x, r, 1.0);
  r = __fma_rn(r, e, r);
  return r;
}

// exp(x) for |x| < 700 (here x = 17.67 (T-273.15)/(T-29.65), |x| < 40 for any T above 100 K)
__device__ __forceinline__ double exp_poly(double x) {
  const double tk = __fma_rn(x, kSat.l2e, kSat.magic);  // low word of tk = round(x log2 e)
  const double kf = tk - kSat.magic;
  double r = __fma_rn(-kf, kSat.ln2hi, x);
  r = __fma_rn(-kf, kSat.ln2lo, r);
  double p = kSat.p[13];
#pragma unroll
  for (int n = 12; n >= 0; --n) p = __fma_rn(p, r, kSat.p[n]);
  int k = __double2loint(tk);
  k = max(-1000, min(1000, k));
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));  // p * 2^k, p in [0.7, 1.42]
}

template <>
__device__ __forceinline__ void sat_adjust_point<double>(double& t, double& qv, double& l, const double pp) {
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double rtm = rcp_newton(t - kSat.c2965);
    const double es = kSat.c6112 * exp_poly(kSat.c1767 * (t - kSat.c27315) * rtm);
    const double rden = rcp_newton(__fma_rn(-kSat.one_m_eps, es, pp));
    const double qs = kSat.eps * es * rden;
    const double des = es * kSat.cdes * rtm * rtm;
    const double dqs = kSat.eps * pp * des * rden * rden;
    double dq = (qv - qs) * rcp_newton(__fma_rn(kSat.lcp, dqs, 1.0));
    dq = fmax(dq, -l);
    t = __fma_rn(kSat.lcp, dq, t);
    qv -= dq;
    l += dq;
  }
}

template <typename T, int W, int U>
__global__ void __launch_bounds__(kBlock) k_saturation_adjust(int niw, int nj, int nk, int ncols, int kchunk,
                                                              F3<const T> p, F3<T> Tt, F3<T> q, F3<T> ql) {
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (c >= ncols) return;
  const Col cc = decompose_column(c, niw, nj);
  const int i = cc.i * W;
  const int k0 = blockIdx.y * kchunk, k1 = min(nk, k0 + kchunk);
  // running pointers: the index arithmetic of four strided fields is paid once per thread, not per level
  T* pt = Tt.at(i, cc.j, k0, cc.b);
  T* pq = q.at(i, cc.j, k0, cc.b);
  T* pl = ql.at(i, cc.j, k0, cc.b);
  const T* pp = p.at(i, cc.j, k0, cc.b);
  for (int kb = k0; kb < k1; kb += U) {
    Vec<T, W> vt[U], vq[U], vl[U], vp[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (kb + u < k1) {
        vt[u] = VecIO<T, W>::ld(pt + u * Tt.sk);
        vq[u] = VecIO<T, W>::ld(pq + u * q.sk);
        vl[u] = VecIO<T, W>::ld(pl + u * ql.sk);
        vp[u] = VecIO<T, W>::ld(pp + u * p.sk);
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (kb + u < k1) {
#pragma unroll
        for (int w = 0; w < W; ++w) sat_adjust_point<T>(vt[u].v[w], vq[u].v[w], vl[u].v[w], vp[u].v[w]);
        VecIO<T, W>::st(pt + u * Tt.sk, vt[u]);
        VecIO<T, W>::st(pq + u * q.sk, vq[u]);
        VecIO<T, W>::st(pl + u * ql.sk, vl[u]);
      }
    pt += U * Tt.sk;
    pq += U * q.sk;
    pl += U * ql.sk;
    pp += U * p.sk;
  }
}

template <typename T>
int saturation_adjust(int ni, int nj, int nk, int nb, F3<const T> p, F3<T> Tt, F3<T> q, F3<T> ql, cudaStream_t s) {
  B2S_ARGCHECK(ni > 0 && nj > 0 && nk > 0 && nb > 0, "saturation_adjust: empty domain %dx%dx%dx%d", ni, nj, nk, nb);
  B2S_ARGCHECK(p.p && Tt.p && q.p && ql.p, "saturation_adjust: null field");
  constexpr int WMAX = MaxWidth<T>::value;
  const bool wide = ni % WMAX == 0 && WidthProbe(WMAX, sizeof(T)).field(p).field(Tt).field(q).field(ql).ok;
  const int W = wide ? WMAX : 1;
  const int ncols = (ni / W) * nj * nb;
  // pointwise: k is free to be split across blockIdx.y
  int kchunk = option("sat_kchunk", 0);
  // measured on C180x72 (profiles/r01_sat_sweep.json): 4 levels per thread; fp64 loads 2 levels per batch
  // (111.6 vs 102.5 Gpts/s), 8 or 12 levels per thread are 3-5 % slower
  if (kchunk <= 0) kchunk = 4;
  if (kchunk > nk) kchunk = nk;
  dim3 grid((ncols + kBlock - 1) / kBlock, (nk + kchunk - 1) / kchunk);
  int unroll = option("sat_unroll", 0);
  if (unroll == 0) unroll = sizeof(T) == 8 ? 2 : 1;
  if (wide) {
    if (unroll == 2)
      k_saturation_adjust<T, WMAX, 2><<<grid, kBlock, 0, s>>>(ni / W, nj, nk, ncols, kchunk, p, Tt, q, ql);
    else if (unroll == 4)
      k_saturation_adjust<T, WMAX, 4><<<grid, kBlock, 0, s>>>(ni / W, nj, nk, ncols, kchunk, p, Tt, q, ql);
    else
      k_saturation_adjust<T, WMAX, 1><<<grid, kBlock, 0, s>>>(ni / W, nj, nk, ncols, kchunk, p, Tt, q, ql);
  } else {
    k_saturation_adjust<T, 1, 4><<<grid, kBlock, 0, s>>>(ni, nj, nk, ncols, kchunk, p, Tt, q, ql);
  }
  return check_launch("saturation_adjust");
}

#define INSTANTIATE(T)                                                                                            \
  template int find_klcl<T>(int, int, int, int, F3<const T>, F2<const T>, F2<T>, F2<IndexOf<T>::type>, cudaStream_t); \
  template int cloud_top<T>(int, int, int, int, T, F3<const T>, F2<IndexOf<T>::type>, cudaStream_t);              \
  template int saturation_adjust<T>(int, int, int, int, F3<const T>, F3<T>, F3<T>, F3<T>, cudaStream_t);
INSTANTIATE(double)
INSTANTIATE(float)

}  // namespace impl
}  // namespace b2s
