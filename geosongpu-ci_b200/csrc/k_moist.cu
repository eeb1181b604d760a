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
// Algorithmic bytes/point: 32 R + 24 W = 56.  In fp64 the two exp() and the divisions put
// ~170 DP operations on every point, close to the DFMA budget per point at HBM speed, so the
// reciprocals of (T - 29.65) and (p - (1-eps) es) are formed once per step and reused.
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

template <typename T, int W, int U>
__global__ void __launch_bounds__(kBlock) k_saturation_adjust(int niw, int nj, int nk, int ncols, int kchunk,
                                                              F3<const T> p, F3<T> Tt, F3<T> q, F3<T> ql) {
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (c >= ncols) return;
  const Col cc = decompose_column(c, niw, nj);
  const int i = cc.i * W;
  const int k0 = blockIdx.y * kchunk, k1 = min(nk, k0 + kchunk);
  for (int kb = k0; kb < k1; kb += U) {
    Vec<T, W> vt[U], vq[U], vl[U], vp[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (kb + u < k1) {
        vt[u] = VecIO<T, W>::ld(Tt.at(i, cc.j, kb + u, cc.b));
        vq[u] = VecIO<T, W>::ld(q.at(i, cc.j, kb + u, cc.b));
        vl[u] = VecIO<T, W>::ld(ql.at(i, cc.j, kb + u, cc.b));
        vp[u] = VecIO<T, W>::ld(p.at(i, cc.j, kb + u, cc.b));
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (kb + u < k1) {
#pragma unroll
        for (int w = 0; w < W; ++w) sat_adjust_point<T>(vt[u].v[w], vq[u].v[w], vl[u].v[w], vp[u].v[w]);
        VecIO<T, W>::st(Tt.at(i, cc.j, kb + u, cc.b), vt[u]);
        VecIO<T, W>::st(q.at(i, cc.j, kb + u, cc.b), vq[u]);
        VecIO<T, W>::st(ql.at(i, cc.j, kb + u, cc.b), vl[u]);
      }
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
  if (kchunk <= 0) kchunk = 4;  // measured best on C180x72 (profiles/): 4 levels per thread, 1 level per load batch
  if (kchunk > nk) kchunk = nk;
  dim3 grid((ncols + kBlock - 1) / kBlock, (nk + kchunk - 1) / kchunk);
  const int unroll = option("sat_unroll", 0);
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
