// K5 (variant 1, "direct"): fv_tp2d with one thread per cell, neighbours read through L1/L2.
// Spec: oracle/numpy_oracle.py fv_tp2d (no source in /root/reference).
// This is the general-stride path and the on-device cross-check of the TMA-pipelined kernel
// (k_fv_tma.cu); it is selected when a field does not meet the TMA alignment rules or with
// b2s_set_option("fv_variant", 1).
// Algorithmic bytes/point: 40 R (q, crx, xfx, cry, yfx) + 8 W + 8/nk (rarea).
#include "fv_math.cuh"
#include "impl.cuh"

namespace b2s {
namespace impl {

template <typename T>
__global__ void __launch_bounds__(256) k_fv_direct(int nk, int i0, int i1, int j0, int j1, F3<const T> q,
                                                   F3<const T> crx, F3<const T> xfx, F3<const T> cry,
                                                   F3<const T> yfx, F2<const T> rarea, F3<T> qout) {
  const int i = i0 + blockIdx.x * 32 + threadIdx.x;
  const int j = j0 + blockIdx.y * 8 + threadIdx.y;
  const int k = blockIdx.z % nk;
  const int b = blockIdx.z / nk;
  if (i >= i1 || j >= j1) return;
  const T* qc = q.at(i, j, k, b);
  const T q0 = __ldg(qc);
  // x direction
  T fx_lo, fx_hi;
  {
    const T m3 = __ldg(qc - 3), m2 = __ldg(qc - 2), m1 = __ldg(qc - 1);
    const T p1 = __ldg(qc + 1), p2 = __ldg(qc + 2), p3 = __ldg(qc + 3);
    const T al_m1 = ppm_al(m3, m2, m1, q0), al_0 = ppm_al(m2, m1, q0, p1);
    const T al_p1 = ppm_al(m1, q0, p1, p2), al_p2 = ppm_al(q0, p1, p2, p3);
    const T* cp = crx.at(i, j, k, b);
    const T* xp = xfx.at(i, j, k, b);
    fx_lo = mul_rn(ppm_flux_from_al(m1, q0, al_m1, al_0, al_p1, __ldg(cp)), __ldg(xp));
    fx_hi = mul_rn(ppm_flux_from_al(q0, p1, al_0, al_p1, al_p2, __ldg(cp + 1)), __ldg(xp + 1));
  }
  // y direction
  T fy_lo, fy_hi;
  {
    const int64_t s = q.sj;
    const T m3 = __ldg(qc - 3 * s), m2 = __ldg(qc - 2 * s), m1 = __ldg(qc - s);
    const T p1 = __ldg(qc + s), p2 = __ldg(qc + 2 * s), p3 = __ldg(qc + 3 * s);
    const T al_m1 = ppm_al(m3, m2, m1, q0), al_0 = ppm_al(m2, m1, q0, p1);
    const T al_p1 = ppm_al(m1, q0, p1, p2), al_p2 = ppm_al(q0, p1, p2, p3);
    const T* cp = cry.at(i, j, k, b);
    const T* yp = yfx.at(i, j, k, b);
    fy_lo = mul_rn(ppm_flux_from_al(m1, q0, al_m1, al_0, al_p1, __ldg(cp)), __ldg(yp));
    fy_hi = mul_rn(ppm_flux_from_al(q0, p1, al_0, al_p1, al_p2, __ldg(cp + cry.sj)), __ldg(yp + yfx.sj));
  }
  const T ra = __ldg(rarea.at(i, j, b));
  __stcs(qout.at(i, j, k, b), fv_update(q0, ra, fx_lo, fx_hi, fy_lo, fy_hi));
}

template <typename T>
int fv_tp2d_direct(int ni, int nj, int nk, int nb, int i0, int i1, int j0, int j1, F3<const T> q, F3<const T> crx,
                   F3<const T> xfx, F3<const T> cry, F3<const T> yfx, F2<const T> rarea, F3<T> q_out,
                   cudaStream_t s) {
  (void)ni;
  (void)nj;
  dim3 block(32, 8);
  dim3 grid((i1 - i0 + 31) / 32, (j1 - j0 + 7) / 8, nk * nb);
  B2S_ARGCHECK(grid.y <= 65535 && grid.z <= 65535, "fv_tp2d: grid too large (%u,%u,%u)", grid.x, grid.y, grid.z);
  k_fv_direct<T><<<grid, block, 0, s>>>(nk, i0, i1, j0, j1, q, crx, xfx, cry, yfx, rarea, q_out);
  return check_launch("fv_tp2d(direct)");
}

template int fv_tp2d_direct<double>(int, int, int, int, int, int, int, int, F3<const double>, F3<const double>,
                                    F3<const double>, F3<const double>, F3<const double>, F2<const double>,
                                    F3<double>, cudaStream_t);
template int fv_tp2d_direct<float>(int, int, int, int, int, int, int, int, F3<const float>, F3<const float>,
                                   F3<const float>, F3<const float>, F3<const float>, F2<const float>, F3<float>,
                                   cudaStream_t);

}  // namespace impl
}  // namespace b2s
