// PPM flux arithmetic shared by the fv_tp2d kernels (spec: oracle/numpy_oracle.py _ppm_flux).
//   al = 7/12 (q[-1] + q) - 1/12 (q[-2] + q[+1])        value at the low-side interface of a cell
//   bl = al - q ; br = al[+1] - q ; b0 = bl + br
//   flux(c > 0) = q[-1] + (1 - c)(br[-1] - c b0[-1]) ;  flux(c <= 0) = q + (1 + c)(bl + c b0)
#pragma once
#include "common.cuh"

namespace b2s {

template <typename T>
__device__ __forceinline__ T ppm_al(T qm2, T qm1, T q0, T qp1) {
  return T(7.0 / 12.0) * (qm1 + q0) - T(1.0 / 12.0) * (qm2 + qp1);
}

// flux through the interface between cell L (low side) and cell H (high side).
//   al_L, al_H, al_HH: interface values at the low side of cells L, H and H+1
template <typename T>
__device__ __forceinline__ T ppm_flux_from_al(T qL, T qH, T al_L, T al_H, T al_HH, T c) {
  const T one = T(1.0);
  if (c > T(0)) {
    const T bl = al_L - qL, br = al_H - qL, b0 = bl + br;
    return qL + (one - c) * (br - c * b0);
  } else {
    const T bl = al_H - qH, br = al_HH - qH, b0 = bl + br;
    return qH + (one + c) * (bl + c * b0);
  }
}

}  // namespace b2s
