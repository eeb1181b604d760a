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
  // Branch-free, two selects.  With qu the upwind cell value, al_H the interface value AT the flux
  // interface and a_far the interface value on the far side of the upwind cell:
  //   c > 0 : br = al_H - qL, bl = al_L  - qL      c <= 0 : bl = al_H - qH, br = al_HH - qH
  // so in both cases  bx = al_H - qu,  b0 = bx + (a_far - qu),  flux = qu + (1-|c|)(bx - |c| b0).
  // Bitwise identical to the two-branch form of the oracle (1 + c == 1 - |c| and
  // bl + c*b0 == bl - |c|*b0 for c <= 0; bl + br commutes).
  const bool pos = c > T(0);
  const T ac = pos ? c : -c;
  const T qu = pos ? qL : qH;
  const T a_far = pos ? al_L : al_HH;
  const T bx = al_H - qu;
  const T b0 = bx + (a_far - qu);
  return qu + (T(1.0) - ac) * (bx - ac * b0);
}

}  // namespace b2s
