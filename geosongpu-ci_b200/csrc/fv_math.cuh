// PPM flux arithmetic shared by the fv_tp2d kernels (spec: oracle/numpy_oracle.py _ppm_flux).
//   al = 7/12 (q[-1] + q) - 1/12 (q[-2] + q[+1])        value at the low-side interface of a cell
//   bl = al - q ; br = al[+1] - q ; b0 = bl + br
//   flux(c > 0) = q[-1] + (1 - c)(br[-1] - c b0[-1]) ;  flux(c <= 0) = q + (1 + c)(bl + c b0)
//
// Every operation is written with an explicit round-to-nearest intrinsic, so the compiler cannot
// choose a different FMA contraction in a different kernel: the direct kernel, every TMA tile
// geometry and every interior/frame split produce the SAME BITS for a cell (tests assert it).
// Three products are fused (one rounding fewer than the un-contracted oracle); the results stay
// within 1e-12 relative of it.
#pragma once
#include "common.cuh"

namespace b2s {

__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }

template <typename T>
__device__ __forceinline__ T ppm_al(T qm2, T qm1, T q0, T qp1) {
  // 7/12 (qm1 + q0) - 1/12 (qm2 + qp1)
  return fma_rn(T(-1.0 / 12.0), add_rn(qm2, qp1), mul_rn(T(7.0 / 12.0), add_rn(qm1, q0)));
}

// flux through the interface between cell L (low side) and cell H (high side).
//   al_L, al_H, al_HH: interface values at the low side of cells L, H and H+1
template <typename T>
__device__ __forceinline__ T ppm_flux_from_al(T qL, T qH, T al_L, T al_H, T al_HH, T c) {
  // Branch-free, two selects.  With qu the upwind cell value, al_H the interface value AT the flux
  // interface and a_far the interface value on the far side of the upwind cell:
  //   c > 0 : br = al_H - qL, bl = al_L  - qL      c <= 0 : bl = al_H - qH, br = al_HH - qH
  // so in both cases  bx = al_H - qu,  b0 = bx + (a_far - qu),  flux = qu + (1-|c|)(bx - |c| b0),
  // the oracle's two-branch form term for term (1 + c == 1 - |c| and bl + c*b0 == bl - |c|*b0 for c <= 0).
  const bool pos = c > T(0);
  const T ac = pos ? c : -c;
  const T qu = pos ? qL : qH;
  const T a_far = pos ? al_L : al_HH;
  const T bx = sub_rn(al_H, qu);
  const T b0 = add_rn(bx, sub_rn(a_far, qu));
  return fma_rn(sub_rn(T(1.0), ac), fma_rn(-ac, b0, bx), qu);
}

// q - rarea * ((fx_hi - fx_lo) + (fy_hi - fy_lo))
template <typename T>
__device__ __forceinline__ T fv_update(T q0, T ra, T fx_lo, T fx_hi, T fy_lo, T fy_hi) {
  return fma_rn(-ra, add_rn(sub_rn(fx_hi, fx_lo), sub_rn(fy_hi, fy_lo)), q0);
}

}  // namespace b2s
