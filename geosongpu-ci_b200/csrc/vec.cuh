// 16-byte vector access helpers: W consecutive i-cells per thread (W = 2 doubles / 4 floats)
// when every pointer and stride of a call allows it, W = 1 otherwise.
#pragma once
#include "common.cuh"

namespace b2s {

template <typename T, int W>
struct Vec {
  T v[W];
};

template <typename T>
struct MaxWidth {
  static constexpr int value = 16 / sizeof(T);
};

// ---- loads (streaming, evict-first: every element of a stencil input is read once) ----
template <typename T, int W>
struct VecIO;

template <typename T>
struct VecIO<T, 1> {
  static __device__ __forceinline__ Vec<T, 1> ld(const T* p) {
    Vec<T, 1> r;
    r.v[0] = __ldcs(p);
    return r;
  }
  static __device__ __forceinline__ void st(T* p, const Vec<T, 1>& x) { __stcs(p, x.v[0]); }
};
template <>
struct VecIO<double, 2> {
  static __device__ __forceinline__ Vec<double, 2> ld(const double* p) {
    double2 t = __ldcs(reinterpret_cast<const double2*>(p));
    Vec<double, 2> r;
    r.v[0] = t.x;
    r.v[1] = t.y;
    return r;
  }
  static __device__ __forceinline__ void st(double* p, const Vec<double, 2>& x) {
    __stcs(reinterpret_cast<double2*>(p), make_double2(x.v[0], x.v[1]));
  }
};
template <>
struct VecIO<float, 4> {
  static __device__ __forceinline__ Vec<float, 4> ld(const float* p) {
    float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    Vec<float, 4> r;
    r.v[0] = t.x;
    r.v[1] = t.y;
    r.v[2] = t.z;
    r.v[3] = t.w;
    return r;
  }
  static __device__ __forceinline__ void st(float* p, const Vec<float, 4>& x) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(x.v[0], x.v[1], x.v[2], x.v[3]));
  }
};
template <>
struct VecIO<int64_t, 2> {
  static __device__ __forceinline__ void st(int64_t* p, const Vec<int64_t, 2>& x) {
    __stcs(reinterpret_cast<longlong2*>(p), make_longlong2(x.v[0], x.v[1]));
  }
};
template <>
struct VecIO<int32_t, 4> {
  static __device__ __forceinline__ void st(int32_t* p, const Vec<int32_t, 4>& x) {
    __stcs(reinterpret_cast<int4*>(p), make_int4(x.v[0], x.v[1], x.v[2], x.v[3]));
  }
};
template <>
struct VecIO<int64_t, 1> {
  static __device__ __forceinline__ void st(int64_t* p, const Vec<int64_t, 1>& x) {
    __stcs(reinterpret_cast<long long*>(p), static_cast<long long>(x.v[0]));
  }
};
template <>
struct VecIO<int32_t, 1> {
  static __device__ __forceinline__ void st(int32_t* p, const Vec<int32_t, 1>& x) { __stcs(p, x.v[0]); }
};

// ---- host-side width selection ----
struct WidthProbe {
  int w;  // candidate width in elements
  size_t elem;
  bool ok = true;
  WidthProbe(int w_, size_t elem_) : w(w_), elem(elem_) {}
  template <typename T>
  WidthProbe& field(F3<T> f) {
    ok = ok && (reinterpret_cast<uintptr_t>(f.p) % (w * elem) == 0) && f.sj % w == 0 && f.sk % w == 0 && f.sb % w == 0;
    return *this;
  }
  template <typename T>
  WidthProbe& field(F2<T> f) {
    ok = ok && (reinterpret_cast<uintptr_t>(f.p) % (w * elem) == 0) && f.sj % w == 0 && f.sb % w == 0;
    return *this;
  }
};

}  // namespace b2s
