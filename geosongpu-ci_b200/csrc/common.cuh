// Shared device/host helpers of libb200stencil (sm_100a only).
//
// Data layout contract (DESIGN.md "Data layout"): every field is i-fastest.  A 3-D field
// travels through the C-ABI as (ptr, sj, sk, sb): element (b, i, j, k) lives at
// ptr[i + j*sj + k*sk + b*sb]; an IJ field as (ptr, sj, sb).  Strides are in elements.
// This is the zero-copy view the reference builds from Fortran memory
// (src/tcn/py_ftn_interface/templates/data_conversion.py:134-148).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2s {

template <typename T>
struct F3 {
  T* p;
  int64_t sj, sk, sb;
  __host__ __device__ __forceinline__ T* at(int64_t i, int64_t j, int64_t k, int64_t b) const {
    return p + i + j * sj + k * sk + b * sb;
  }
};

template <typename T>
struct F2 {
  T* p;
  int64_t sj, sb;
  __host__ __device__ __forceinline__ T* at(int64_t i, int64_t j, int64_t b) const { return p + i + j * sj + b * sb; }
};

template <typename T>
struct IndexOf;
template <>
struct IndexOf<double> {
  using type = int64_t;
};
template <>
struct IndexOf<float> {
  using type = int32_t;
};

// status codes of the C-ABI
enum : int { B2S_OK = 0, B2S_EINVAL = -1, B2S_ENOTINIT = -2, B2S_EUNSUPPORTED = -3 };

int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> status
int sm_count();
int option(const char* name, int fallback);

// streaming (evict-first) loads/stores for data touched exactly once
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) {
  return __ldcs(p);
}
template <typename T>
__device__ __forceinline__ void st_stream(T* p, T v) {
  __stcs(p, v);
}

struct Col {
  int i, j, b;
};
__device__ __forceinline__ Col decompose_column(int c, int ni, int nj) {
  Col r;
  r.i = c % ni;
  int t = c / ni;
  r.j = t % nj;
  r.b = t / nj;
  return r;
}

}  // namespace b2s

#define B2S_ARGCHECK(cond, ...)                                        \
  do {                                                                 \
    if (!(cond)) return b2s::set_error(b2s::B2S_EINVAL, __VA_ARGS__); \
  } while (0)

// flat C-ABI arguments -> field structs
#define B2S_F3(T, name) \
  b2s::F3<T> { name, name##_sj, name##_sk, name##_sb }
#define B2S_F2(T, name) \
  b2s::F2<T> { name, name##_sj, name##_sb }
