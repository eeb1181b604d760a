// Host side of the TMA kernels: cuTensorMapEncodeTiled through the runtime's driver entry point
// (no link against libcuda), with a small cache keyed by everything a descriptor depends on.
#include <map>
#include <mutex>
#include <tuple>

#include "tma.cuh"

namespace b2s {
namespace impl {

namespace {

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* base;
  int64_t sj, sk, sb;
  int e0, e1, e2, e3, b0, b1, b2, es;
  bool operator<(const MapKey& o) const {
    return std::tie(base, sj, sk, sb, e0, e1, e2, e3, b0, b1, b2, es) <
           std::tie(o.base, o.sj, o.sk, o.sb, o.e0, o.e1, o.e2, o.e3, o.b0, o.b1, o.b2, o.es);
  }
};
std::map<MapKey, CUtensorMap> g_maps;
std::mutex g_maps_mu;

}  // namespace

bool make_tensor_map(CUtensorMap* out, const void* base, int elem_size, int64_t sj, int64_t sk, int64_t sb, int e0,
                     int e1, int e2, int e3, int b0, int b1, int b2) {
  MapKey key{base, sj, sk, sb, e0, e1, e2, e3, b0, b1, b2, elem_size};
  {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) {
      *out = it->second;
      return true;
    }
  }
  EncodeFn enc = encode_fn();
  if (!enc) return false;
  // size-1 axes still need a legal (multiple-of-16, non-zero) stride
  const int64_t sk_b = (e2 > 1 ? sk : sj * e1) * (int64_t)elem_size;
  const int64_t sb_b = (e3 > 1 ? sb : (e2 > 1 ? sk * e2 : sj * e1)) * (int64_t)elem_size;
  cuuint64_t dims[4] = {(cuuint64_t)e0, (cuuint64_t)e1, (cuuint64_t)e2, (cuuint64_t)e3};
  cuuint64_t strides[3] = {(cuuint64_t)(sj * elem_size), (cuuint64_t)sk_b, (cuuint64_t)sb_b};
  cuuint32_t box[4] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapDataType dt = elem_size == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(out, dt, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  std::lock_guard<std::mutex> lk(g_maps_mu);
  if (g_maps.size() > 256) g_maps.clear();
  g_maps[key] = *out;
  return true;
}

}  // namespace impl
}  // namespace b2s
