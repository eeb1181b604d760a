// K5 (variant 2, TMA-pipelined) -- placeholder until the kernel lands: reports "not applicable".
#include "impl.cuh"

namespace b2s {
namespace impl {

template <typename T>
int fv_tp2d_tma(int, int, int, int, int, int, int, int, F3<const T>, F3<const T>, F3<const T>, F3<const T>,
                F3<const T>, F2<const T>, F3<T>, cudaStream_t, bool* applicable) {
  *applicable = false;
  return B2S_OK;
}

template int fv_tp2d_tma<double>(int, int, int, int, int, int, int, int, F3<const double>, F3<const double>,
                                 F3<const double>, F3<const double>, F3<const double>, F2<const double>,
                                 F3<double>, cudaStream_t, bool*);
template int fv_tp2d_tma<float>(int, int, int, int, int, int, int, int, F3<const float>, F3<const float>,
                                F3<const float>, F3<const float>, F3<const float>, F2<const float>, F3<float>,
                                cudaStream_t, bool*);

}  // namespace impl
}  // namespace b2s
