// K7 halo_move: pack / unpack / same-GPU neighbour copy of halo strips, one kernel for all three.
// Replaces the pack-unpack half of NDSL's HaloUpdater (not in /root/reference; SURVEY.md 2.1 K7,
// 8e).  A "link" is an affine strip copy described by 10 int64 (device array `links`):
//   [0] src_off [1] src_sd [2] src_sp [3] src_sk   element offset/strides into `src`
//   [4] dst_off [5] dst_sd [6] dst_sp [7] dst_sk   element offset/strides into `dst`
//   [8] nd  (halo depth, 3)   [9] np (edge length)
//   dst[dst_off + d*dst_sd + p*dst_sp + k*dst_sk] = src[src_off + d*src_sd + p*src_sp + k*src_sk]
// Negative strides express the index reversal / rotation across cubed-sphere tile edges.
//   pack:   src = field, dst = send buffer      unpack: src = recv buffer, dst = field
//   local:  src = dst = field (neighbouring sub-domains resident on the same GPU)
#include "impl.cuh"

namespace b2s {
namespace impl {

static constexpr int kLinkWords = 10;

template <typename T>
__global__ void __launch_bounds__(256) k_halo_move(int nk, const int64_t* __restrict__ links, const T* src, T* dst) {
  const int64_t* L = links + (int64_t)blockIdx.y * kLinkWords;
  const int nd = (int)L[8], np = (int)L[9];
  const int64_t total = (int64_t)nd * np * nk;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(t % np);
    const int64_t r = t / np;
    const int d = (int)(r % nd);
    const int k = (int)(r / nd);
    dst[L[4] + d * L[5] + p * L[6] + k * L[7]] = src[L[0] + d * L[1] + p * L[2] + k * L[3]];
  }
}

template <typename T>
int halo_move(int nlinks, int nk, const int64_t* links, const T* src, T* dst, cudaStream_t s) {
  B2S_ARGCHECK(nlinks >= 0 && nk > 0, "halo_move: bad sizes nlinks=%d nk=%d", nlinks, nk);
  if (nlinks == 0) return B2S_OK;
  B2S_ARGCHECK(links && src && dst, "halo_move: null pointer");
  dim3 grid(16, nlinks);
  k_halo_move<T><<<grid, 256, 0, s>>>(nk, links, src, dst);
  return check_launch("halo_move");
}

template int halo_move<double>(int, int, const int64_t*, const double*, double*, cudaStream_t);
template int halo_move<float>(int, int, const int64_t*, const float*, float*, cudaStream_t);

}  // namespace impl
}  // namespace b2s
