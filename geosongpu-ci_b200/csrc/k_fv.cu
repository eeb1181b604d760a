// K5 fv_tp2d dispatch: argument checks and variant selection (direct vs TMA-pipelined).
#include "halo_device.cuh"
#include "impl.cuh"

namespace b2s {
namespace impl {

template <typename T>
int fv_tp2d_direct(int ni, int nj, int nk, int nb, int i0, int i1, int j0, int j1, F3<const T> q, F3<const T> crx,
                   F3<const T> xfx, F3<const T> cry, F3<const T> yfx, F2<const T> rarea, F3<T> q_out, cudaStream_t s);
template <typename T>
int fv_tp2d_tma(int ni, int nj, int nk, int nb, int i0, int i1, int j0, int j1, F3<const T> q, F3<const T> crx,
                F3<const T> xfx, F3<const T> cry, F3<const T> yfx, F2<const T> rarea, F3<T> q_out, cudaStream_t s,
                bool* applicable, int* gate, const HaloXchg* xchg);

template <typename T>
int fv_tp2d_stream(int ni, int nj, int nk, int nb, int i0, int i1, int j0, int j1, F3<const T> q, F3<const T> crx,
                   F3<const T> xfx, F3<const T> cry, F3<const T> yfx, F2<const T> rarea, F3<T> q_out, cudaStream_t s,
                   bool* applicable, int* gate, const HaloXchg* xchg);

template <typename T>
static int fv_tp2d_any(int ni, int nj, int nk, int nb, int i0, int i1, int j0, int j1, F3<const T> q, F3<const T> crx,
                       F3<const T> xfx, F3<const T> cry, F3<const T> yfx, F2<const T> rarea, F3<T> q_out, int* gate,
                       const HaloXchg* xchg, cudaStream_t s) {
  B2S_ARGCHECK(ni > 0 && nj > 0 && nk > 0 && nb > 0, "fv_tp2d: empty domain %dx%dx%dx%d", ni, nj, nk, nb);
  B2S_ARGCHECK(0 <= i0 && i0 <= i1 && i1 <= ni && 0 <= j0 && j0 <= j1 && j1 <= nj,
               "fv_tp2d: rectangle [%d,%d)x[%d,%d) outside the %dx%d domain", i0, i1, j0, j1, ni, nj);
  B2S_ARGCHECK(q.p && crx.p && xfx.p && cry.p && yfx.p && rarea.p && q_out.p, "fv_tp2d: null field");
  if (i0 == i1 || j0 == j1) return B2S_OK;
  const int variant = option("fv_variant", 0);
  // auto: the streaming kernel (every row staged once; 453 us against the tile kernel's 516 us on C384x72 fp64,
  // 237 against 282 us on 3 x 384 x 384 x 72), except for launches below ~12 M points where the tile kernel's
  // finer work items fill the machine better (3 x 192 x 192 x 72, the 8-GPU sub-domains: 61.0 against 63.1 us);
  // the tile kernel as the second TMA choice, the direct kernel for fields TMA cannot address
  const int64_t small_below = option("fv_small_points", 0) > 0 ? (int64_t)option("fv_small_points", 0) : 12000000;
  const bool small = (int64_t)(i1 - i0) * (j1 - j0) * nk * nb < small_below;
  if (variant == 3 || (variant == 0 && !small)) {
    bool applicable = false;
    int rc = fv_tp2d_stream<T>(ni, nj, nk, nb, i0, i1, j0, j1, q, crx, xfx, cry, yfx, rarea, q_out, s, &applicable, gate, xchg);
    if (applicable) return rc;
    if (variant == 3) return set_error(B2S_EUNSUPPORTED, "fv_tp2d: fv_variant=3 forced but fields do not meet the TMA alignment rules");
  }
  if (variant != 1) {
    bool applicable = false;
    int rc = fv_tp2d_tma<T>(ni, nj, nk, nb, i0, i1, j0, j1, q, crx, xfx, cry, yfx, rarea, q_out, s, &applicable, gate, xchg);
    if (applicable) return rc;
    if (variant == 2) return set_error(B2S_EUNSUPPORTED, "fv_tp2d: fv_variant=2 forced but fields do not meet the TMA alignment rules");
  }
  if (gate != nullptr)
    return set_error(B2S_EUNSUPPORTED, "fv_tp2d_gated / halo_fv_tp2d: the fields do not meet the TMA alignment rules (16-byte aligned rows); "
                                       "use b2s_halo_exchange + b2s_fv_tp2d instead");
  return fv_tp2d_direct<T>(ni, nj, nk, nb, i0, i1, j0, j1, q, crx, xfx, cry, yfx, rarea, q_out, s);
}

template <typename T>
int fv_tp2d(int ni, int nj, int nk, int nb, int i0, int i1, int j0, int j1, F3<const T> q, F3<const T> crx,
            F3<const T> xfx, F3<const T> cry, F3<const T> yfx, F2<const T> rarea, F3<T> q_out, cudaStream_t s) {
  return fv_tp2d_any<T>(ni, nj, nk, nb, i0, i1, j0, j1, q, crx, xfx, cry, yfx, rarea, q_out, nullptr, nullptr, s);
}

// fv_tp2d on the whole batch, overlapped with the halo update of q that is still in flight (b2s_halo_exchange_start
// with gated = 1): the kernel walks the sub-domains in batch order and the producer warp of every CTA acquires
// gate[b] (b2s_halo_gate) before its first TMA load of sub-domain b, so sub-domain b is computed while the halos
// of b+1.. are still arriving.  nb must be the number of sub-domains of the exchanged field.  Same bits as fv_tp2d.
template <typename T>
int fv_tp2d_gated(int ni, int nj, int nk, int nb, F3<const T> q, F3<const T> crx, F3<const T> xfx, F3<const T> cry,
                  F3<const T> yfx, F2<const T> rarea, int* gate, F3<T> q_out, cudaStream_t s) {
  B2S_ARGCHECK(gate != nullptr, "fv_tp2d_gated: gate is NULL (b2s_halo_gate)");
  B2S_ARGCHECK(nb <= 64, "fv_tp2d_gated: %d sub-domains, the gate has 64 slots", nb);
  return fv_tp2d_any<T>(ni, nj, nk, nb, 0, ni, 0, nj, q, crx, xfx, cry, yfx, rarea, q_out, gate, nullptr, s);
}

// Halo update of q + fv_tp2d in ONE launch (b2s_halo_fv_tp2d, csrc/halo_ctx.cu): every CTA of the persistent stencil grid
// first takes its share of the exchange (neighbour handshake + strip copies over peer memory), then walks its stencil
// items in sub-domain order behind the gates the exchange opens.  Needs the TMA kernels (all CTAs are co-resident).
template <typename T>
int fv_tp2d_fused(int ni, int nj, int nk, int nb, F3<const T> q, F3<const T> crx, F3<const T> xfx, F3<const T> cry,
                  F3<const T> yfx, F2<const T> rarea, F3<T> q_out, const HaloXchg& xchg, cudaStream_t s) {
  B2S_ARGCHECK(nb <= 64, "halo_fv_tp2d: %d sub-domains, the gate has 64 slots", nb);
  B2S_ARGCHECK(xchg.links != nullptr && xchg.gated, "halo_fv_tp2d: no exchange to fuse");
  return fv_tp2d_any<T>(ni, nj, nk, nb, 0, ni, 0, nj, q, crx, xfx, cry, yfx, rarea, q_out, xchg.state + kGateWord, &xchg, s);
}

template int fv_tp2d_fused<double>(int, int, int, int, F3<const double>, F3<const double>, F3<const double>, F3<const double>,
                                   F3<const double>, F2<const double>, F3<double>, const HaloXchg&, cudaStream_t);
template int fv_tp2d_fused<float>(int, int, int, int, F3<const float>, F3<const float>, F3<const float>, F3<const float>,
                                  F3<const float>, F2<const float>, F3<float>, const HaloXchg&, cudaStream_t);

template int fv_tp2d_gated<double>(int, int, int, int, F3<const double>, F3<const double>, F3<const double>, F3<const double>,
                                   F3<const double>, F2<const double>, int*, F3<double>, cudaStream_t);
template int fv_tp2d_gated<float>(int, int, int, int, F3<const float>, F3<const float>, F3<const float>, F3<const float>,
                                  F3<const float>, F2<const float>, int*, F3<float>, cudaStream_t);

template int fv_tp2d<double>(int, int, int, int, int, int, int, int, F3<const double>, F3<const double>,
                             F3<const double>, F3<const double>, F3<const double>, F2<const double>, F3<double>,
                             cudaStream_t);
template int fv_tp2d<float>(int, int, int, int, int, int, int, int, F3<const float>, F3<const float>,
                            F3<const float>, F3<const float>, F3<const float>, F2<const float>, F3<float>,
                            cudaStream_t);

}  // namespace impl
}  // namespace b2s
