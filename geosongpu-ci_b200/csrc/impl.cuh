// Declarations of the stencil launchers behind the C-ABI (definitions in k_*.cu, explicit
// instantiation for double and float).  Argument order = YAML order (inputs, inouts, outputs).
#pragma once
#include "common.cuh"

namespace b2s {
namespace impl {

template <typename T>
int top_of_column(int ni, int nj, int nk, int nb, F3<const T> PLEmb, F2<T> PLEmb_top, F3<T> out_field, cudaStream_t s);
template <typename T>
int while_in_function(int ni, int nj, int nk, int nb, T threshold, F3<const T> in_field, F3<T> out_field,
                      int64_t* undefined_count, cudaStream_t s);
template <typename T>
int hybrid_index_2dout(int ni, int nj, int nk, int nb, F3<const T> data_field, F3<const T> k_mask,
                       F2<const T> k_index_desired, F2<T> out_field, cudaStream_t s);

template <typename T>
int find_klcl(int ni, int nj, int nk, int nb, F3<const T> PLmb, F2<const T> PLCL, F2<T> PLmb_at_KLCL,
              F2<typename IndexOf<T>::type> KLCL, cudaStream_t s);
template <typename T>
int saturation_adjust(int ni, int nj, int nk, int nb, F3<const T> p, F3<T> Tt, F3<T> q, F3<T> ql, cudaStream_t s);
template <typename T>
int cloud_top(int ni, int nj, int nk, int nb, T ql_min, F3<const T> ql, F2<typename IndexOf<T>::type> ktop,
              cudaStream_t s);

template <typename T>
int fv_tp2d(int ni, int nj, int nk, int nb, int i0, int i1, int j0, int j1, F3<const T> q, F3<const T> crx,
            F3<const T> xfx, F3<const T> cry, F3<const T> yfx, F2<const T> rarea, F3<T> q_out, cudaStream_t s);

template <typename T>
int fv_tp2d_gated(int ni, int nj, int nk, int nb, F3<const T> q, F3<const T> crx, F3<const T> xfx, F3<const T> cry,
                  F3<const T> yfx, F2<const T> rarea, int* gate, F3<T> q_out, cudaStream_t s);

// halo update of q + fv_tp2d in one launch; ctx / plan from b2s_halo_init / b2s_halo_plan (csrc/halo_ctx.cu)
template <typename T>
int halo_fv_tp2d(int64_t ctx, int plan, int ni, int nj, int nk, int nb, F3<const T> crx, F3<const T> xfx, F3<const T> cry,
                 F3<const T> yfx, F2<const T> rarea, F3<T> q, F3<T> q_out, cudaStream_t s);

template <typename T>
int fv_tp2d_split(int ni, int nj, int nk, int nb, F3<const T> q, F3<const T> crx, F3<const T> xfx, F3<const T> cry,
                  F3<const T> yfx, F2<const T> area, F2<const T> rarea, const int* corner_flags, F3<T> q_out,
                  F3<T> fx_out, F3<T> fy_out, cudaStream_t s);

template <typename T>
int pe_prefix(int ni, int nj, int nk, int nb, T ptop, F3<const T> delp, F3<T> pe, cudaStream_t s);
template <typename T>
int remap(int ni, int nj, int nk1, int nk2, int nb, F3<const T> pe1, F3<const T> q1, F3<const T> pe2, F3<T> q2,
          cudaStream_t s);
template <typename T>
int remap_delp(int ni, int nj, int nk1, int nk2, int nb, T ptop, F3<const T> delp, F3<const T> q1, F3<const T> pe2,
               F3<T> q2, cudaStream_t s);
template <typename T>
int remap_ppm(int ni, int nj, int nk1, int nk2, int nb, int kord, int iv, F3<const T> pe1, F3<const T> q1,
              F3<const T> pe2, F3<T> q2, cudaStream_t s);
template <typename T>
int tridiag(int ni, int nj, int nk, int nb, F3<const T> a, F3<const T> b, F3<const T> c, F3<const T> d, F3<T> w,
            F3<T> x, cudaStream_t s);

template <typename T>
int halo_move(int nlinks, int nk, int max_strip, const int64_t* links, const T* src, T* dst, cudaStream_t s);

template <typename T>
int halo_pull(int nlinks, int nk, int max_strip, const int64_t* links, T* dst, cudaStream_t s);

// Per-device kernel set-up caches are indexed by the CUDA device ordinal.
constexpr int kMaxDevices = 64;
// Load (CUDA loads kernels lazily) and set up, on the current device, every kernel that can run beside -- or be
// waited for by -- a spinning exchange kernel: called by b2s_halo_init, before the first exchange of the context.
int fv_tma_preload();
int fv_stream_preload();
int halo_kernels_preload();
}  // namespace impl
}  // namespace b2s
