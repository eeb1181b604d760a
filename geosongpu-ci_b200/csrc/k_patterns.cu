// K1-K3: the three gt4py/NDSL column patterns of /root/reference/dsl_patterns as sm_100a kernels.
//
// Mapping (DESIGN.md "Kernels"): one thread owns W adjacent columns (W*sizeof(T) = 16 B when
// alignment allows), consecutive threads own consecutive i -> every warp access is a run of
// full 128-B lines.  The k loop lives inside the thread, the k-carry in registers; U levels
// are loaded before they are consumed so each thread keeps U independent 16-B requests in
// flight.  All three are HBM-bound; no shared memory, no tensor cores.
#include "impl.cuh"
#include "vec.cuh"

namespace b2s {
namespace impl {

static constexpr int kBlock = 128;

// -------------------------------------------------------------------------------------------
// K1 top_of_column -- dsl_patterns/Do__get_top_of_the_column.py:33-38
//   FORWARD interval(-1,None): PLEmb_top = PLEmb ; PARALLEL interval(...): out_field = PLEmb_top
// Algorithmic bytes/point: 8 W + 16/nk (read bottom level, write the IJ field).
// k is split across blockIdx.y when there are too few columns to fill the machine.
// -------------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(kBlock) k_top_of_column(int niw, int nj, int nk, int ncols, int kchunk,
                                                          F3<const T> in, F2<T> top, F3<T> out) {
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (c >= ncols) return;
  const Col cc = decompose_column(c, niw, nj);
  const int i = cc.i * W;
  const Vec<T, W> v = VecIO<T, W>::ld(in.at(i, cc.j, nk - 1, cc.b));
  const int k0 = blockIdx.y * kchunk;
  const int k1 = min(nk, k0 + kchunk);
  if (k1 == nk) VecIO<T, W>::st(top.at(i, cc.j, cc.b), v);  // the chunk owning the last level
  T* o = out.at(i, cc.j, k0, cc.b);
#pragma unroll 8
  for (int k = k0; k < k1; ++k, o += out.sk) VecIO<T, W>::st(o, v);
}

template <typename T>
int top_of_column(int ni, int nj, int nk, int nb, F3<const T> PLEmb, F2<T> PLEmb_top, F3<T> out_field,
                  cudaStream_t s) {
  B2S_ARGCHECK(ni > 0 && nj > 0 && nk > 0 && nb > 0, "top_of_column: empty domain %dx%dx%dx%d", ni, nj, nk, nb);
  B2S_ARGCHECK(PLEmb.p && PLEmb_top.p && out_field.p, "top_of_column: null field");
  constexpr int WMAX = MaxWidth<T>::value;
  const bool wide = ni % WMAX == 0 && WidthProbe(WMAX, sizeof(T)).field(PLEmb).field(PLEmb_top).field(out_field).ok;
  const int W = wide ? WMAX : 1;
  const int ncols = (ni / W) * nj * nb;
  // aim for >= 8 resident warps per SM scheduler slot: split k while the grid is thin
  const int target_threads = sm_count() * 2048;
  int ksplit = 1;
  while (ksplit < nk && (int64_t)ncols * ksplit < target_threads && nk / (ksplit * 2) >= 8) ksplit *= 2;
  const int kchunk = (nk + ksplit - 1) / ksplit;
  dim3 grid((ncols + kBlock - 1) / kBlock, (nk + kchunk - 1) / kchunk);
  if (wide)
    k_top_of_column<T, WMAX><<<grid, kBlock, 0, s>>>(ni / W, nj, nk, ncols, kchunk, PLEmb, PLEmb_top, out_field);
  else
    k_top_of_column<T, 1><<<grid, kBlock, 0, s>>>(ni, nj, nk, ncols, kchunk, PLEmb, PLEmb_top, out_field);
  return check_launch("top_of_column");
}

// -------------------------------------------------------------------------------------------
// K2 while_in_function -- dsl_patterns/Do__while_in_gt_functions.py:22-32
//   per point: lev = 0; while in[k+lev] < thr: lev += 1; out = lev
// Restated as ONE backward scan: nxt = (in[k] >= thr) ? k : nxt; out = nxt-k,
// identical wherever the reference is defined (a level >= thr exists at or below k).  Where it
// is not, out = nk-k and the point is counted in *undefined_count.
// Algorithmic bytes/point: 8 R + 8 W.
// -------------------------------------------------------------------------------------------
template <typename T, int W, int U>
__global__ void __launch_bounds__(kBlock) k_while_in_function(int niw, int nj, int nk, int ncols, T thr,
                                                              F3<const T> in, F3<T> out,
                                                              unsigned long long* undefined_count) {
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (c >= ncols) return;
  const Col cc = decompose_column(c, niw, nj);
  const int i = cc.i * W;
  const T* ip = in.at(i, cc.j, 0, cc.b);
  T* op = out.at(i, cc.j, 0, cc.b);
  int nxt[W];
#pragma unroll
  for (int w = 0; w < W; ++w) nxt[w] = nk;
  unsigned int undefined = 0;
  for (int kb = nk; kb > 0; kb -= U) {
    Vec<T, W> x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = kb - 1 - u;
      if (k >= 0) x[u] = VecIO<T, W>::ld(ip + (int64_t)k * in.sk);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = kb - 1 - u;
      if (k >= 0) {
        Vec<T, W> r;
#pragma unroll
        for (int w = 0; w < W; ++w) {
          if (!(x[u].v[w] < thr)) nxt[w] = k;  // the reference's loop stops on `not (field < 4)`
          undefined += (nxt[w] == nk);
          r.v[w] = static_cast<T>(nxt[w] - k);
        }
        VecIO<T, W>::st(op + (int64_t)k * out.sk, r);
      }
    }
  }
  if (undefined && undefined_count) atomicAdd(undefined_count, (unsigned long long)undefined);
}

// Thin grids (C96x72: 55 296 columns, a third of the machine's thread slots) make the scan above
// latency-bound: a column is one thread and its levels arrive in nk/U dependent batches.  Here the
// column is cut into S segments of L levels, one thread each (CTA = 32*W columns x S segments): every
// thread issues ALL its loads at once, finds the first hit of its segment, the segments exchange
// that through shared memory (the k-carry of the BACKWARD scan is "the first hit below me"), and
// each thread then writes its own levels.  Same values as the scan, one DRAM round trip deep.
template <typename T, int W, int L>
__global__ void __launch_bounds__(1024) k_while_ksplit(int niw, int nj, int nk, int ncols, T thr, F3<const T> in,
                                                       F3<T> out, unsigned long long* undefined_count) {
  extern __shared__ int first_hit[];  // [segment][32*W]
  const int lane = threadIdx.x, seg = threadIdx.y, nseg = blockDim.y;
  const int c = blockIdx.x * 32 + lane;
  const bool valid = c < ncols;
  const Col cc = decompose_column(valid ? c : 0, niw, nj);
  const int i = cc.i * W;
  const int k0 = seg * L;
  const T* ip = in.at(i, cc.j, k0, cc.b);
  T* op = out.at(i, cc.j, k0, cc.b);
  Vec<T, W> x[L];
#pragma unroll
  for (int u = 0; u < L; ++u)
    if (valid && k0 + u < nk) x[u] = VecIO<T, W>::ld(ip + (int64_t)u * in.sk);
  int nxt[W];
#pragma unroll
  for (int w = 0; w < W; ++w) nxt[w] = nk;
#pragma unroll
  for (int u = L - 1; u >= 0; --u)
    if (valid && k0 + u < nk) {
#pragma unroll
      for (int w = 0; w < W; ++w)
        if (!(x[u].v[w] < thr)) nxt[w] = k0 + u;
    }
#pragma unroll
  for (int w = 0; w < W; ++w) first_hit[(seg * 32 + lane) * W + w] = nxt[w];
  __syncthreads();
  if (!valid) return;
  // the carry entering this segment from below: the first hit of the nearest lower segment that has one
#pragma unroll
  for (int w = 0; w < W; ++w) nxt[w] = nk;
  for (int s2 = nseg - 1; s2 > seg; --s2) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int f = first_hit[(s2 * 32 + lane) * W + w];
      if (f < nk) nxt[w] = f;
    }
  }
  unsigned int undefined = 0;
#pragma unroll
  for (int u = L - 1; u >= 0; --u) {
    const int k = k0 + u;
    if (k < nk) {
      Vec<T, W> r;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        if (!(x[u].v[w] < thr)) nxt[w] = k;
        undefined += (nxt[w] == nk);
        r.v[w] = static_cast<T>(nxt[w] - k);
      }
      VecIO<T, W>::st(op + (int64_t)u * out.sk, r);
    }
  }
  if (undefined && undefined_count) atomicAdd(undefined_count, (unsigned long long)undefined);
}

// b2s_set_option("while_variant", v): 0 auto (k-split on thin grids), 1 column scan, 2 k-split
template <typename T, int W>
int while_ksplit_launch(int niw, int nj, int nk, int ncols, T thr, F3<const T> in, F3<T> out, unsigned long long* cnt,
                        cudaStream_t s) {
  constexpr int L = 9;
  const int nseg = (nk + L - 1) / L;
  dim3 block(32, nseg);
  const size_t smem = (size_t)nseg * 32 * W * sizeof(int);
  k_while_ksplit<T, W, L><<<(ncols + 31) / 32, block, smem, s>>>(niw, nj, nk, ncols, thr, in, out, cnt);
  return check_launch("while_in_function");
}

template <typename T>
int while_in_function(int ni, int nj, int nk, int nb, T threshold, F3<const T> in_field, F3<T> out_field,
                      int64_t* undefined_count, cudaStream_t s) {
  B2S_ARGCHECK(ni > 0 && nj > 0 && nk > 0 && nb > 0, "while_in_function: empty domain %dx%dx%dx%d", ni, nj, nk, nb);
  B2S_ARGCHECK(in_field.p && out_field.p, "while_in_function: null field");
  constexpr int WMAX = MaxWidth<T>::value;
  const bool aligned = ni % WMAX == 0 && WidthProbe(WMAX, sizeof(T)).field(in_field).field(out_field).ok;
  const bool wide = aligned && (int64_t)(ni / WMAX) * nj * nb >= (int64_t)sm_count() * 1024;
  auto* cnt = reinterpret_cast<unsigned long long*>(undefined_count);
  const int variant = option("while_variant", 0);
  const bool thin = (int64_t)ni * nj * nb < (int64_t)sm_count() * 2048;  // fewer columns than thread slots
  if (variant != 1 && nk <= 9 * 32 && (variant == 2 || thin)) {
    if (aligned) return while_ksplit_launch<T, WMAX>(ni / WMAX, nj, nk, (ni / WMAX) * nj * nb, threshold, in_field, out_field, cnt, s);
    return while_ksplit_launch<T, 1>(ni, nj, nk, ni * nj * nb, threshold, in_field, out_field, cnt, s);
  }
  if (wide) {
    const int ncols = (ni / WMAX) * nj * nb;
    k_while_in_function<T, WMAX, 8>
        <<<(ncols + kBlock - 1) / kBlock, kBlock, 0, s>>>(ni / WMAX, nj, nk, ncols, threshold, in_field, out_field, cnt);
  } else {
    const int ncols = ni * nj * nb;
    // thin grids (C96: 55k columns, a fraction of one wave) are latency-bound: 24 levels in flight per thread
    if ((int64_t)ncols < (int64_t)sm_count() * 1024)
      k_while_in_function<T, 1, 24>
          <<<(ncols + kBlock - 1) / kBlock, kBlock, 0, s>>>(ni, nj, nk, ncols, threshold, in_field, out_field, cnt);
    else
      k_while_in_function<T, 1, 12>
          <<<(ncols + kBlock - 1) / kBlock, kBlock, 0, s>>>(ni, nj, nk, ncols, threshold, in_field, out_field, cnt);
  }
  return check_launch("while_in_function");
}

// -------------------------------------------------------------------------------------------
// K3 hybrid_index_2dout -- dsl_patterns/WIP__hybrid_index_2dout.py:34-42
//   FORWARD over all k: if k_mask == k_index_desired: out_field = data_field   (last match wins)
// The scan keeps the LAST matching level in a register and reads data_field once, at that
// level; a column without a match leaves out_field untouched.  Arbitrary k_mask contents are
// honoured (every level of k_mask is read).
// Algorithmic bytes/point: 16 R + 16/nk; DRAM traffic is ~8 R because data_field
// is only touched in the sectors that hold a match.
// -------------------------------------------------------------------------------------------
template <typename T, int W, int U>
__global__ void __launch_bounds__(kBlock) k_hybrid_index(int niw, int nj, int nk, int ncols, F3<const T> data,
                                                         F3<const T> kmask, F2<const T> kidx, F2<T> out) {
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (c >= ncols) return;
  const Col cc = decompose_column(c, niw, nj);
  const int i = cc.i * W;
  const Vec<T, W> want = VecIO<T, W>::ld(kidx.at(i, cc.j, cc.b));
  const T* mp = kmask.at(i, cc.j, 0, cc.b);
  int last[W];
#pragma unroll
  for (int w = 0; w < W; ++w) last[w] = -1;
  for (int kb = 0; kb < nk; kb += U) {
    Vec<T, W> m[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (kb + u < nk) m[u] = VecIO<T, W>::ld(mp + (int64_t)(kb + u) * kmask.sk);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (kb + u < nk) {
#pragma unroll
        for (int w = 0; w < W; ++w)
          if (m[u].v[w] == want.v[w]) last[w] = kb + u;
      }
  }
#pragma unroll
  for (int w = 0; w < W; ++w)
    if (last[w] >= 0) *out.at(i + w, cc.j, cc.b) = __ldg(data.at(i + w, cc.j, last[w], cc.b));
}

template <typename T>
int hybrid_index_2dout(int ni, int nj, int nk, int nb, F3<const T> data_field, F3<const T> k_mask,
                       F2<const T> k_index_desired, F2<T> out_field, cudaStream_t s) {
  B2S_ARGCHECK(ni > 0 && nj > 0 && nk > 0 && nb > 0, "hybrid_index_2dout: empty domain %dx%dx%dx%d", ni, nj, nk, nb);
  B2S_ARGCHECK(data_field.p && k_mask.p && k_index_desired.p && out_field.p, "hybrid_index_2dout: null field");
  constexpr int WMAX = MaxWidth<T>::value;
  const bool wide = ni % WMAX == 0 && WidthProbe(WMAX, sizeof(T)).field(k_mask).field(k_index_desired).ok &&
                    (int64_t)(ni / WMAX) * nj * nb >= (int64_t)sm_count() * 1024;
  if (wide) {
    const int ncols = (ni / WMAX) * nj * nb;
    k_hybrid_index<T, WMAX, 8>
        <<<(ncols + kBlock - 1) / kBlock, kBlock, 0, s>>>(ni / WMAX, nj, nk, ncols, data_field, k_mask, k_index_desired, out_field);
  } else {
    const int ncols = ni * nj * nb;
    k_hybrid_index<T, 1, 12>
        <<<(ncols + kBlock - 1) / kBlock, kBlock, 0, s>>>(ni, nj, nk, ncols, data_field, k_mask, k_index_desired, out_field);
  }
  return check_launch("hybrid_index_2dout");
}

#define INSTANTIATE(T)                                                                                             \
  template int top_of_column<T>(int, int, int, int, F3<const T>, F2<T>, F3<T>, cudaStream_t);                      \
  template int while_in_function<T>(int, int, int, int, T, F3<const T>, F3<T>, int64_t*, cudaStream_t);            \
  template int hybrid_index_2dout<T>(int, int, int, int, F3<const T>, F3<const T>, F2<const T>, F2<T>, cudaStream_t);
INSTANTIATE(double)
INSTANTIATE(float)

}  // namespace impl
}  // namespace b2s
