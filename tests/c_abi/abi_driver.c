/* A compiled caller of libb200stencil's C-ABI with no Python in the process: the role the reference's
 * acceptance program plays for its generated bridge (test/py_ftn_interface/data/fortran_program.f90:24-36:
 * call through the bind(c) symbols, then check the array that came back).  It owns its device buffers
 * (cudaMalloc), passes raw pointers + strides + a stream, and checks the reference's golden vectors:
 *   Do__get_top_of_the_column.py:59-68  ones with 42 at the last level -> 42 everywhere
 *   Do__while_in_gt_functions.py:52-62  same input                     -> [3, 2, 1, 0] in every column
 * and one fv_tp2d call on a constant field (flux divergence of a constant with matching mass fluxes = 0).
 * Build: gcc abi_driver.c -I include -I $CUDA/include -L <libdir> -lb200stencil -L $CUDA/lib64 -lcudart
 * Exit code 0 = all checks passed; the failing check is printed otherwise. */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>

#include "b200stencil.h"

#define CHECK_CUDA(x)                                                         \
  do {                                                                        \
    cudaError_t e_ = (x);                                                     \
    if (e_ != cudaSuccess) {                                                  \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                \
      return 2;                                                               \
    }                                                                         \
  } while (0)
#define CHECK_B2S(x)                                                          \
  do {                                                                        \
    int rc_ = (x);                                                            \
    if (rc_ != 0) {                                                           \
      fprintf(stderr, "%s -> %d: %s\n", #x, rc_, b2s_last_error());           \
      return 3;                                                               \
    }                                                                         \
  } while (0)

int main(void) {
  enum { NI = 3, NJ = 3, NK = 4 };
  /* i-fastest: element (i, j, k) at i + j*sj + k*sk */
  const int64_t sj = NI, sk = NI * NJ, n3 = NI * NJ * NK, n2 = NI * NJ;
  double h_in[NI * NJ * NK], h_out[NI * NJ * NK], h_top[NI * NJ];
  for (int k = 0; k < NK; ++k)
    for (int c = 0; c < NI * NJ; ++c) h_in[c + k * sk] = k == NK - 1 ? 42.0 : 1.0;

  if (b2s_abi_version() <= 0) return 1;
  CHECK_B2S(b2s_init(0));
  cudaStream_t stream;
  CHECK_CUDA(cudaStreamCreate(&stream));
  double *d_in, *d_out, *d_top;
  CHECK_CUDA(cudaMalloc((void**)&d_in, n3 * sizeof(double)));
  CHECK_CUDA(cudaMalloc((void**)&d_out, n3 * sizeof(double)));
  CHECK_CUDA(cudaMalloc((void**)&d_top, n2 * sizeof(double)));
  CHECK_CUDA(cudaMemcpyAsync(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice, stream));
  CHECK_CUDA(cudaMemsetAsync(d_out, 0, sizeof(h_out), stream));
  CHECK_CUDA(cudaMemsetAsync(d_top, 0, sizeof(h_top), stream));

  /* S1 */
  CHECK_B2S(b2s_top_of_column_c(NI, NJ, NK, 1, d_in, sj, sk, 0, d_top, sj, 0, d_out, sj, sk, 0, stream));
  CHECK_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaMemcpyAsync(h_top, d_top, sizeof(h_top), cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaStreamSynchronize(stream));
  for (int c = 0; c < n3; ++c)
    if (h_out[c] != 42.0) return fprintf(stderr, "top_of_column: out[%d] = %g, expected 42\n", c, h_out[c]), 10;
  for (int c = 0; c < n2; ++c)
    if (h_top[c] != 42.0) return fprintf(stderr, "top_of_column: top[%d] = %g, expected 42\n", c, h_top[c]), 11;

  /* S2 */
  int64_t* d_cnt;
  int64_t h_cnt = -1;
  CHECK_CUDA(cudaMalloc((void**)&d_cnt, sizeof(int64_t)));
  CHECK_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int64_t), stream));
  CHECK_B2S(b2s_while_in_function_c(NI, NJ, NK, 1, 4.0, d_in, sj, sk, 0, d_out, sj, sk, 0, d_cnt, stream));
  CHECK_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaMemcpyAsync(&h_cnt, d_cnt, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaStreamSynchronize(stream));
  for (int k = 0; k < NK; ++k)
    for (int c = 0; c < n2; ++c)
      if (h_out[c + k * sk] != (double)(NK - 1 - k))
        return fprintf(stderr, "while_in_function: out[%d, k=%d] = %g, expected %d\n", c, k, h_out[c + k * sk], NK - 1 - k), 12;
  if (h_cnt != 0) return fprintf(stderr, "while_in_function: %lld undefined reads reported\n", (long long)h_cnt), 13;

  /* the error channel: a null field must come back as an argument error with a message, not a crash */
  if (b2s_top_of_column_c(NI, NJ, NK, 1, NULL, sj, sk, 0, d_top, sj, 0, d_out, sj, sk, 0, stream) >= 0 || !b2s_last_error()[0])
    return fprintf(stderr, "error channel: a NULL field was accepted\n"), 14;

  /* S5 on a constant field: q = 2 everywhere (halo included), crx = cry = 0.5, xfx = yfx = 1, rarea = 1 -> q_out = 2 */
  enum { FI = 40, FJ = 12, FK = 2, H = 3 };
  const int64_t qsj = FI + 2 * H, qsk = qsj * (FJ + 2 * H), xsj = FI + 1, xsk = xsj * FJ, ysj = FI, ysk = ysj * (FJ + 1);
  const int64_t osj = FI, osk = osj * FJ;
  const size_t nq = qsk * FK, nx = xsk * FK, ny = ysk * FK, no = osk * FK, nr = FI * FJ;
  double* h = (double*)malloc((nq > nx ? nq : nx) * sizeof(double));
  double *d_q, *d_cx, *d_xf, *d_cy, *d_yf, *d_ra, *d_qo;
  CHECK_CUDA(cudaMalloc((void**)&d_q, nq * sizeof(double)));
  CHECK_CUDA(cudaMalloc((void**)&d_cx, nx * sizeof(double)));
  CHECK_CUDA(cudaMalloc((void**)&d_xf, nx * sizeof(double)));
  CHECK_CUDA(cudaMalloc((void**)&d_cy, ny * sizeof(double)));
  CHECK_CUDA(cudaMalloc((void**)&d_yf, ny * sizeof(double)));
  CHECK_CUDA(cudaMalloc((void**)&d_ra, nr * sizeof(double)));
  CHECK_CUDA(cudaMalloc((void**)&d_qo, no * sizeof(double)));
#define FILL(dst, n, v)                                                                 \
  do {                                                                                  \
    for (size_t i_ = 0; i_ < (n); ++i_) h[i_] = (v);                                    \
    CHECK_CUDA(cudaMemcpy(dst, h, (n) * sizeof(double), cudaMemcpyHostToDevice));       \
  } while (0)
  FILL(d_q, nq, 2.0);
  FILL(d_cx, nx, 0.5);
  FILL(d_xf, nx, 1.0);
  FILL(d_cy, ny, 0.5);
  FILL(d_yf, ny, 1.0);
  FILL(d_ra, nr, 1.0);
  CHECK_CUDA(cudaMemset(d_qo, 0, no * sizeof(double)));
  /* q points at compute cell (0, 0, 0): the halo sits at negative offsets */
  CHECK_B2S(b2s_fv_tp2d_c(FI, FJ, FK, 1, 0, FI, 0, FJ, d_q + H + H * qsj, qsj, qsk, 0, d_cx, xsj, xsk, 0, d_xf, xsj, xsk, 0, d_cy, ysj,
                          ysk, 0, d_yf, ysj, ysk, 0, d_ra, FI, 0, d_qo, osj, osk, 0, stream));
  CHECK_CUDA(cudaMemcpyAsync(h, d_qo, no * sizeof(double), cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaStreamSynchronize(stream));
  for (size_t c = 0; c < no; ++c)
    if (h[c] < 2.0 - 1e-12 || h[c] > 2.0 + 1e-12) return fprintf(stderr, "fv_tp2d: q_out[%zu] = %.17g, expected 2\n", c, h[c]), 15;

  free(h);
  CHECK_B2S(b2s_finalize());
  printf("abi_driver ok: top_of_column, while_in_function, error channel, fv_tp2d through the C-ABI (abi %d, %d SMs)\n",
         b2s_abi_version(), b2s_sm_count());
  return 0;
}
