/* The call sequence of include/b2s_example.f90, statement for statement, in C: this image has no Fortran compiler,
 * so the GPU box executes the mirror.  A bind(c) call with `value` scalars, `type(c_ptr), value` addresses and
 * by-reference intent(out) results IS this C call (tests/fortran_check.py checks the declarations against the
 * header); what is left to prove is that the sequence and its constants are right.  In addition to the status checks
 * of the Fortran program the mirror fills the fields and checks the east halo after the exchange. */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "b200stencil.h"

enum { ni = 48, nj = 48, nk = 4 };

#define STOP(code, what)                                                          \
  do {                                                                            \
    fprintf(stderr, "stop %d: %s: %s\n", code, what, b2s_last_error());            \
    return code;                                                                  \
  } while (0)

int main(void) {
  const int64_t sj = ni + 6 + ni % 2, sk = sj * (nj + 6), sb = sk * nk;
  const int64_t fj = ni + 2, fk = fj * (nj + 1), fb = fk * nk;
  int64_t ctx = 0, links[12];
  int status, plan = -1, epoch = -1, xstatus = -1;
  void *q = NULL, *crx = NULL, *xfx = NULL, *cry = NULL, *yfx = NULL, *rarea = NULL, *q_out = NULL;

  status = b2s_init(0);
  if (status != 0) {
    fprintf(stderr, "b2s_init: %s\n", b2s_last_error());
    return 3; /* the no-GPU exit code the other drivers use */
  }
  if (b2s_abi_version() != 1) STOP(2, "b2s_abi_version");
  status = b2s_halo_init("", 0, 1, 0, &ctx);
  if (status != 0) STOP(3, "b2s_halo_init");
  status = b2s_halo_alloc(ctx, 8 * sb, &q);
  if (status != 0 || q == NULL) STOP(4, "b2s_halo_alloc");
  if (b2s_halo_alloc(ctx, 8 * fb, &crx) != 0) STOP(4, "b2s_halo_alloc");
  if (b2s_halo_alloc(ctx, 8 * fb, &xfx) != 0) STOP(4, "b2s_halo_alloc");
  if (b2s_halo_alloc(ctx, 8 * fb, &cry) != 0) STOP(4, "b2s_halo_alloc");
  if (b2s_halo_alloc(ctx, 8 * fb, &yfx) != 0) STOP(4, "b2s_halo_alloc");
  if (b2s_halo_alloc(ctx, 8 * fb, &rarea) != 0) STOP(4, "b2s_halo_alloc");
  if (b2s_halo_alloc(ctx, 8 * fb, &q_out) != 0) STOP(4, "b2s_halo_alloc");

  /* (mirror only) known contents: q(i, j, k) = i + 100 j + 10000 k in halo-origin coordinates, zeros elsewhere */
  double* hq = (double*)malloc(8 * sb);
  for (int k = 0; k < nk; ++k)
    for (int j = 0; j < nj + 6; ++j)
      for (int i = 0; i < sj; ++i) hq[i + j * sj + k * sk] = i + 100.0 * j + 10000.0 * k;
  cudaMemcpy(q, hq, 8 * sb, cudaMemcpyHostToDevice);
  cudaMemset(crx, 0, 8 * fb), cudaMemset(xfx, 0, 8 * fb), cudaMemset(cry, 0, 8 * fb), cudaMemset(yfx, 0, 8 * fb);
  cudaMemset(rarea, 0, 8 * fb), cudaMemset(q_out, 0xff, 8 * fb);

  /* one link: the first three interior columns of the tile -> its own east halo (a periodic strip) */
  const int64_t row[12] = {3 + 3 * sj, 1, sj, sk, ni + 3 + 3 * sj, 1, sj, sk, 3, nj, 0, 0};
  for (int w = 0; w < 12; ++w) links[w] = row[w];
  status = b2s_halo_plan(ctx, q, 8, nk, 1, links, &plan);
  if (status != 0) STOP(5, "b2s_halo_plan");
  status = b2s_halo_exchange(ctx, plan, NULL);
  if (status != 0) STOP(6, "b2s_halo_exchange");
  status = b2s_halo_status(ctx, &epoch, &xstatus);
  if (status != 0 || epoch != 1 || xstatus != 0) STOP(7, "b2s_halo_status");

  /* (mirror only) the east halo now holds the first three interior columns */
  cudaMemcpy(hq, q, 8 * sb, cudaMemcpyDeviceToHost);
  for (int k = 0; k < nk; ++k)
    for (int j = 3; j < nj + 3; ++j)
      for (int d = 0; d < 3; ++d)
        if (hq[ni + 3 + d + j * sj + k * sk] != (3 + d) + 100.0 * j + 10000.0 * k) {
          fprintf(stderr, "east halo (%d, %d, %d) = %g\n", d, j, k, hq[ni + 3 + d + j * sj + k * sk]);
          return 10;
        }

  /* the stencils take the address of compute cell (0,0,0): three rows and three columns into the halo-carrying field */
  double* q_core = (double*)((intptr_t)q + 8 * (3 + 3 * sj));
  status = b2s_fv_tp2d_c(ni, nj, nk, 1, 0, ni, 0, nj, q_core, sj, sk, sb, (const double*)crx, fj, fk, fb, (const double*)xfx, fj, fk, fb,
                         (const double*)cry, fj, fk, fb, (const double*)yfx, fj, fk, fb, (const double*)rarea, fj, fb, (double*)q_out, fj,
                         fk, fb, NULL);
  if (status != 0) STOP(8, "b2s_fv_tp2d_c");
  /* (mirror only) zero Courant numbers, zero fluxes, zero rarea: q_out = q on the compute domain */
  double* ho = (double*)malloc(8 * fb);
  if (cudaMemcpy(ho, q_out, 8 * fb, cudaMemcpyDeviceToHost) != cudaSuccess) return 11;
  for (int k = 0; k < nk; ++k)
    for (int j = 0; j < nj; ++j)
      for (int i = 0; i < ni; ++i)
        if (ho[i + j * fj + k * fk] != (i + 3) + 100.0 * (j + 3) + 10000.0 * k) {
          fprintf(stderr, "q_out (%d, %d, %d) = %g\n", i, j, k, ho[i + j * fj + k * fk]);
          return 12;
        }
  status = b2s_halo_finalize(ctx);
  if (status != 0) STOP(9, "b2s_halo_finalize");
  status = b2s_finalize();
  free(hq), free(ho);
  printf("b2s_example: ok\n");
  return 0;
}
