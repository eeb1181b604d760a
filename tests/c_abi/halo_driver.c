/* The multi-GPU halo lifecycle of libb200stencil driven from C with no Python in the process: two ranks (threads here,
 * one per virtual GPU on device 0; processes on separate GPUs take the same path through cudaIpc) meet through
 * b2s_halo_init, allocate a symmetric field, bind hand-written link tables of a periodic two-sub-domain ring and
 * exchange halos -- b2s_halo_exchange and the forked b2s_halo_exchange_start / _wait -- then run the gated stencil.
 * It stands where the reference's Fortran acceptance program passes its MPI communicator through the bridge
 * (/root/reference/src/tcn/py_ftn_interface/example_def_dycore.yaml:4-21 `comm`, argument.py:54-86 type MPI).
 * Build: gcc halo_driver.c -I include -I $CUDA/include -L <libdir> -lb200stencil -L $CUDA/lib64 -lcudart -lpthread
 * Exit code 0 = all checks passed. */
#include <cuda_runtime_api.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "b200stencil.h"

enum { NI = 40, NJ = 12, NK = 3, H = 3, WORLD = 2 };
static const int64_t SJ = NI + 2 * H, SK = (NI + 2 * H) * (NJ + 2 * H);
static char g_session[64];
static int g_fail[WORLD];

static double value_of(int rank, int i, int j, int k) { return rank * 1e6 + i * 1e4 + j * 100 + k; }

#define FAIL(code, ...)                  \
  do {                                   \
    fprintf(stderr, __VA_ARGS__);        \
    g_fail[rank] = (code);               \
    return NULL;                         \
  } while (0)
#define B2S(x)                                                                       \
  do {                                                                               \
    int rc_ = (x);                                                                   \
    if (rc_ != 0) FAIL(3, "rank %d: %s -> %d: %s\n", rank, #x, rc_, b2s_last_error()); \
  } while (0)
#define CUDA(x)                                                                      \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) FAIL(2, "rank %d: %s: %s\n", rank, #x, cudaGetErrorString(e_)); \
  } while (0)

static void* rank_main(void* arg) {
  const int rank = (int)(intptr_t)arg, other = 1 - rank;
  const size_t n = (size_t)SK * NK;
  CUDA(cudaSetDevice(0));
  cudaStream_t stream;
  CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  int64_t ctx = 0;
  B2S(b2s_halo_init(g_session, rank, WORLD, 0, &ctx));
  if (b2s_halo_rank(ctx) != rank || b2s_halo_world(ctx) != WORLD) FAIL(20, "rank %d: context reports %d of %d\n", rank, b2s_halo_rank(ctx), b2s_halo_world(ctx));
  void* p = NULL;
  B2S(b2s_halo_alloc(ctx, (int64_t)(n * sizeof(double)), &p));
  double* q = (double*)p;
  void* peer = NULL;
  B2S(b2s_halo_peer_ptr(ctx, q, other, &peer));
  if (peer == NULL || peer == (void*)q) FAIL(21, "rank %d: peer pointer %p\n", rank, peer);

  /* periodic ring in i: west halo <- the other rank's east columns, east halo <- its west columns */
  const int64_t links[2][12] = {
      {H + NI - 1 + H * SJ, -1, SJ, SK, H - 1 + H * SJ, -1, SJ, SK, H, NJ, other, 0},
      {H + H * SJ, 1, SJ, SK, H + NI + H * SJ, 1, SJ, SK, H, NJ, other, 0},
  };
  int plan = -1;
  B2S(b2s_halo_plan(ctx, q, (int)sizeof(double), NK, 2, &links[0][0], &plan));
  if (b2s_halo_plan_remote_bytes(ctx, plan) != (int64_t)2 * H * NJ * NK * (int64_t)sizeof(double))
    FAIL(22, "rank %d: remote bytes %lld\n", rank, (long long)b2s_halo_plan_remote_bytes(ctx, plan));

  double* h = (double*)malloc(n * sizeof(double));
  for (int rep = 0; rep < 3; ++rep) {
    for (size_t c = 0; c < n; ++c) h[c] = -1.0;
    for (int k = 0; k < NK; ++k)
      for (int j = 0; j < NJ; ++j)
        for (int i = 0; i < NI; ++i) h[(i + H) + (j + H) * SJ + k * SK] = value_of(rank, i, j, k) + rep;
    CUDA(cudaMemcpyAsync(q, h, n * sizeof(double), cudaMemcpyHostToDevice, stream));
    CUDA(cudaStreamSynchronize(stream));
    B2S(b2s_halo_barrier(ctx)); /* both interiors are in place (the host rewrote them) */
    if (rep == 1) {
      B2S(b2s_halo_exchange_start(ctx, plan, 0, stream));
      B2S(b2s_halo_exchange_wait(ctx, stream));
    } else {
      B2S(b2s_halo_exchange(ctx, plan, stream));
    }
    CUDA(cudaMemcpyAsync(h, q, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CUDA(cudaStreamSynchronize(stream));
    for (int k = 0; k < NK; ++k)
      for (int j = 0; j < NJ; ++j)
        for (int d = 0; d < H; ++d) {
          const double w = h[(H - 1 - d) + (j + H) * SJ + k * SK], e = h[(H + NI + d) + (j + H) * SJ + k * SK];
          if (w != value_of(other, NI - 1 - d, j, k) + rep) FAIL(23, "rank %d rep %d: west halo (%d,%d,%d) = %.1f\n", rank, rep, d, j, k, w);
          if (e != value_of(other, d, j, k) + rep) FAIL(24, "rank %d rep %d: east halo (%d,%d,%d) = %.1f\n", rank, rep, d, j, k, e);
        }
    B2S(b2s_halo_barrier(ctx)); /* the peer has pulled: the interior may be rewritten */
  }

  /* overlapped step: constant field, exchange forked, gated stencil; flux divergence of a constant is zero */
  const int64_t xsj = NI + 2, xsk = xsj * NJ, ysj = NI, ysk = ysj * (NJ + 1), osj = NI, osk = osj * NJ;
  const size_t nx = xsk * NK, ny = ysk * NK, no = osk * NK, nr = NI * NJ;
  double *cx, *xf, *cy, *yf, *ra, *qo;
  CUDA(cudaMalloc((void**)&cx, nx * sizeof(double)));
  CUDA(cudaMalloc((void**)&xf, nx * sizeof(double)));
  CUDA(cudaMalloc((void**)&cy, ny * sizeof(double)));
  CUDA(cudaMalloc((void**)&yf, ny * sizeof(double)));
  CUDA(cudaMalloc((void**)&ra, nr * sizeof(double)));
  CUDA(cudaMalloc((void**)&qo, no * sizeof(double)));
  double* hh = (double*)malloc((nx > n ? nx : n) * sizeof(double));
#define FILL(dst, cnt, v)                                                              \
  do {                                                                                 \
    for (size_t i_ = 0; i_ < (cnt); ++i_) hh[i_] = (v);                                \
    CUDA(cudaMemcpy(dst, hh, (cnt) * sizeof(double), cudaMemcpyHostToDevice));         \
  } while (0)
  FILL(q, n, 2.0);
  FILL(cx, nx, 0.5);
  FILL(xf, nx, 1.0);
  FILL(cy, ny, -0.5);
  FILL(yf, ny, 1.0);
  FILL(ra, nr, 1.0);
  CUDA(cudaMemset(qo, 0, no * sizeof(double)));
  int* gate = NULL;
  B2S(b2s_halo_gate(ctx, &gate));
  B2S(b2s_halo_barrier(ctx));
  for (int rep = 0; rep < 2; ++rep) {
    B2S(b2s_halo_exchange_start(ctx, plan, 1, stream));
    B2S(b2s_fv_tp2d_gated_c(NI, NJ, NK, 1, q + H + H * SJ, SJ, SK, 0, cx, xsj, xsk, 0, xf, xsj, xsk, 0, cy, ysj, ysk, 0, yf, ysj, ysk, 0,
                            ra, NI, 0, gate, qo, osj, osk, 0, stream));
    B2S(b2s_halo_exchange_wait(ctx, stream));
  }
  CUDA(cudaMemcpyAsync(hh, qo, no * sizeof(double), cudaMemcpyDeviceToHost, stream));
  CUDA(cudaStreamSynchronize(stream));
  for (size_t c = 0; c < no; ++c)
    if (hh[c] < 2.0 - 1e-12 || hh[c] > 2.0 + 1e-12) FAIL(25, "rank %d: gated fv_tp2d q_out[%zu] = %.17g, expected 2\n", rank, c, hh[c]);
  int epoch = -1, status = -1;
  B2S(b2s_halo_status(ctx, &epoch, &status));
  if (epoch != 5 || status != 0) FAIL(26, "rank %d: epoch %d status %d, expected 5 and 0\n", rank, epoch, status);
  B2S(b2s_halo_free(ctx, q));
  B2S(b2s_halo_finalize(ctx));
  if (b2s_halo_barrier(ctx) == 0) FAIL(27, "rank %d: a finalized context was accepted\n", rank);
  free(h);
  free(hh);
  return NULL;
}

int main(void) {
  int rc = b2s_init(0);
  if (rc != 0) {
    fprintf(stderr, "b2s_init(0) -> %d: %s\n", rc, b2s_last_error());
    return 3;
  }
  snprintf(g_session, sizeof(g_session), "halo_driver_%ld", (long)getpid());
  pthread_t t[WORLD];
  for (int r = 0; r < WORLD; ++r) pthread_create(&t[r], NULL, rank_main, (void*)(intptr_t)r);
  for (int r = 0; r < WORLD; ++r) pthread_join(t[r], NULL);
  for (int r = 0; r < WORLD; ++r)
    if (g_fail[r]) return g_fail[r];
  printf("halo_driver ok: 2 ranks, symmetric allocation, 3 exchanges (plain + forked), gated fv_tp2d, finalize\n");
  return 0;
}
