// K5 (variant 3, streaming): fv_tp2d marching down j -- every input row enters shared memory ONCE.
// Spec: oracle/numpy_oracle.py fv_tp2d (no source in /root/reference).
//
// The tile kernel (k_fv_tma.cu) stages R + 6 rows of q for every R output rows (R = 4: each q row is
// brought into shared memory 2.5 times, from L2 mostly) and warms the y-window up for every tile (three
// extra interface values and one extra flux per four rows).  Here a work item is a strip of TI columns of
// one (k, b) level and JB rows (128 by default), and the CTA marches through rows r = j0-3 .. j1+2 once:
//   * the PRODUCER warp streams 4-row chunks through an NSTAGE-deep full/empty mbarrier ring: q with the
//     x-apron, crx / xfx of the same rows, cry / yfx two rows behind (row r brings the y-interface r-2 that
//     can be closed once r has arrived); five TMA tile loads per chunk.  The ring runs on across items;
//   * CONSUMER thread = one column.  When row r arrives it (1) computes the x-flux difference of row r from
//     the shared row (seven neighbours) and parks it in a four-deep register ring, (2) adds row r to the
//     y-window (three q values, two interface values, the previous interface's flux in registers), closes
//     y-interface r-2, and (3) stores row r-3: q - rarea ((fx_hi - fx_lo) + (fy_hi - fy_lo)).
//     Nothing crosses threads except through the TMA-written rows: no __syncthreads in the loop.
// Same formulas, same explicit-rounding arithmetic as the other variants: identical bits (tests assert it).
// Algorithmic bytes/point: 40 R + 8 W + 8/nk; smem fill per point: (JB + 6) / JB of it.
#include <mutex>

#include "fv_math.cuh"
#include "halo_device.cuh"
#include "impl.cuh"
#include "tma.cuh"

namespace b2s {
namespace impl {

namespace {

constexpr int ru(int x, int m) { return (x + m - 1) / m * m; }
constexpr int RB = 4;  // rows per chunk = depth of the x-flux register ring

template <typename T, int TI, int NSTAGE>
struct STile {
  static constexpr int V = 16 / sizeof(T);
  static constexpr int BQ = ru(TI + 6 + V - 1, V);  // q box: columns i_s-3 .. i_s+TI+2 (+ alignment shift)
  static constexpr int BX = ru(TI + 1 + V - 1, V);  // crx / xfx box: interfaces i_s .. i_s+TI
  static constexpr int BY = ru(TI + V - 1, V);      // cry / yfx box
  static constexpr int Q_OFF = 0;
  static constexpr int CRX_OFF = ru(RB * BQ * (int)sizeof(T), 128);
  static constexpr int XFX_OFF = CRX_OFF + ru(RB * BX * (int)sizeof(T), 128);
  static constexpr int CRY_OFF = XFX_OFF + ru(RB * BX * (int)sizeof(T), 128);
  static constexpr int YFX_OFF = CRY_OFF + ru(RB * BY * (int)sizeof(T), 128);
  static constexpr int STAGE_BYTES = YFX_OFF + ru(RB * BY * (int)sizeof(T), 128);
  static constexpr int TX_BYTES = RB * (BQ + 2 * BX + 2 * BY) * (int)sizeof(T);
  static constexpr int BAR_OFF = NSTAGE * STAGE_BYTES;
  static constexpr int SCRATCH_OFF = BAR_OFF + 2 * NSTAGE * 8;  // one int for the fused exchange phase
  static constexpr int SMEM_BYTES = SCRATCH_OFF + 16;
  static constexpr int THREADS = TI + 32;
  static_assert(BQ <= 256 && TI % 32 == 0 && TI % V == 0, "tile width");
};

template <typename T>
struct FvStreamParams {
  int nk, i0, i1, j0, j1;
  int nstrips, njblk, jb, nitems;
  int c_q, c_crx, c_xfx, c_cry, c_yfx;       // TMA coordinate of compute column i0 (tensor bases are 16-byte aligned)
  int sh_q, sh_crx, sh_xfx, sh_cry, sh_yfx;  // c % V: the box starts at the aligned column before
  F2<const T> rarea;
  F3<T> qout;
  // gated launch (fv_tp2d_gated): gate[b] is raised by the halo exchange kernel when every halo cell of sub-domain b
  // has landed; the producer acquires it before the first load of an item of that sub-domain.  nullptr: not gated.
  int* gate;
  // fused step (b2s_halo_fv_tp2d): x.links != nullptr -> every CTA first takes its share of the halo exchange
  // (halo_device.cuh), then walks its stencil items behind the gates; halo update + transport are ONE launch
  HaloXchg x;
  int nb;
};

struct StreamItem {
  int strip, jb0, nrows, k, b, nchunk;
};

template <typename T>
__device__ __forceinline__ StreamItem stream_item(int item, const FvStreamParams<T>& P) {
  StreamItem it;
  const int jblk = item % P.njblk;
  int t = item / P.njblk;
  it.strip = t % P.nstrips;
  t /= P.nstrips;
  it.k = t % P.nk;
  it.b = t / P.nk;
  it.jb0 = P.j0 + jblk * P.jb;
  it.nrows = min(P.jb, P.j1 - it.jb0);
  it.nchunk = (it.nrows + 6 + RB - 1) / RB;
  return it;
}

// MODE: 0 plain stencil, 1 gated, 2 exchange fused in front of the gates (see k_fv_tma.cu).
template <typename T, int TI, int NSTAGE, int MODE>
__global__ void __launch_bounds__(TI + 32) k_fv_stream(const __grid_constant__ CUtensorMap tm_q,
                                                       const __grid_constant__ CUtensorMap tm_crx,
                                                       const __grid_constant__ CUtensorMap tm_xfx,
                                                       const __grid_constant__ CUtensorMap tm_cry,
                                                       const __grid_constant__ CUtensorMap tm_yfx,
                                                       const FvStreamParams<T> P) {
  using G = STile<T, TI, NSTAGE>;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + G::BAR_OFF);
  uint64_t* empty = full + NSTAGE;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int NCONS_WARPS = TI / 32;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 127u) __trap();
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCONS_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if constexpr (MODE == 2) {
    if (P.x.links != nullptr) halo_exchange_call<T>(&P.x, reinterpret_cast<int*>(smem + G::SCRATCH_OFF));
  }

  if (warp == NCONS_WARPS) {
    // =============================== PRODUCER ===============================
    if (lane == 0) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_crx);
      tma_prefetch_desc(&tm_xfx);
      tma_prefetch_desc(&tm_cry);
      tma_prefetch_desc(&tm_yfx);
      int stage = 0;
      uint32_t phase = 0;
      [[maybe_unused]] int b_open = 0;  // sub-domains below b_open have their halos (items run in b order)
      for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        const StreamItem it = stream_item(item, P);
        const int io = it.strip * TI;
        if constexpr (MODE != 0) {
          while (b_open <= it.b) {
            if (blockIdx.x == 0 && b_open == 0) gate_trace(P.gate, 3);
            gate_acquire(P.gate, b_open++);
            if (blockIdx.x == 0 && b_open == 1) gate_trace(P.gate, 4);
          }
        }
        for (int m = 0; m < it.nchunk; ++m) {
          // chunk row rr of chunk m is iteration n = m*RB + rr: q row r = jb0 - 3 + n (tensor row r + 3: the map is
          // based at the halo origin), crx / xfx row r, y-interface r - 2.  Rows outside a tensor are zero-filled.
          const int row = it.jb0 + m * RB;
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* st = smem + stage * G::STAGE_BYTES;
          mbar_arrive_expect_tx(&full[stage], G::TX_BYTES);
          tma_load_4d(st + G::Q_OFF, &tm_q, &full[stage], P.c_q - P.sh_q + io, row, it.k, it.b);
          tma_load_4d(st + G::CRX_OFF, &tm_crx, &full[stage], P.c_crx - P.sh_crx + io, row - 3, it.k, it.b);
          tma_load_4d(st + G::XFX_OFF, &tm_xfx, &full[stage], P.c_xfx - P.sh_xfx + io, row - 3, it.k, it.b);
          tma_load_4d(st + G::CRY_OFF, &tm_cry, &full[stage], P.c_cry - P.sh_cry + io, row - 5, it.k, it.b);
          tma_load_4d(st + G::YFX_OFF, &tm_yfx, &full[stage], P.c_yfx - P.sh_yfx + io, row - 5, it.k, it.b);
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    return;
  }

  // =============================== CONSUMERS ===============================
  const int ci = threadIdx.x;  // column of the strip owned by this thread
  const int oq = G::Q_OFF / (int)sizeof(T) + ci + P.sh_q;
  const int ocx = G::CRX_OFF / (int)sizeof(T) + ci + P.sh_crx;
  const int oxf = G::XFX_OFF / (int)sizeof(T) + ci + P.sh_xfx;
  const int ocy = G::CRY_OFF / (int)sizeof(T) + ci + P.sh_cry;
  const int oyf = G::YFX_OFF / (int)sizeof(T) + ci + P.sh_yfx;
  const int64_t out_sj = P.qout.sj, ra_sj = P.rarea.sj;
  int stage = 0;
  uint32_t phase = 0;
  for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
    const StreamItem it = stream_item(item, P);
    const int i = P.i0 + it.strip * TI + ci;
    const bool col_ok = i < P.i1;
    const int nrows = it.nrows;
    // running pointers of the row stored at iteration n (row jb0 - 6 + n); dereferenced for 6 <= n < 6 + nrows only
    T* outp = P.qout.at(i, it.jb0 - 6, it.k, it.b);
    const T* rap = P.rarea.at(i, it.jb0 - 6, it.b);
    // y-window: q rows r-3 .. r-1, interface values of cells r-3 and r-2, flux x yfx at interface r-3
    T w1 = T(0), w2 = T(0), w3 = T(0), al_a = T(0), al_b = T(0), fy_prev = T(0);
    T dfx[RB];  // x-flux difference of rows r-3 .. r
#pragma unroll
    for (int u = 0; u < RB; ++u) dfx[u] = T(0);
    T ra_cur[RB], ra_nxt[RB];  // rarea of the rows stored by this chunk / the next one
#pragma unroll
    for (int u = 0; u < RB; ++u) ra_cur[u] = T(0);

    for (int m = 0; m < it.nchunk; ++m) {
      // rarea (L2-resident, re-read for every k) of the rows the NEXT chunk stores, in flight across this chunk
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const int n2 = (m + 1) * RB + u;
        ra_nxt[u] = (col_ok && (unsigned)(n2 - 6) < (unsigned)nrows) ? __ldg(rap + (int64_t)(n2)*ra_sj) : T(0);
      }
      mbar_wait(&full[stage], phase);
      const T* st = reinterpret_cast<const T*>(smem + stage * G::STAGE_BYTES);
      const T* qs = st + oq;    // qs[rr*BQ + c]: row r, column i-3+c
      const T* cxs = st + ocx;  // [rr*BX + {0,1}]: x-interfaces i, i+1 of row r
      const T* xfs = st + oxf;
      const T* cys = st + ocy;  // [rr*BY]: y-interface r-2
      const T* yfs = st + oyf;
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
        const int n = m * RB + rr;
        // ---- x direction on row r ----
        const T* row = qs + rr * G::BQ;
        const T xm3 = row[0], xm2 = row[1], xm1 = row[2], qn = row[3], xp1 = row[4], xp2 = row[5], xp3 = row[6];
        const T ax_m1 = ppm_al(xm3, xm2, xm1, qn), ax_0 = ppm_al(xm2, xm1, qn, xp1);
        const T ax_p1 = ppm_al(xm1, qn, xp1, xp2), ax_p2 = ppm_al(qn, xp1, xp2, xp3);
        const T fx_lo = mul_rn(ppm_flux_from_al(xm1, qn, ax_m1, ax_0, ax_p1, cxs[rr * G::BX]), xfs[rr * G::BX]);
        const T fx_hi = mul_rn(ppm_flux_from_al(qn, xp1, ax_0, ax_p1, ax_p2, cxs[rr * G::BX + 1]), xfs[rr * G::BX + 1]);
        const T dfx_out = dfx[(rr + 1) % RB];  // row r-3's, parked three iterations ago
        dfx[rr] = sub_rn(fx_hi, fx_lo);
        // ---- y direction: cell r-1 gets its interface value, interface r-2 its flux ----
        const T al_c = ppm_al(w1, w2, w3, qn);
        const T fy = mul_rn(ppm_flux_from_al(w1, w2, al_a, al_b, al_c, cys[rr * G::BY]), yfs[rr * G::BY]);
        // ---- row r-3 ----
        if (col_ok && (unsigned)(n - 6) < (unsigned)nrows)
          __stcs(outp, fma_rn(-ra_cur[rr], add_rn(dfx_out, sub_rn(fy, fy_prev)), w1));
        outp += out_sj;
        w1 = w2, w2 = w3, w3 = qn;
        al_a = al_b, al_b = al_c;
        fy_prev = fy;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == NSTAGE) {
        stage = 0;
        phase ^= 1;
      }
#pragma unroll
      for (int u = 0; u < RB; ++u) ra_cur[u] = ra_nxt[u];
    }
  }
  if constexpr (MODE != 0) {
    if (threadIdx.x == 0) gate_release(P.gate, P.nb, gridDim.x);  // this CTA has consumed all its loads
  }
}

// Shared-memory opt-in + resident-CTA count of one instance on the current device (cached per device); the first call
// loads the kernel -- see kernel_setup in k_fv_tma.cu for why b2s_halo_init does that up front (fv_stream_preload).
template <typename T, int TI, int NSTAGE>
int stream_kernel_setup(int mode, int* ctas_per_sm) {
  using G = STile<T, TI, NSTAGE>;
  static std::mutex mu;
  static int cache[3][kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  std::lock_guard<std::mutex> lk(mu);
  int& slot = cache[mode][dev];
  if (slot == 0) {
    auto kern = mode == 2 ? k_fv_stream<T, TI, NSTAGE, 2> : (mode == 1 ? k_fv_stream<T, TI, NSTAGE, 1> : k_fv_stream<T, TI, NSTAGE, 0>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
    if (e != cudaSuccess) return set_error((int)e, "fv_tp2d(stream): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    int nblk = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, kern, G::THREADS, G::SMEM_BYTES);
    if (e != cudaSuccess || nblk < 1) return set_error(e == cudaSuccess ? B2S_EUNSUPPORTED : (int)e, "fv_tp2d(stream): occupancy query failed");
    slot = nblk;
  }
  *ctas_per_sm = slot;
  return B2S_OK;
}

template <typename T, int TI, int NSTAGE>
int launch_stream(int ni, int nj, int nk, int nb, int i0, int i1, int j0, int j1, F3<const T> q, F3<const T> crx,
                  F3<const T> xfx, F3<const T> cry, F3<const T> yfx, F2<const T> rarea, F3<T> q_out, cudaStream_t s,
                  bool* applicable, int* gate, const HaloXchg* xchg) {
  using G = STile<T, TI, NSTAGE>;
  constexpr int V = G::V;
  *applicable = false;
  const TmaField<T> fq = tma_field<T>(q.p - 3 - 3 * q.sj, q.sj, q.sk, q.sb, nk, nb);
  const TmaField<T> fcx = tma_field<T>(crx.p, crx.sj, crx.sk, crx.sb, nk, nb);
  const TmaField<T> fxx = tma_field<T>(xfx.p, xfx.sj, xfx.sk, xfx.sb, nk, nb);
  const TmaField<T> fcy = tma_field<T>(cry.p, cry.sj, cry.sk, cry.sb, nk, nb);
  const TmaField<T> fyx = tma_field<T>(yfx.p, yfx.sj, yfx.sk, yfx.sb, nk, nb);
  if (!(fq.ok && fcx.ok && fxx.ok && fcy.ok && fyx.ok)) return B2S_OK;
  CUtensorMap mq, mcx, mxx, mcy, myx;
  const bool ok = make_map<T>(&mq, fq.base, q.sj, q.sk, q.sb, ni + 6 + fq.off, nj + 6, nk, nb, G::BQ, RB) &&
                  make_map<T>(&mcx, fcx.base, crx.sj, crx.sk, crx.sb, ni + 1 + fcx.off, nj, nk, nb, G::BX, RB) &&
                  make_map<T>(&mxx, fxx.base, xfx.sj, xfx.sk, xfx.sb, ni + 1 + fxx.off, nj, nk, nb, G::BX, RB) &&
                  make_map<T>(&mcy, fcy.base, cry.sj, cry.sk, cry.sb, ni + fcy.off, nj + 1, nk, nb, G::BY, RB) &&
                  make_map<T>(&myx, fyx.base, yfx.sj, yfx.sk, yfx.sb, ni + fyx.off, nj + 1, nk, nb, G::BY, RB);
  if (!ok) return B2S_OK;
  const int mode = xchg != nullptr ? 2 : (gate != nullptr ? 1 : 0);  // fused exchange + gates | gates only | plain
  auto kern = mode == 2 ? k_fv_stream<T, TI, NSTAGE, 2> : (mode == 1 ? k_fv_stream<T, TI, NSTAGE, 1> : k_fv_stream<T, TI, NSTAGE, 0>);
  int ctas_per_sm = 0;
  if (int rc = stream_kernel_setup<T, TI, NSTAGE>(mode, &ctas_per_sm)) return rc;
  FvStreamParams<T> P;
  P.nk = nk, P.i0 = i0, P.i1 = i1, P.j0 = j0, P.j1 = j1;
  P.nstrips = (i1 - i0 + TI - 1) / TI;
  // Rows per item and grid size.  Items are long (a whole column height where possible) and the persistent grid
  // hands them out statically, so the split must come out even: for 1 .. 16 row blocks per column, model the
  // busiest SM (waves of items per CTA x CTAs on that SM) and keep the best of
  //     items / (SMs x that)  x  JB / (JB + 42).
  // The 42 is measured, not derived: an item costs its six warm-up rows plus what fits the sweeps as ~36 more
  // (C384x72 fp64: 453 us with 384-row items, 477 with 128, 519 with 64; C720x137 fp32: 1.87 ms with 720-row items,
  // 2.00 with 360, 2.08 with 240), while balance decides when there are few waves (fp32 C384x72, 888 CTAs: one
  // block per column = 1.46 items per CTA, 271 us; two blocks 254 us; three 279 us).
  // b2s_set_option("fv_jb", n) overrides the choice.
  const int h = j1 - j0;
  const int64_t max_ctas = (int64_t)sm_count() * ctas_per_sm;
  auto even_grid = [&](int64_t n) { const int64_t waves = (n + max_ctas - 1) / max_ctas; return (n + waves - 1) / waves; };
  int jb = option("fv_jb", 0);
  if (jb <= 0) {
    double best = -1.0;
    for (int nb_j = 1; nb_j <= 16 && (nb_j == 1 || h / nb_j >= 32); ++nb_j) {
      const int cand = (h + nb_j - 1) / nb_j;
      const int64_t n = (int64_t)P.nstrips * ((h + cand - 1) / cand) * nk * nb;
      const int64_t waves = (n + max_ctas - 1) / max_ctas, g = even_grid(n);
      const int64_t busiest = (g + sm_count() - 1) / sm_count() * waves;
      const double score = (double)n / ((double)sm_count() * busiest) * cand / (cand + 42.0);
      if (score > best) best = score, jb = cand;
    }
  }
  const int nblk_j = (h + jb - 1) / jb;
  P.jb = (h + nblk_j - 1) / nblk_j;
  P.njblk = (h + P.jb - 1) / P.jb;
  const int64_t nitems = (int64_t)P.nstrips * P.njblk * nk * nb;
  if (nitems > (int64_t)1 << 30) return B2S_OK;
  P.nitems = (int)nitems;
  P.c_q = i0 + fq.off, P.sh_q = P.c_q % V;
  P.c_crx = i0 + fcx.off, P.sh_crx = P.c_crx % V;
  P.c_xfx = i0 + fxx.off, P.sh_xfx = P.c_xfx % V;
  P.c_cry = i0 + fcy.off, P.sh_cry = P.c_cry % V;
  P.c_yfx = i0 + fyx.off, P.sh_yfx = P.c_yfx % V;
  P.rarea = rarea;
  P.qout = q_out;
  P.gate = gate;
  if (xchg != nullptr) P.x = *xchg; else P.x.links = nullptr;
  P.nb = nb;
  const int grid = (int)even_grid(nitems);
  *applicable = true;
  kern<<<grid, G::THREADS, G::SMEM_BYTES, s>>>(mq, mcx, mxx, mcy, myx, P);
  return check_launch("fv_tp2d(stream)");
}

}  // namespace

#define B2S_FVS_ARGS ni, nj, nk, nb, i0, i1, j0, j1, q, crx, xfx, cry, yfx, rarea, q_out, s, applicable, gate, xchg

template <typename T>
int fv_tp2d_stream(int ni, int nj, int nk, int nb, int i0, int i1, int j0, int j1, F3<const T> q, F3<const T> crx,
                   F3<const T> xfx, F3<const T> cry, F3<const T> yfx, F2<const T> rarea, F3<T> q_out, cudaStream_t s,
                   bool* applicable, int* gate, const HaloXchg* xchg) {
  // strip width: the one of 128 / 96 / 64 columns that pads the rectangle least, the wider on ties (192-wide
  // sub-domains of the 8-GPU layout: two 96-column strips), 32 for narrow rectangles; ring depth
  // b2s_set_option("fv_stages", 2 | 3 | 4) for the 128- and 64-column kernels
  const int w = i1 - i0;
  int ti = option("fv_ti", 0);
  if (ti != 32 && ti != 64 && ti != 96 && ti != 128) {
    ti = 32;
    if (w > 32) {
      int best_waste = 1 << 30;
      for (int c : {128, 96, 64}) {
        const int waste = (w + c - 1) / c * c - w;
        if (waste < best_waste) best_waste = waste, ti = c;
      }
    }
  }
  int stages = option("fv_stages", 0);
  if (stages < 2 || stages > 4) stages = 3;
  if (ti == 32) return launch_stream<T, 32, 4>(B2S_FVS_ARGS);
  if (ti == 96) return launch_stream<T, 96, 3>(B2S_FVS_ARGS);
  if (ti == 64) return stages == 2 ? launch_stream<T, 64, 2>(B2S_FVS_ARGS) : (stages == 4 ? launch_stream<T, 64, 4>(B2S_FVS_ARGS) : launch_stream<T, 64, 3>(B2S_FVS_ARGS));
  return stages == 2 ? launch_stream<T, 128, 2>(B2S_FVS_ARGS) : (stages == 4 ? launch_stream<T, 128, 4>(B2S_FVS_ARGS) : launch_stream<T, 128, 3>(B2S_FVS_ARGS));
}

template <typename T>
static int stream_preload_all() {
  int n = 0, rc = 0;
  for (int gated = 0; gated < 3; ++gated) {  // mode: plain, gated, fused
    if ((rc = stream_kernel_setup<T, 32, 4>(gated, &n))) return rc;
    if ((rc = stream_kernel_setup<T, 96, 3>(gated, &n))) return rc;
    if ((rc = stream_kernel_setup<T, 64, 2>(gated, &n))) return rc;
    if ((rc = stream_kernel_setup<T, 64, 3>(gated, &n))) return rc;
    if ((rc = stream_kernel_setup<T, 64, 4>(gated, &n))) return rc;
    if ((rc = stream_kernel_setup<T, 128, 2>(gated, &n))) return rc;
    if ((rc = stream_kernel_setup<T, 128, 3>(gated, &n))) return rc;
    if ((rc = stream_kernel_setup<T, 128, 4>(gated, &n))) return rc;
  }
  return B2S_OK;
}
// every instance fv_tp2d_stream can dispatch to, loaded and set up on the current device
int fv_stream_preload() {
  if (int rc = stream_preload_all<double>()) return rc;
  return stream_preload_all<float>();
}

template int fv_tp2d_stream<double>(int, int, int, int, int, int, int, int, F3<const double>, F3<const double>,
                                    F3<const double>, F3<const double>, F3<const double>, F2<const double>, F3<double>,
                                    cudaStream_t, bool*, int*, const HaloXchg*);
template int fv_tp2d_stream<float>(int, int, int, int, int, int, int, int, F3<const float>, F3<const float>,
                                   F3<const float>, F3<const float>, F3<const float>, F2<const float>, F3<float>,
                                   cudaStream_t, bool*, int*, const HaloXchg*);

}  // namespace impl
}  // namespace b2s
