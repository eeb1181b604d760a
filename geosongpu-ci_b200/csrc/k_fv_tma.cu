// K5 (variant 2, TMA tiles): fv_tp2d as a persistent, warp-specialised sm_100a kernel.  Default only for launches
// below ~12 M points since k_fv_stream.cu (variant 3) took over; kept bit-identical to it.
// Spec: oracle/numpy_oracle.py fv_tp2d (no source in /root/reference).
//
// Design (DESIGN.md "K5"):
//   * work item = one TI x R tile of one (k, b) level; a persistent grid (CTAs/SM x 148 SMs)
//     walks the item list round-robin, so there is no tail wave and no per-tile launch cost;
//   * warp 4 is the PRODUCER: one elected lane issues five 4-D TMA tile loads per item
//     (cp.async.bulk.tensor.4d -> SASS UTMALDG): q with its 3-cell apron ((TI+6) x (R+6)),
//     crx/xfx ((TI+1) x R) and cry/yfx (TI x (R+1)), completing on an mbarrier (expect_tx);
//   * warps 0-3 are CONSUMERS: thread = one i-column of the tile, marching the R rows with the
//     y-direction window (7 q values, 3 interface values, the low-side flux) carried in
//     registers; x-direction neighbours come from the shared-memory tile (conflict-free:
//     consecutive lanes read consecutive elements).  Results go straight to HBM (coalesced,
//     streaming stores);
//   * NSTAGE-deep full/empty mbarrier ring: loads of the next items are in flight while the
//     current one is computed, independent of occupancy;
//   * every input element crosses HBM once; apron re-reads (q: (R+6)/R, cry/yfx: (R+1)/R) are
//     L2 hits because neighbouring tiles are consecutive in the item order.
// Algorithmic bytes/point: 40 R + 8 W + 8/nk (same as variant 1).
#include <mutex>

#include "fv_math.cuh"
#include "halo_device.cuh"
#include "impl.cuh"
#include "tma.cuh"

namespace b2s {
namespace impl {

namespace {

constexpr int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---- compile-time tile geometry -------------------------------------------------------------------

template <typename T, int TI, int R, int NSTAGE>
struct Tile {
  static constexpr int V = 16 / sizeof(T);          // elements per 16 B
  // A TMA box must START on a 16-byte boundary of global memory (an odd fp64 start coordinate
  // faults with "illegal instruction"), so every box begins at the aligned column at or before
  // the one needed and is V-1 elements wider; the consumer skips `shift` leading elements.
  static constexpr int BQ = round_up(TI + 6 + V - 1, V);  // q box: columns i_s-3 .. i_s+TI+2
  static constexpr int BX = round_up(TI + 1 + V - 1, V);  // crx/xfx box: interfaces i_s .. i_s+TI
  static constexpr int BY = round_up(TI + V - 1, V);      // cry/yfx box
  static constexpr int RQ = R + 6, RX = R, RY = R + 1;
  static constexpr int Q_BYTES = RQ * BQ * sizeof(T);
  static constexpr int X_BYTES = RX * BX * sizeof(T);
  static constexpr int Y_BYTES = RY * BY * sizeof(T);
  static constexpr int Q_OFF = 0;
  static constexpr int CRX_OFF = round_up(Q_BYTES, 128);
  static constexpr int XFX_OFF = CRX_OFF + round_up(X_BYTES, 128);
  static constexpr int CRY_OFF = XFX_OFF + round_up(X_BYTES, 128);
  static constexpr int YFX_OFF = CRY_OFF + round_up(Y_BYTES, 128);
  static constexpr int STAGE_BYTES = YFX_OFF + round_up(Y_BYTES, 128);
  static constexpr int TX_BYTES = Q_BYTES + 2 * X_BYTES + 2 * Y_BYTES;
  static constexpr int BAR_OFF = NSTAGE * STAGE_BYTES;
  static constexpr int SCRATCH_OFF = BAR_OFF + 2 * NSTAGE * 8;  // one int for the fused exchange phase
  static constexpr int SMEM_BYTES = SCRATCH_OFF + 16;
  static constexpr int THREADS = TI + 32;
  static_assert(BQ <= 256 && BX <= 256 && RQ <= 256, "TMA box dimensions are limited to 256 elements");
};

template <typename T>
struct FvTmaParams {
  int nk, nb, i0, i1, j0, j1;
  int nstrips, njblk;
  int nitems;
  // per field: c = TMA coordinate of compute column i0 (tensor base is 16-byte aligned), sh = c % V.
  // The box of a tile starts at c - sh + strip*TI; the tile's first needed element sits sh further.
  int c_q, c_crx, c_xfx, c_cry, c_yfx;
  int sh_q, sh_crx, sh_xfx, sh_cry, sh_yfx;
  F2<const T> rarea;
  F3<T> qout;
  // gated launch (fv_tp2d_gated): gate[b] is raised by the halo exchange kernel when every halo cell of sub-domain b
  // has landed; the producer acquires it before the first load of an item of that sub-domain.  nullptr: not gated.
  int* gate;
  // fused step (b2s_halo_fv_tp2d): x.links != nullptr -> every CTA first takes its share of the halo exchange
  // (halo_device.cuh), then walks its stencil items behind the gates; halo update + transport are ONE launch
  HaloXchg x;
};

// Position of a work item, advanced by gridDim.x items per step without divisions:
// item -> (jb fastest, strip, k, b).
struct ItemCursor {
  int jb, strip, k, b;
  int d_jb, d_strip, d_k, d_b;
  __device__ __forceinline__ void init(int item, int step, int njblk, int nstrips, int nk) {
    jb = item % njblk;
    int t = item / njblk;
    strip = t % nstrips;
    t /= nstrips;
    k = t % nk;
    b = t / nk;
    d_jb = step % njblk;
    t = step / njblk;
    d_strip = t % nstrips;
    t /= nstrips;
    d_k = t % nk;
    d_b = t / nk;
  }
  __device__ __forceinline__ void advance(int njblk, int nstrips, int nk) {
    jb += d_jb;
    if (jb >= njblk) {
      jb -= njblk;
      strip += 1;
    }
    strip += d_strip;
    if (strip >= nstrips) {
      strip -= nstrips;
      k += 1;
    }
    k += d_k;
    if (k >= nk) {
      k -= nk;
      b += 1;
    }
    b += d_b;
  }
};

// MODE 0 is the plain stencil, 1 adds the gates (fv_tp2d_gated), 2 the exchange in front of them (halo_fv_tp2d: the call
// into halo_exchange_call and its stack frame).  Mode 0 contains none of the gate / exchange code, so its register allocation and
// schedule are exactly those of the kernel without the multi-GPU machinery (with the code merely switched off at run
// time the consumers' loop came out 8 % slower at 3 x 192 x 192 x 72: ptxas re-allocated it, 91 -> 79 registers).
template <typename T, int TI, int R, int NSTAGE, int MODE>
__global__ void __launch_bounds__(TI + 32) k_fv_tma(const __grid_constant__ CUtensorMap tm_q,
                                                    const __grid_constant__ CUtensorMap tm_crx,
                                                    const __grid_constant__ CUtensorMap tm_xfx,
                                                    const __grid_constant__ CUtensorMap tm_cry,
                                                    const __grid_constant__ CUtensorMap tm_yfx,
                                                    const FvTmaParams<T> P) {
  using G = Tile<T, TI, R, NSTAGE>;
  // Indexed straight off the __shared__ symbol so the compiler keeps the address space (LDS with
  // immediate offsets, not generic LD).  TMA needs 128-byte aligned destinations: the dynamic
  // shared window of a CTA without static shared memory starts 1024-byte aligned.
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + G::BAR_OFF);
  uint64_t* empty = full + NSTAGE;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int NCONS_WARPS = TI / 32;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 127u) __trap();
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);             // producer's arrive.expect_tx
      mbar_init(&empty[s], NCONS_WARPS);  // one arrive per consumer warp
    }
    fence_barrier_init();
  }
  __syncthreads();
  if constexpr (MODE == 2) {
    if (P.x.links != nullptr) halo_exchange_call<T>(&P.x, reinterpret_cast<int*>(smem + G::SCRATCH_OFF));
  }

  ItemCursor cur;
  cur.init(blockIdx.x, gridDim.x, P.njblk, P.nstrips, P.nk);
  const int nmine = (P.nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

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
      // one item: wait for a free stage, five tile loads
      auto issue = [&]() {
        const int io = cur.strip * TI;     // column offset of the tile inside the rectangle
        const int js = P.j0 + cur.jb * R;  // first compute row of the tile
        mbar_wait(&empty[stage], phase ^ 1);
        unsigned char* st = smem + stage * G::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[stage], G::TX_BYTES);
        // q's tensor map is based at the halo origin (-3,-3): compute cell i-3 is coordinate i, row j-3 is row j
        tma_load_4d(st + G::Q_OFF, &tm_q, &full[stage], P.c_q - P.sh_q + io, js, cur.k, cur.b);
        tma_load_4d(st + G::CRX_OFF, &tm_crx, &full[stage], P.c_crx - P.sh_crx + io, js, cur.k, cur.b);
        tma_load_4d(st + G::XFX_OFF, &tm_xfx, &full[stage], P.c_xfx - P.sh_xfx + io, js, cur.k, cur.b);
        tma_load_4d(st + G::CRY_OFF, &tm_cry, &full[stage], P.c_cry - P.sh_cry + io, js, cur.k, cur.b);
        tma_load_4d(st + G::YFX_OFF, &tm_yfx, &full[stage], P.c_yfx - P.sh_yfx + io, js, cur.k, cur.b);
        cur.advance(P.njblk, P.nstrips, P.nk);
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      };
      if constexpr (MODE == 0) {
        for (int n = 0; n < nmine; ++n) issue();
      } else {
        // Items run in sub-domain order: the gate of sub-domain b is acquired ONCE, in front of this CTA's run of
        // items of b, and the run itself is the plain loop (a gate test inside the item loop kept ptxas from
        // unrolling it and made the gated kernel ~5 % slower than the plain one).
        const int per_b = P.nitems / P.nb;
        int n = 0;
        for (int b = 0; b < P.nb && n < nmine; ++b) {
          const int end_item = (b + 1) * per_b;  // my items below it: blockIdx.x + m * gridDim.x < end_item
          int n_end = end_item > (int)blockIdx.x ? (end_item - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
          n_end = min(n_end, nmine);
          if (n_end <= n) continue;
          if (blockIdx.x == 0 && b == 0) gate_trace(P.gate, 3);
          gate_acquire(P.gate, b);
          if (blockIdx.x == 0 && b == 0) gate_trace(P.gate, 4);
          for (; n < n_end; ++n) issue();
        }
      }
    }
    return;
  }

  // =============================== CONSUMERS ===============================
  const int ci = threadIdx.x;  // column of the tile owned by this thread
  // per-thread element offsets into a stage (constant for the whole kernel)
  const int oq = G::Q_OFF / (int)sizeof(T) + ci + P.sh_q;
  const int ocx = G::CRX_OFF / (int)sizeof(T) + ci + P.sh_crx;
  const int oxf = G::XFX_OFF / (int)sizeof(T) + ci + P.sh_xfx;
  const int ocy = G::CRY_OFF / (int)sizeof(T) + ci + P.sh_cry;
  const int oyf = G::YFX_OFF / (int)sizeof(T) + ci + P.sh_yfx;
  int stage = 0;
  uint32_t phase = 0;
  for (int n = 0; n < nmine; ++n) {
    const int i = P.i0 + cur.strip * TI + ci;
    const int js = P.j0 + cur.jb * R;
    const int nrows = i < P.i1 ? min(R, P.j1 - js) : 0;  // rows of this tile this thread stores
    const T* rap = P.rarea.at(i, js, cur.b);
    T* outp = P.qout.at(i, js, cur.k, cur.b);
    cur.advance(P.njblk, P.nstrips, P.nk);

    // rarea does not go through shared memory: R independent loads issued before the wait
    T ra[R];
#pragma unroll
    for (int r = 0; r < R; ++r) ra[r] = r < nrows ? __ldg(rap + (int64_t)r * P.rarea.sj) : T(0);

    mbar_wait(&full[stage], phase);
    const T* st = reinterpret_cast<const T*>(smem + stage * G::STAGE_BYTES);
    const T* qs = st + oq;    // qs[r*BQ + c]: row js-3+r, column i-3+c
    const T* cxs = st + ocx;  // [r*BX + {0,1}]: interfaces i, i+1
    const T* xfs = st + oxf;
    const T* cys = st + ocy;  // [r*BY]: interface js+r
    const T* yfs = st + oyf;

    // y-direction window centred on row js (q0), carried down the tile
    T qm2 = qs[1 * G::BQ + 3], qm1 = qs[2 * G::BQ + 3], q0 = qs[3 * G::BQ + 3];
    T qp1 = qs[4 * G::BQ + 3], qp2 = qs[5 * G::BQ + 3];
    T al_0 = ppm_al(qm2, qm1, q0, qp1);
    T al_p1 = ppm_al(qm1, q0, qp1, qp2);
    T fy_lo;
    {
      const T al_m1 = ppm_al(qs[0 * G::BQ + 3], qm2, qm1, q0);
      fy_lo = mul_rn(ppm_flux_from_al(qm1, q0, al_m1, al_0, al_p1, cys[0]), yfs[0]);
    }

#pragma unroll
    for (int r = 0; r < R; ++r) {
      const T qp3 = qs[(r + 6) * G::BQ + 3];
      const T al_p2 = ppm_al(q0, qp1, qp2, qp3);
      const T fy_hi = mul_rn(ppm_flux_from_al(q0, qp1, al_0, al_p1, al_p2, cys[(r + 1) * G::BY]), yfs[(r + 1) * G::BY]);
      // x direction on row js + r (shared-memory row r + 3)
      const T* row = qs + (r + 3) * G::BQ;
      const T xm3 = row[0], xm2 = row[1], xm1 = row[2], xp1 = row[4], xp2 = row[5], xp3 = row[6];
      const T ax_m1 = ppm_al(xm3, xm2, xm1, q0), ax_0 = ppm_al(xm2, xm1, q0, xp1);
      const T ax_p1 = ppm_al(xm1, q0, xp1, xp2), ax_p2 = ppm_al(q0, xp1, xp2, xp3);
      const T fx_lo = mul_rn(ppm_flux_from_al(xm1, q0, ax_m1, ax_0, ax_p1, cxs[r * G::BX]), xfs[r * G::BX]);
      const T fx_hi = mul_rn(ppm_flux_from_al(q0, xp1, ax_0, ax_p1, ax_p2, cxs[r * G::BX + 1]), xfs[r * G::BX + 1]);
      if (r < nrows) __stcs(outp + (int64_t)r * P.qout.sj, fv_update(q0, ra[r], fx_lo, fx_hi, fy_lo, fy_hi));
      // slide the window one row down
      qm1 = q0;
      q0 = qp1;
      qp1 = qp2;
      qp2 = qp3;
      al_0 = al_p1;
      al_p1 = al_p2;
      fy_lo = fy_hi;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    if (++stage == NSTAGE) {
      stage = 0;
      phase ^= 1;
    }
  }
  if constexpr (MODE != 0) {
    if (threadIdx.x == 0) gate_release(P.gate, P.nb, gridDim.x);  // this CTA has consumed all its loads
  }
}

// ---- host side (tensor maps: tma.cuh / tma_host.cu) ------------------------------------------------

// Shared-memory opt-in and the resident-CTA count of one kernel instance on the CURRENT device, cached per device.
// The first call also LOADS the kernel (CUDA loads modules lazily, and loading may have to wait for the context to go
// idle): b2s_halo_init runs it for every instance (fv_tma_preload below) so that no load can happen later, while an
// exchange kernel of this context is spinning on a neighbour that cannot launch until the load is over.
template <typename T, int TI, int R, int NSTAGE>
int kernel_setup(int mode, int* ctas_per_sm) {
  using G = Tile<T, TI, R, NSTAGE>;
  static std::mutex mu;
  static int cache[3][kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  std::lock_guard<std::mutex> lk(mu);
  int& slot = cache[mode][dev];
  if (slot == 0) {
    auto kern = mode == 2 ? k_fv_tma<T, TI, R, NSTAGE, 2> : (mode == 1 ? k_fv_tma<T, TI, R, NSTAGE, 1> : k_fv_tma<T, TI, R, NSTAGE, 0>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
    if (e != cudaSuccess) return set_error((int)e, "fv_tp2d(tma): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    int nblk = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, kern, G::THREADS, G::SMEM_BYTES);
    if (e != cudaSuccess || nblk < 1) return set_error(e == cudaSuccess ? B2S_EUNSUPPORTED : (int)e, "fv_tp2d(tma): occupancy query failed");
    slot = nblk;
  }
  *ctas_per_sm = slot;
  return B2S_OK;
}

template <typename T, int TI, int R, int NSTAGE>
int launch(int ni, int nj, int nk, int nb, int i0, int i1, int j0, int j1, F3<const T> q, F3<const T> crx,
           F3<const T> xfx, F3<const T> cry, F3<const T> yfx, F2<const T> rarea, F3<T> q_out, cudaStream_t s,
           bool* applicable, int* gate, const HaloXchg* xchg) {
  using G = Tile<T, TI, R, NSTAGE>;
  *applicable = false;
  // q's tensor starts at the halo origin (-3, -3)
  const TmaField<T> fq = tma_field<T>(q.p - 3 - 3 * q.sj, q.sj, q.sk, q.sb, nk, nb);
  const TmaField<T> fcx = tma_field<T>(crx.p, crx.sj, crx.sk, crx.sb, nk, nb);
  const TmaField<T> fxx = tma_field<T>(xfx.p, xfx.sj, xfx.sk, xfx.sb, nk, nb);
  const TmaField<T> fcy = tma_field<T>(cry.p, cry.sj, cry.sk, cry.sb, nk, nb);
  const TmaField<T> fyx = tma_field<T>(yfx.p, yfx.sj, yfx.sk, yfx.sb, nk, nb);
  if (!(fq.ok && fcx.ok && fxx.ok && fcy.ok && fyx.ok)) return B2S_OK;

  CUtensorMap mq, mcx, mxx, mcy, myx;
  bool ok = make_map<T>(&mq, fq.base, q.sj, q.sk, q.sb, ni + 6 + fq.off, nj + 6, nk, nb, G::BQ, G::RQ) &&
            make_map<T>(&mcx, fcx.base, crx.sj, crx.sk, crx.sb, ni + 1 + fcx.off, nj, nk, nb, G::BX, G::RX) &&
            make_map<T>(&mxx, fxx.base, xfx.sj, xfx.sk, xfx.sb, ni + 1 + fxx.off, nj, nk, nb, G::BX, G::RX) &&
            make_map<T>(&mcy, fcy.base, cry.sj, cry.sk, cry.sb, ni + fcy.off, nj + 1, nk, nb, G::BY, G::RY) &&
            make_map<T>(&myx, fyx.base, yfx.sj, yfx.sk, yfx.sb, ni + fyx.off, nj + 1, nk, nb, G::BY, G::RY);
  if (!ok) return B2S_OK;  // driver refused the descriptor: let the direct kernel handle the call

  const int mode = xchg != nullptr ? 2 : (gate != nullptr ? 1 : 0);  // fused exchange + gates | gates only | plain
  auto kern = mode == 2 ? k_fv_tma<T, TI, R, NSTAGE, 2> : (mode == 1 ? k_fv_tma<T, TI, R, NSTAGE, 1> : k_fv_tma<T, TI, R, NSTAGE, 0>);
  int ctas_per_sm = 0;
  if (int rc = kernel_setup<T, TI, R, NSTAGE>(mode, &ctas_per_sm)) return rc;
  FvTmaParams<T> P;
  P.nk = nk;
  P.nb = nb;
  P.i0 = i0;
  P.i1 = i1;
  P.j0 = j0;
  P.j1 = j1;
  P.nstrips = (i1 - i0 + TI - 1) / TI;
  P.njblk = (j1 - j0 + R - 1) / R;
  const int64_t nitems = (int64_t)P.nstrips * P.njblk * nk * nb;
  if (nitems > (int64_t)1 << 30) return B2S_OK;  // beyond the 32-bit item cursor: direct kernel
  P.nitems = (int)nitems;
  constexpr int V = G::V;
  static_assert(TI % V == 0, "tile width must keep the box start alignment from strip to strip");
  P.c_q = i0 + fq.off, P.sh_q = P.c_q % V;
  P.c_crx = i0 + fcx.off, P.sh_crx = P.c_crx % V;
  P.c_xfx = i0 + fxx.off, P.sh_xfx = P.c_xfx % V;
  P.c_cry = i0 + fcy.off, P.sh_cry = P.c_cry % V;
  P.c_yfx = i0 + fyx.off, P.sh_yfx = P.c_yfx % V;
  P.rarea = rarea;
  P.qout = q_out;
  P.gate = gate;
  if (xchg != nullptr) P.x = *xchg; else P.x.links = nullptr;
  const int64_t max_ctas = (int64_t)sm_count() * ctas_per_sm;
  const int grid = (int)(nitems < max_ctas ? nitems : max_ctas);
  *applicable = true;
  kern<<<grid, G::THREADS, G::SMEM_BYTES, s>>>(mq, mcx, mxx, mcy, myx, P);
  return check_launch("fv_tp2d(tma)");
}

}  // namespace

#define B2S_FV_ARGS ni, nj, nk, nb, i0, i1, j0, j1, q, crx, xfx, cry, yfx, rarea, q_out, s, applicable, gate, xchg

template <typename T, int TI>
int launch_rows_stages(int rows, int stages, int ni, int nj, int nk, int nb, int i0, int i1, int j0, int j1,
                       F3<const T> q, F3<const T> crx, F3<const T> xfx, F3<const T> cry, F3<const T> yfx,
                       F2<const T> rarea, F3<T> q_out, cudaStream_t s, bool* applicable, int* gate, const HaloXchg* xchg) {
  if (rows == 8) return stages == 3 ? launch<T, TI, 8, 3>(B2S_FV_ARGS) : launch<T, TI, 8, 2>(B2S_FV_ARGS);
  return stages == 3 ? launch<T, TI, 4, 3>(B2S_FV_ARGS) : launch<T, TI, 4, 2>(B2S_FV_ARGS);
}

template <typename T>
int fv_tp2d_tma(int ni, int nj, int nk, int nb, int i0, int i1, int j0, int j1, F3<const T> q, F3<const T> crx,
                F3<const T> xfx, F3<const T> cry, F3<const T> yfx, F2<const T> rarea, F3<T> q_out, cudaStream_t s,
                bool* applicable, int* gate, const HaloXchg* xchg) {
  // Tile geometry.  Width TI = consumer threads per CTA (+ 1 producer warp); widths on offer: 32
  // (narrow boundary strips), 64, 96, 128, 192 columns; R rows per stage; NSTAGE-deep ring.
  // Automatic choice, from the sweeps in profiles/r01_fv_tile_sweep.md:
  //   width a multiple of 128 (C384 tiles, 384-wide sub-domains): 128 x 4 rows (fp64) / 8 rows (fp32)
  //   width a multiple of 64 only (the 192-wide sub-domains of the 8-GPU layout): 64 x 8 rows --
  //     no half-empty tile, 4 CTAs/SM; 6.2 TB/s vs 5.2 TB/s for 128-column tiles on 3x192x192x72
  //   anything else: the width among 128/96/64 that wastes the fewest columns, ties towards 128.
  // b2s_set_option("fv_ti" | "fv_rows" | "fv_stages", n) overrides each for tuning.
  const int w = i1 - i0;
  int ti = option("fv_ti", 0);
  int rows = option("fv_rows", 0), stages = option("fv_stages", 0);
  if (ti == 0) {
    if (w <= 32) {
      ti = 32;
    } else if (w % 128 == 0) {
      ti = 128;
    } else if (w % 64 == 0) {
      ti = 64;
      if (rows == 0) rows = 8;
    } else {
      const int cand[3] = {128, 96, 64};
      int best_waste = 1 << 30;
      for (int c : cand) {
        const int waste = (w + c - 1) / c * c - w;
        if (waste < best_waste) {
          best_waste = waste;
          ti = c;
        }
      }
    }
  }
  // fp64: 4-row stages (3 CTAs/SM) win on C384-sized sub-domains, 8-row stages (apron re-read 1.75x
  // instead of 2.5x) on anything larger than ~512^2 (C720x137: 3.85 ms vs 4.17 ms); fp32: always 8.
  if (rows == 0) rows = (sizeof(T) == 8 && (int64_t)w * (j1 - j0) <= 512 * 512) ? 4 : 8;
  if (stages == 0) stages = 2;
  switch (ti) {
    case 32:
      return launch<T, 32, 8, 3>(B2S_FV_ARGS);
    case 64:
      return launch_rows_stages<T, 64>(rows, stages, B2S_FV_ARGS);
    case 96:
      return launch_rows_stages<T, 96>(rows, stages, B2S_FV_ARGS);
    case 192:
      return launch_rows_stages<T, 192>(rows, stages, B2S_FV_ARGS);
    default:
      return launch_rows_stages<T, 128>(rows, stages, B2S_FV_ARGS);
  }
}

template <typename T, int TI>
static int preload_rows_stages() {
  int n = 0, rc = 0;
  for (int g = 0; g < 3; ++g) {
    if ((rc = kernel_setup<T, TI, 4, 2>(g, &n))) return rc;
    if ((rc = kernel_setup<T, TI, 4, 3>(g, &n))) return rc;
    if ((rc = kernel_setup<T, TI, 8, 2>(g, &n))) return rc;
    if ((rc = kernel_setup<T, TI, 8, 3>(g, &n))) return rc;
  }
  return B2S_OK;
}
template <typename T>
static int preload_all() {
  int n = 0, rc = 0;
  for (int g = 0; g < 3; ++g)
    if ((rc = kernel_setup<T, 32, 8, 3>(g, &n))) return rc;
  if ((rc = preload_rows_stages<T, 64>())) return rc;
  if ((rc = preload_rows_stages<T, 96>())) return rc;
  if ((rc = preload_rows_stages<T, 128>())) return rc;
  return preload_rows_stages<T, 192>();
}
// every instance fv_tp2d_tma can dispatch to, loaded and set up on the current device
int fv_tma_preload() {
  if (int rc = preload_all<double>()) return rc;
  return preload_all<float>();
}

template int fv_tp2d_tma<double>(int, int, int, int, int, int, int, int, F3<const double>, F3<const double>,
                                 F3<const double>, F3<const double>, F3<const double>, F2<const double>,
                                 F3<double>, cudaStream_t, bool*, int*, const HaloXchg*);
template int fv_tp2d_tma<float>(int, int, int, int, int, int, int, int, F3<const float>, F3<const float>,
                                F3<const float>, F3<const float>, F3<const float>, F2<const float>, F3<float>,
                                cudaStream_t, bool*, int*, const HaloXchg*);

}  // namespace impl
}  // namespace b2s
