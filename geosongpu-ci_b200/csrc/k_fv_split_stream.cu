// K5b (streaming variant): fv_tp2d_split marching down j, every row of every sweep computed ONCE.
// Spec: oracle/numpy_oracle.py fv_tp2d_split (see k_fv_split.cu for the scheme and its provenance).
//
// The tile kernel (k_fv_split.cu) recomputes the inner sweeps on the 3-row aprons of every R-row tile
// ((R+6)/R times the x-sweep work) and is issue-bound.  Here a CTA owns a strip of tw <= TI columns of one
// (k, b) level and a block of JB rows, and marches through the rows r = j0-3 .. j0+JB+2 once:
//   * inputs arrive in 4-row chunks through a two-stage TMA ring (q and area with the i-apron, crx/xfx,
//     and cry/yfx shifted by two rows so that a chunk row holds the y-interface it is used for);
//   * thread t is tile column t (compute column i0-3+t) in the y-direction phases and x-interface i0+t
//     in the x-direction phases.  Everything that is reused down the column lives in registers: the six-row
//     q window and the six-row q_j window of the two y-sweeps, the previous interface's fluxes, and
//     four-deep rings of the x-interface's crx / xfx / fx2 and of the cell areas (row r-3 is needed when
//     row r arrives) -- indexed by the chunk row under full unrolling, so they stay registers;
//   * three quantities cross threads and go through double-buffered shared-memory rows: fx2 of row r (for
//     q_j), q_i of row r-3 (for the outer x-sweep) and the averaged x-flux of row r-3 (for the update);
//     two barriers per row.
// Per row r:   1. fx2(r)            2. fy2(r-2), q_i(r-3)          -- barrier --
//              3. q_j(r), fy(r-2)   4. fx(r-3)                     -- barrier --   5. q_out(r-3)
// Same formulas and the same explicit-rounding arithmetic as the tile kernel: the two variants produce
// identical bits (tests assert it).  Cube corners: as in the tile kernel (`corner_flags`).
#include "fv_math.cuh"
#include "impl.cuh"
#include "tma.cuh"

namespace b2s {
namespace impl {

namespace {

constexpr int round_up_(int x, int m) { return (x + m - 1) / m * m; }

__device__ __forceinline__ double rcp_fast_(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = __fma_rn(-x, r, 1.0);
  r = __fma_rn(r, e, r);
  e = __fma_rn(-x, r, 1.0);
  return __fma_rn(r, e, r);
}
__device__ __forceinline__ float rcp_fast_(float x) { return __frcp_rn(x); }

constexpr int RB = 4;  // rows per chunk = depth of the register rings

template <typename T, int TI>
struct StreamTile {
  static constexpr int V = 16 / sizeof(T);
  static constexpr int WQ = round_up_(TI + 6 + V - 1, V);
  static constexpr int WX = round_up_(TI + 1 + V - 1, V);
  static constexpr int Q_OFF = 0;
  static constexpr int AR_OFF = Q_OFF + round_up_(RB * WQ * (int)sizeof(T), 128);
  static constexpr int CX_OFF = AR_OFF + round_up_(RB * WQ * (int)sizeof(T), 128);
  static constexpr int XF_OFF = CX_OFF + round_up_(RB * WX * (int)sizeof(T), 128);
  static constexpr int CY_OFF = XF_OFF + round_up_(RB * WX * (int)sizeof(T), 128);
  static constexpr int YF_OFF = CY_OFF + round_up_(RB * WQ * (int)sizeof(T), 128);
  static constexpr int STAGE_BYTES = YF_OFF + round_up_(RB * WQ * (int)sizeof(T), 128);
  static constexpr int TX_BYTES = RB * (4 * WQ + 2 * WX) * (int)sizeof(T);
  static constexpr int P = TI + 8;                                      // pitch of the exchange rows
  static constexpr int XROWS_OFF = 2 * STAGE_BYTES;                     // fx2[2][P], q_i[2][P], fxa[2][P]
  static constexpr int BAR_OFF = round_up_(XROWS_OFF + 6 * P * (int)sizeof(T), 16);
  static constexpr int SMEM_BYTES = BAR_OFF + 16;
  static constexpr int THREADS = round_up_(TI + 6, 32);
};

template <typename T>
struct StreamParams {
  int ni, nj, nk, nstrips, njblk, tw, jb;  // jb = rows per CTA
  int c_q, c_ar, c_cx, c_xf, c_cy, c_yf;   // TMA column coordinate of strip 0 (aligned down)
  int s_q, s_ar, s_cx, s_xf, s_cy, s_yf;   // elements to skip inside a box row
  F2<const T> rarea;
  F3<const T> q;
  const int* corner_flags;
  F3<T> qout, fxo, fyo;
};

template <typename T, int TI>
__global__ void __launch_bounds__(StreamTile<T, TI>::THREADS) k_fv_split_stream(
    const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_ar,
    const __grid_constant__ CUtensorMap tm_cx, const __grid_constant__ CUtensorMap tm_xf,
    const __grid_constant__ CUtensorMap tm_cy, const __grid_constant__ CUtensorMap tm_yf, const StreamParams<T> P) {
  using G = StreamTile<T, TI>;
  constexpr int WQ = G::WQ, WX = G::WX, PP = G::P;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + G::BAR_OFF);  // [2]
  T* xrows = reinterpret_cast<T*>(smem + G::XROWS_OFF);
  T* fx2row = xrows;            // [2][P] by row parity
  T* qirow = xrows + 2 * PP;    // [2][P]
  T* fxarow = xrows + 4 * PP;   // [2][P]
  const int tid = threadIdx.x;

  int t = blockIdx.x;
  const int jblk = t % P.njblk;
  t /= P.njblk;
  const int strip = t % P.nstrips;
  t /= P.nstrips;
  const int k = t % P.nk;
  const int b = t / P.nk;
  const int tw = P.tw;
  const int i0 = strip * tw;
  const int jc0 = jblk * P.jb, jc1 = min(P.nj, jc0 + P.jb);  // output rows [jc0, jc1)
  const int r0 = jc0 - 3;                                    // first row that enters the windows
  const int niter = (jc1 - jc0) + 6;                         // rows r0 .. jc1+2
  const int nchunk = (niter + RB - 1) / RB;

  auto issue = [&](int m) {  // chunk m -> stage m & 1 (one thread)
    unsigned char* st = smem + (m & 1) * G::STAGE_BYTES;
    uint64_t* bar = &full[m & 1];
    const int row = r0 + RB * m + 3;  // tensor row of compute row r0 + RB*m (tensors start at the halo origin, row -3)
    mbar_arrive_expect_tx(bar, G::TX_BYTES);
    tma_load_4d(st + G::Q_OFF, &tm_q, bar, P.c_q + i0, row, k, b);
    tma_load_4d(st + G::AR_OFF, &tm_ar, bar, P.c_ar + i0, row, 0, b);
    tma_load_4d(st + G::CX_OFF, &tm_cx, bar, P.c_cx + i0, row, k, b);
    tma_load_4d(st + G::XF_OFF, &tm_xf, bar, P.c_xf + i0, row, k, b);
    // y-interfaces: chunk row rr of iteration row r holds interface r - 2 (tensor row = interface index)
    tma_load_4d(st + G::CY_OFF, &tm_cy, bar, P.c_cy + i0, row - 3 - 2, k, b);
    tma_load_4d(st + G::YF_OFF, &tm_yf, bar, P.c_yf + i0, row - 3 - 2, k, b);
  };
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_barrier_init();
    issue(0);
    if (nchunk > 1) issue(1);
  }
  __syncthreads();

  // roles of this thread
  const int c = tid;                   // tile column (compute column i0 - 3 + c)
  const bool ycol = c < tw + 6;        // takes part in the y-direction phases
  const bool xint = c <= tw;           // x-interface i0 + c
  const int cc = c - 3;                // compute column of the tile, valid for 3 <= c < tw + 3
  const bool ccol = c >= 3 && c < tw + 3;
  const int ig = i0 + cc;              // global compute column
  const bool store_col = ccol && ig < P.ni;
  const int flags = P.corner_flags ? __ldg(P.corner_flags + b) : 0;
  const int ia = i0 - 3 + c;
  const bool patch = flags != 0 && ycol && (ia < 0 || (ia >= P.ni && ia < P.ni + 3));

  // register state
  T qw0 = T(0), qw1 = T(0), qw2 = T(0), qw3 = T(0), qw4 = T(0), qw5 = T(0);  // q rows r-5 .. r (y-sweep view)
  T jw0 = T(0), jw1 = T(0), jw2 = T(0), jw3 = T(0), jw4 = T(0), jw5 = T(0);  // q_j rows r-5 .. r
  T qal_a = T(0), qal_b = T(0), jal_a = T(0), jal_b = T(0);                  // carried interface values of the two windows
  T q_m3 = T(0), q_m2 = T(0), q_m1 = T(0);                                   // q rows r-3 .. r-1 as stored (direction-1 corners)
  T fyy_prev = T(0), yf_prev = T(0), fya_prev = T(0);                        // interface r-3: yfx*fy2, yfx, averaged flux
  T cx_ring[RB], xf_ring[RB], fx2_ring[RB], ar_ring[RB];                     // rows r-3 .. r of this thread's interface / column
#pragma unroll
  for (int u = 0; u < RB; ++u) cx_ring[u] = xf_ring[u] = fx2_ring[u] = ar_ring[u] = T(0);
  T ra_next = T(0);  // rarea of the row that will be stored next iteration

  for (int m = 0; m < nchunk; ++m) {
    mbar_wait(&full[m & 1], (m >> 1) & 1);
    const unsigned char* st = smem + (m & 1) * G::STAGE_BYTES;
    const T* Qs = reinterpret_cast<const T*>(st + G::Q_OFF) + P.s_q;    // [RB][WQ] (row, column i0-3+c)
    const T* ARs = reinterpret_cast<const T*>(st + G::AR_OFF) + P.s_ar;
    const T* CXs = reinterpret_cast<const T*>(st + G::CX_OFF) + P.s_cx;  // [RB][WX] (row, interface i0+c)
    const T* XFs = reinterpret_cast<const T*>(st + G::XF_OFF) + P.s_xf;
    const T* CYs = reinterpret_cast<const T*>(st + G::CY_OFF) + P.s_cy;  // [RB][WQ] (interface r-2, column i0-3+c)
    const T* YFs = reinterpret_cast<const T*>(st + G::YF_OFF) + P.s_yf;
#pragma unroll
    for (int rr = 0; rr < RB; ++rr) {
      const int n = m * RB + rr;  // iteration; uniform over the CTA
      if (n < niter) {
        const int r = r0 + n;
        const int par = n & 1;
        constexpr int kNow = 0;  // ring slot helpers below use rr directly
        (void)kNow;
        // ---- 1. inner x-sweep of row r, thread = interface ----
        T fx2 = T(0), cx = T(0), xf = T(0);
        if (xint) {
          const T* row = Qs + rr * WQ + c;
          const T x0 = row[0], x1 = row[1], x2 = row[2], x3 = row[3], x4 = row[4], x5 = row[5];
          cx = CXs[rr * WX + c];
          xf = XFs[rr * WX + c];
          fx2 = ppm_flux_from_al(x2, x3, ppm_al(x0, x1, x2, x3), ppm_al(x1, x2, x3, x4), ppm_al(x2, x3, x4, x5), cx);
          fx2row[par * PP + c] = fx2;
        }
        // ---- 2. inner y-sweep: interface r-2, then q_i of row r-3; thread = tile column ----
        T fy2 = T(0), yf = T(0), cy = T(0), q_new = T(0), ar_new = T(0);
        if (ycol) {
          q_new = Qs[rr * WQ + c];
          ar_new = ARs[rr * WQ + c];
          T qy = q_new;
          if (patch && (r < 0 || r >= P.nj) && r < P.nj + 3) {  // copy_corners direction 2 (see k_fv_split.cu)
            const int bit = ia < 0 ? (r < 0 ? 1 : 4) : (r < 0 ? 2 : 8);
            if (flags & bit) {
              int si, sj;
              if (bit == 1) si = -r - 1, sj = ia;
              else if (bit == 2) si = P.ni + r, sj = P.ni - 1 - ia;
              else if (bit == 8) si = P.ni + P.nj - 1 - r, sj = ia - P.ni + P.nj;
              else si = r - P.nj, sj = P.nj - 1 - ia;
              qy = __ldg(P.q.at(si, sj, k, b));
            }
          }
          qw0 = qw1, qw1 = qw2, qw2 = qw3, qw3 = qw4, qw4 = qw5, qw5 = qy;
          const T al_c = ppm_al(qw2, qw3, qw4, qw5);
          cy = CYs[rr * WQ + c];
          yf = YFs[rr * WQ + c];
          fy2 = ppm_flux_from_al(qw2, qw3, qal_a, qal_b, al_c, cy);  // interface r-2: between rows r-3 (qw2) and r-2 (qw3)
          qal_a = qal_b, qal_b = al_c;
          const T fyy = mul_rn(yf, fy2);
          // q_i of row r-3: interfaces r-3 (previous iteration) and r-2; q of row r-3 as stored, its area from the ring
          const T arj = ar_ring[(rr + 1) % RB];
          const T ra = add_rn(arj, sub_rn(yf_prev, yf));
          qirow[par * PP + c] = mul_rn(fma_rn(q_m3, arj, sub_rn(fyy_prev, fyy)), rcp_fast_(ra));
          fyy_prev = fyy, yf_prev = yf;
        }
        __syncthreads();
        // ---- 3. q_j of row r, outer y-sweep at interface r-2; thread = compute column ----
        T fya = T(0);
        if (ccol) {
          const T xl = XFs[rr * WX + cc], xh = XFs[rr * WX + cc + 1];
          const T ra = add_rn(ar_new, sub_rn(xl, xh));
          const T num = fma_rn(q_new, ar_new, sub_rn(mul_rn(xl, fx2row[par * PP + cc]), mul_rn(xh, fx2row[par * PP + cc + 1])));
          const T qj = mul_rn(num, rcp_fast_(ra));
          jw0 = jw1, jw1 = jw2, jw2 = jw3, jw3 = jw4, jw4 = jw5, jw5 = qj;
          const T al_c = ppm_al(jw2, jw3, jw4, jw5);
          const T fo = ppm_flux_from_al(jw2, jw3, jal_a, jal_b, al_c, cy);
          jal_a = jal_b, jal_b = al_c;
          fya = mul_rn(mul_rn(T(0.5), add_rn(fo, fy2)), yf);  // averaged y-flux at interface r-2
          const int jf = r - 2;
          if (n >= 5 && jf >= jc0 && jf <= jc1 && store_col && P.fyo.p) __stcs(P.fyo.at(ig, jf, k, b), fya);
        }
        // ---- 4. outer x-sweep of row r-3 on q_i; thread = interface ----
        if (xint) {
          const T* row = qirow + par * PP + c;
          const T x0 = row[0], x1 = row[1], x2 = row[2], x3 = row[3], x4 = row[4], x5 = row[5];
          const T cxj = cx_ring[(rr + 1) % RB], xfj = xf_ring[(rr + 1) % RB], f2j = fx2_ring[(rr + 1) % RB];
          const T fo = ppm_flux_from_al(x2, x3, ppm_al(x0, x1, x2, x3), ppm_al(x1, x2, x3, x4), ppm_al(x2, x3, x4, x5), cxj);
          const T fxa = mul_rn(mul_rn(T(0.5), add_rn(fo, f2j)), xfj);
          fxarow[par * PP + c] = fxa;
          const int j = r - 3;
          if (n >= 6 && j < jc1 && P.fxo.p && i0 + c <= P.ni) __stcs(P.fxo.at(i0 + c, j, k, b), fxa);
        }
        // rings: this row's values replace row r-4's
        cx_ring[rr] = cx, xf_ring[rr] = xf, fx2_ring[rr] = fx2, ar_ring[rr] = ar_new;
        __syncthreads();
        // ---- 5. update of row r-3; thread = compute column ----
        {
          const int j = r - 3;
          if (ccol) {
            if (n >= 6 && j < jc1 && store_col) {
              const T fxl = fxarow[par * PP + cc], fxh = fxarow[par * PP + cc + 1];
              __stcs(P.qout.at(ig, j, k, b), fma_rn(ra_next, add_rn(sub_rn(fxl, fxh), sub_rn(fya_prev, fya)), q_m3));
            }
            fya_prev = fya;
            // rarea of the row stored next iteration (r - 2), in flight across one iteration
            const int jn = j + 1;
            if (store_col && jn >= jc0 && jn < jc1) ra_next = __ldg(P.rarea.at(ig, jn, b));
          }
          q_m3 = q_m2, q_m2 = q_m1, q_m1 = q_new;
        }
      }
    }
    // every read of this stage is behind the last barrier of its last row: refill it with chunk m + 2
    if (tid == 0 && m + 2 < nchunk) issue(m + 2);
  }
}

template <typename T, int TI>
int launch_stream(int ni, int nj, int nk, int nb, F3<const T> q, F3<const T> crx, F3<const T> xfx, F3<const T> cry,
                  F3<const T> yfx, F2<const T> area, F2<const T> rarea, const int* corner_flags, F3<T> q_out, F3<T> fxo,
                  F3<T> fyo, cudaStream_t s, bool* applicable) {
  using G = StreamTile<T, TI>;
  constexpr int V = G::V;
  *applicable = false;
  const TmaField<T> fq = tma_field<T>(q.p - 3 - 3 * q.sj, q.sj, q.sk, q.sb, nk, nb);
  const TmaField<T> fa = tma_field<T>(area.p - 3 - 3 * area.sj, area.sj, area.sj * (nj + 6), area.sb, 1, nb);
  const TmaField<T> fcx = tma_field<T>(crx.p - 3 * crx.sj, crx.sj, crx.sk, crx.sb, nk, nb);
  const TmaField<T> fxx = tma_field<T>(xfx.p - 3 * xfx.sj, xfx.sj, xfx.sk, xfx.sb, nk, nb);
  const TmaField<T> fcy = tma_field<T>(cry.p - 3, cry.sj, cry.sk, cry.sb, nk, nb);
  const TmaField<T> fyx = tma_field<T>(yfx.p - 3, yfx.sj, yfx.sk, yfx.sb, nk, nb);
  if (!(fq.ok && fa.ok && fcx.ok && fxx.ok && fcy.ok && fyx.ok)) return B2S_OK;
  CUtensorMap mq, ma, mcx, mxx, mcy, myx;
  const bool ok =
      make_map<T>(&mq, fq.base, q.sj, q.sk, q.sb, ni + 6 + fq.off, nj + 6, nk, nb, G::WQ, RB) &&
      make_map<T>(&ma, fa.base, area.sj, area.sj * (nj + 6), area.sb, ni + 6 + fa.off, nj + 6, 1, nb, G::WQ, RB) &&
      make_map<T>(&mcx, fcx.base, crx.sj, crx.sk, crx.sb, ni + 1 + fcx.off, nj + 6, nk, nb, G::WX, RB) &&
      make_map<T>(&mxx, fxx.base, xfx.sj, xfx.sk, xfx.sb, ni + 1 + fxx.off, nj + 6, nk, nb, G::WX, RB) &&
      make_map<T>(&mcy, fcy.base, cry.sj, cry.sk, cry.sb, ni + 6 + fcy.off, nj + 1, nk, nb, G::WQ, RB) &&
      make_map<T>(&myx, fyx.base, yfx.sj, yfx.sk, yfx.sb, ni + 6 + fyx.off, nj + 1, nk, nb, G::WQ, RB);
  if (!ok) return B2S_OK;
  auto kern = k_fv_split_stream<T, TI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return set_error((int)e, "fv_tp2d_split(stream): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  StreamParams<T> P;
  P.ni = ni, P.nj = nj, P.nk = nk;
  P.nstrips = (ni + TI - 1) / TI;
  P.tw = round_up_((ni + P.nstrips - 1) / P.nstrips, V);
  // rows per CTA: enough CTAs to fill the machine several times over, few enough rows of warm-up (6 per block);
  // measured on C384x72 fp64: 32 rows 1.24 ms, 64 rows 1.18 ms, 128 rows 1.14 ms
  int jb = option("fv_split_jb", 0);
  if (jb <= 0) jb = 128;
  const int nblk = (nj + jb - 1) / jb;
  P.jb = (nj + nblk - 1) / nblk;
  P.njblk = (nj + P.jb - 1) / P.jb;
  const int64_t nitems = (int64_t)P.nstrips * P.njblk * nk * nb;
  if (nitems > 0x7fffffffLL) return B2S_OK;
  static_assert(TI % V == 0, "tile width must keep the box start alignment from strip to strip");
  P.s_q = fq.off % V, P.c_q = fq.off - P.s_q;
  P.s_ar = fa.off % V, P.c_ar = fa.off - P.s_ar;
  P.s_cx = fcx.off % V, P.c_cx = fcx.off - P.s_cx;
  P.s_xf = fxx.off % V, P.c_xf = fxx.off - P.s_xf;
  P.s_cy = fcy.off % V, P.c_cy = fcy.off - P.s_cy;
  P.s_yf = fyx.off % V, P.c_yf = fyx.off - P.s_yf;
  P.rarea = rarea;
  P.q = q, P.corner_flags = corner_flags;
  P.qout = q_out, P.fxo = fxo, P.fyo = fyo;
  *applicable = true;
  kern<<<(unsigned)nitems, G::THREADS, G::SMEM_BYTES, s>>>(mq, ma, mcx, mxx, mcy, myx, P);
  return check_launch("fv_tp2d_split");
}

}  // namespace

template <typename T>
int fv_tp2d_split_stream(int ti, int ni, int nj, int nk, int nb, F3<const T> q, F3<const T> crx, F3<const T> xfx,
                         F3<const T> cry, F3<const T> yfx, F2<const T> area, F2<const T> rarea, const int* corner_flags,
                         F3<T> q_out, F3<T> fx_out, F3<T> fy_out, cudaStream_t s, bool* applicable) {
  if (ti == 120)
    return launch_stream<T, 120>(ni, nj, nk, nb, q, crx, xfx, cry, yfx, area, rarea, corner_flags, q_out, fx_out, fy_out, s, applicable);
  return launch_stream<T, 56>(ni, nj, nk, nb, q, crx, xfx, cry, yfx, area, rarea, corner_flags, q_out, fx_out, fy_out, s, applicable);
}

#define INSTANTIATE(T)                                                                                                 \
  template int fv_tp2d_split_stream<T>(int, int, int, int, int, F3<const T>, F3<const T>, F3<const T>, F3<const T>,    \
                                       F3<const T>, F2<const T>, F2<const T>, const int*, F3<T>, F3<T>, F3<T>,          \
                                       cudaStream_t, bool*);
INSTANTIATE(double)
INSTANTIATE(float)

}  // namespace impl
}  // namespace b2s
