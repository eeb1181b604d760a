// K5b fv_tp2d_split: FV3's fv_tp_2d (inner/outer operator splitting) as one sm_100a kernel.
// Spec: oracle/numpy_oracle.py fv_tp2d_split ([recalled] FV3 tp_core.F90 fv_tp_2d with the unlimited PPM
// of S5 as xppm/yppm; no source in /root/reference).
//
//   fy2 = yppm(q)            q_i = (q area + d_y(yfx fy2)) / (area + d_y yfx)      fx = xppm(q_i)
//   fx2 = xppm(q)            q_j = (q area + d_x(xfx fx2)) / (area + d_x xfx)      fy = yppm(q_j)
//   fx <- 0.5 (fx + fx2) xfx ;  fy <- 0.5 (fy + fy2) yfx ;  q_out = q + rarea (d_x fx + d_y fy)
//
// Work item = one tw x R tile of one (k, b) level (tw <= TI columns: the domain width is cut into equal
// strips so that no strip is mostly empty), one CTA per item, 2-3 CTAs per SM so that the TMA loads of one
// tile overlap the arithmetic of another.  TI + 6 <= THREADS = a whole number of warps: with TI = 128 a
// fifth warp would run every FP64 instruction for 6 apron columns and double the load of one of the four
// schedulers (measured: 2.2 ms vs the 4-warp shape on C384x72).  One elected thread issues five TMA tile
// loads (q with the full 3-cell apron INCLUDING corners, crx/xfx with the j-apron, cry/yfx with the
// i-apron); cell areas (an IJ field, L2-resident across the nk levels) are loaded straight into registers
// before the wait.  The four sweeps then run out of shared memory:
//   A  thread = column (TI+6 of them): inner y-sweep down the column with the q window in registers
//      -> fy2 (kept for the average) and q_i, both into shared memory;
//   B  thread = x-interface (TI+1): inner x-sweep over the R+6 rows -> fx2, then (after a barrier)
//      thread = column: q_j for the R+6 rows;
//   C  thread = x-interface: outer x-sweep on q_i -> averaged flux fx; thread = column: outer y-sweep
//      on q_j -> averaged flux fy in registers;
//   D  thread = column: q_out, streamed to HBM (and fx / fy when the caller wants the fluxes).
// Cube corners (FV3 copy_corners): the halo-corner cells of q hold the values for x-sweeps (direction 1, filled
// by the halo update); where `corner_flags[b]` marks a halo corner as a cube corner, the inner y-sweep reads
// the direction-2 values instead -- a rotated copy of the sub-domain's own south / north halo, fetched
// straight from global memory by the three apron-column threads concerned.
// All arithmetic goes through the explicit-rounding helpers of fv_math.cuh; the two divisions per cell are
// MUFU.RCP64H + two Newton steps in fp64 (<= 1 ulp).  Parity with the oracle: 1e-12 (fp64) / 1e-5 (fp32).
// Algorithmic bytes/point: 40 R + 8 W (+ 16 W with the flux outputs) + 16/nk for area and rarea.
#include "fv_math.cuh"
#include "impl.cuh"
#include "tma.cuh"

namespace b2s {
namespace impl {

namespace {

constexpr int round_up(int x, int m) { return (x + m - 1) / m * m; }

__device__ __forceinline__ double rcp_fast(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = __fma_rn(-x, r, 1.0);
  r = __fma_rn(r, e, r);
  e = __fma_rn(-x, r, 1.0);
  return __fma_rn(r, e, r);
}
__device__ __forceinline__ float rcp_fast(float x) { return __frcp_rn(x); }

template <typename T, int TI, int R>
struct SplitTile {
  static constexpr int V = 16 / sizeof(T);
  static constexpr int WQ = round_up(TI + 6 + V - 1, V);   // q / area / cry / yfx box width (columns -3 .. TI+2, + shift)
  static constexpr int WX = round_up(TI + 1 + V - 1, V);   // crx / xfx box width (interfaces 0 .. TI, + shift)
  static constexpr int RQ = R + 6, RY = R + 1;
  static constexpr int Q_OFF = 0;
  static constexpr int CRX_OFF = Q_OFF + round_up(RQ * WQ * (int)sizeof(T), 128);
  static constexpr int XFX_OFF = CRX_OFF + round_up(RQ * WX * (int)sizeof(T), 128);
  static constexpr int CRY_OFF = XFX_OFF + round_up(RQ * WX * (int)sizeof(T), 128);
  static constexpr int YFX_OFF = CRY_OFF + round_up(RY * WQ * (int)sizeof(T), 128);
  static constexpr int TX_BYTES = (RQ * WQ + 2 * RQ * WX + 2 * RY * WQ) * (int)sizeof(T);
  // intermediates (plain shared arrays, pitch = TI + 8 elements)
  static constexpr int P = TI + 8;
  static constexpr int FY2_OFF = YFX_OFF + round_up(RY * WQ * (int)sizeof(T), 128);   // [R+1][P] column c+3
  static constexpr int QI_OFF = FY2_OFF + RY * P * (int)sizeof(T);                    // [R][P]   column c+3
  static constexpr int FX2_OFF = QI_OFF + R * P * (int)sizeof(T);                     // [R+6][P] interface i
  static constexpr int QJ_OFF = FX2_OFF + RQ * P * (int)sizeof(T);                    // [R+6][P] column c
  static constexpr int FXA_OFF = QJ_OFF + RQ * P * (int)sizeof(T);                    // [R][P]   interface i
  static constexpr int BAR_OFF = round_up(FXA_OFF + R * P * (int)sizeof(T), 16);
  static constexpr int SMEM_BYTES = BAR_OFF + 16;
  static constexpr int THREADS = round_up(TI + 6, 32);
  static_assert(WQ <= 256 && RQ <= 256, "TMA box dimensions are limited to 256 elements");
};

template <typename T>
struct SplitParams {
  int ni, nj, nk, nstrips, njblk, tw;    // tw = columns per strip (<= TI, multiple of 16 bytes)
  int c_q, c_crx, c_xfx, c_cry, c_yfx;  // TMA coordinate of the first box column of strip 0
  int s_q, s_crx, s_xfx, s_cry, s_yfx;  // elements to skip inside a box row (16-byte alignment shift)
  F2<const T> area, rarea;              // area addresses compute cell (0, 0); its halo sits at negative offsets
  F3<const T> q;                        // compute cell (0, 0, 0): only read for the direction-2 corner values
  const int* corner_flags;              // per sub-domain: 1 SW | 2 SE | 4 NW | 8 NE halo corner is a cube corner; may be NULL
  F3<T> qout, fxo, fyo;
};

template <typename T, int TI, int R>
__global__ void __launch_bounds__(SplitTile<T, TI, R>::THREADS) k_fv_split(const __grid_constant__ CUtensorMap tm_q,
                                                                           const __grid_constant__ CUtensorMap tm_crx,
                                                                           const __grid_constant__ CUtensorMap tm_xfx,
                                                                           const __grid_constant__ CUtensorMap tm_cry,
                                                                           const __grid_constant__ CUtensorMap tm_yfx,
                                                                           const SplitParams<T> P) {
  using G = SplitTile<T, TI, R>;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + G::BAR_OFF);
  const int tid = threadIdx.x;

  int t = blockIdx.x;
  const int jb = t % P.njblk;
  t /= P.njblk;
  const int strip = t % P.nstrips;
  t /= P.nstrips;
  const int k = t % P.nk;
  const int b = t / P.nk;
  const int tw = P.tw;
  const int i0 = strip * tw, j0 = jb * R;  // first compute cell of the tile

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(bar, G::TX_BYTES);
    // tensor maps are based at the halo origin of each field: row coordinate j0 = compute row j0 - 3 where a
    // field has a j-apron, column coordinate likewise
    tma_load_4d(smem + G::Q_OFF, &tm_q, bar, P.c_q + i0, j0, k, b);
    tma_load_4d(smem + G::CRX_OFF, &tm_crx, bar, P.c_crx + i0, j0, k, b);
    tma_load_4d(smem + G::XFX_OFF, &tm_xfx, bar, P.c_xfx + i0, j0, k, b);
    tma_load_4d(smem + G::CRY_OFF, &tm_cry, bar, P.c_cry + i0, j0, k, b);
    tma_load_4d(smem + G::YFX_OFF, &tm_yfx, bar, P.c_yfx + i0, j0, k, b);
  }
  // cell areas of this thread's columns, straight from global memory (L2) while the tiles arrive:
  // area_a[r] = column i0-3+tid, rows j0 .. j0+R-1 (phase A);  area_b[r] = column i0+tid, rows j0-3 .. j0+R+2 (phase B2).
  // Rows / columns beyond the halo-extended domain are clamped (their results are never stored).
  T area_a[R], area_b[R + 6];
  {
    const int ia = min(i0 - 3 + tid, P.ni + 2), ib = min(i0 + tid, P.ni + 2);
#pragma unroll
    for (int r = 0; r < R; ++r) area_a[r] = __ldg(P.area.at(ia, min(j0 + r, P.nj + 2), b));
#pragma unroll
    for (int r = 0; r < R + 6; ++r) area_b[r] = __ldg(P.area.at(ib, min(j0 - 3 + r, P.nj + 2), b));
  }
  __syncthreads();
  mbar_wait(bar, 0);

  // shared views; X(r, c): r = tile row index (0 = compute row j0-3 for fields with a j-apron, else j0),
  // c = tile column index (0 = compute column i0-3 for fields with an i-apron, else interface i0)
  const T* qs = reinterpret_cast<const T*>(smem + G::Q_OFF) + P.s_q;        // [R+6][WQ]  (j0-3+r, i0-3+c)
  const T* cxs = reinterpret_cast<const T*>(smem + G::CRX_OFF) + P.s_crx;   // [R+6][WX]  (j0-3+r, interface i0+c)
  const T* xfs = reinterpret_cast<const T*>(smem + G::XFX_OFF) + P.s_xfx;
  const T* cys = reinterpret_cast<const T*>(smem + G::CRY_OFF) + P.s_cry;   // [R+1][WQ]  (interface j0+r, i0-3+c)
  const T* yfs = reinterpret_cast<const T*>(smem + G::YFX_OFF) + P.s_yfx;
  T* fy2s = reinterpret_cast<T*>(smem + G::FY2_OFF);  // [R+1][P] (interface j0+r, column i0-3+c)
  T* qis = reinterpret_cast<T*>(smem + G::QI_OFF);    // [R][P]   (row j0+r, column i0-3+c)
  T* fx2s = reinterpret_cast<T*>(smem + G::FX2_OFF);  // [R+6][P] (row j0-3+r, interface i0+c)
  T* qjs = reinterpret_cast<T*>(smem + G::QJ_OFF);    // [R+6][P] (row j0-3+r, column i0+c)
  T* fxas = reinterpret_cast<T*>(smem + G::FXA_OFF);  // [R][P]   (row j0+r, interface i0+c)
  constexpr int WQ = G::WQ, WX = G::WX, PP = G::P;

  // ---- A: inner y-sweep, thread = column c of the apron-extended tile (compute column i0 - 3 + c) ----
  if (tid < tw + 6) {
    const int c = tid;
    // q of tile row r in this thread's column as the y-sweep must see it (oracle: copy_corners direction 2)
    const int flags = P.corner_flags ? __ldg(P.corner_flags + b) : 0;
    const int ia = i0 - 3 + c;
    const bool patch = flags != 0 && (ia < 0 || (ia >= P.ni && ia < P.ni + 3));
    auto qy = [&](int r) -> T {
      if (patch) {
        const int j = j0 - 3 + r;
        if ((j < 0 || j >= P.nj) && j < P.nj + 3) {
          const int bit = ia < 0 ? (j < 0 ? 1 : 4) : (j < 0 ? 2 : 8);
          if (flags & bit) {
            int si, sj;
            if (bit == 1) si = -j - 1, sj = ia;                                  // SW: q(i,j) = q(1-j, i)
            else if (bit == 2) si = P.ni + j, sj = P.ni - 1 - ia;                // SE: q(npy+j-1, npx-i)
            else if (bit == 8) si = P.ni + P.nj - 1 - j, sj = ia - P.ni + P.nj;  // NE: q(2*npy-1-j, i)
            else si = j - P.nj, sj = P.nj - 1 - ia;                              // NW: q(j+1-npx, npy-i)
            return __ldg(P.q.at(si, sj, k, b));
          }
        }
      }
      return qs[r * WQ + c];
    };
    // window of six q values around interface j0 + r: rows (r .. r+5) of the q tile
    T w0 = qy(0), w1 = qy(1), w2 = qy(2), w3 = qy(3), w4 = qy(4);
    T al_a = ppm_al(w0, w1, w2, w3);  // low-side interface value of the cell in row r+2 ... advanced below
    T al_b = ppm_al(w1, w2, w3, w4);
    T f_lo = T(0), y_lo = T(0);
#pragma unroll
    for (int r = 0; r <= R; ++r) {
      const T w5 = qy(r + 5);
      const T al_c = ppm_al(w2, w3, w4, w5);
      // interface j0+r lies between rows r+2 (low side) and r+3 (high side) of the tile
      const T f = ppm_flux_from_al(w2, w3, al_a, al_b, al_c, cys[r * WQ + c]);
      const T yf = yfs[r * WQ + c];
      fy2s[r * PP + c] = f;
      const T fyy = mul_rn(yf, f);
      if (r > 0) {
        // cell row j0 + r - 1 = tile row r + 2 - ... its value is w2 of the PREVIOUS step = current w1's successor:
        // at step r the low-side cell of interface r is tile row r+2; the cell just closed is tile row r+2 (between
        // interfaces r-1 and r), i.e. w2.
        const T ar = area_a[r - 1];
        const T ra = add_rn(ar, sub_rn(y_lo, yf));
        qis[(r - 1) * PP + c] = mul_rn(fma_rn(w2, ar, sub_rn(f_lo, fyy)), rcp_fast(ra));
      }
      f_lo = fyy;
      y_lo = yf;
      w0 = w1, w1 = w2, w2 = w3, w3 = w4, w4 = w5;
      al_a = al_b, al_b = al_c;
    }
  }
  // ---- B1: inner x-sweep, thread = x-interface c (interface i0 + c), all R+6 rows ----
  if (tid <= tw) {
    const int c = tid;
#pragma unroll
    for (int r = 0; r < R + 6; ++r) {
      const T* row = qs + r * WQ + c;  // row[m] = q at column i0 - 3 + c + m; interface c between row[2] and row[3]
      const T x0 = row[0], x1 = row[1], x2 = row[2], x3 = row[3], x4 = row[4], x5 = row[5];
      const T f = ppm_flux_from_al(x2, x3, ppm_al(x0, x1, x2, x3), ppm_al(x1, x2, x3, x4), ppm_al(x2, x3, x4, x5), cxs[r * WX + c]);
      fx2s[r * PP + c] = f;
    }
  }
  __syncthreads();
  // ---- B2: q_j, thread = compute column c (column i0 + c), all R+6 rows ----
  if (tid < tw) {
    const int c = tid;
#pragma unroll
    for (int r = 0; r < R + 6; ++r) {
      const T xl = xfs[r * WX + c], xh = xfs[r * WX + c + 1];
      const T ar = area_b[r];
      const T ra = add_rn(ar, sub_rn(xl, xh));
      const T num = fma_rn(qs[r * WQ + c + 3], ar, sub_rn(mul_rn(xl, fx2s[r * PP + c]), mul_rn(xh, fx2s[r * PP + c + 1])));
      qjs[r * PP + c] = mul_rn(num, rcp_fast(ra));
    }
  }
  // ---- C1: outer x-sweep on q_i, thread = x-interface c, R compute rows -> averaged flux ----
  if (tid <= tw) {
    const int c = tid;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const T* row = qis + r * PP + c;
      const T x0 = row[0], x1 = row[1], x2 = row[2], x3 = row[3], x4 = row[4], x5 = row[5];
      const T fo = ppm_flux_from_al(x2, x3, ppm_al(x0, x1, x2, x3), ppm_al(x1, x2, x3, x4), ppm_al(x2, x3, x4, x5), cxs[(r + 3) * WX + c]);
      fxas[r * PP + c] = mul_rn(mul_rn(T(0.5), add_rn(fo, fx2s[(r + 3) * PP + c])), xfs[(r + 3) * WX + c]);
    }
  }
  __syncthreads();
  // ---- C2 + D: outer y-sweep on q_j and the update, thread = compute column c ----
  if (tid < tw) {
    const int c = tid;
    const int i = i0 + c;
    const bool col_ok = i < P.ni;
    T w0 = qjs[0 * PP + c], w1 = qjs[1 * PP + c], w2 = qjs[2 * PP + c], w3 = qjs[3 * PP + c], w4 = qjs[4 * PP + c];
    T al_a = ppm_al(w0, w1, w2, w3);
    T al_b = ppm_al(w1, w2, w3, w4);
    T fy_lo = T(0);
#pragma unroll
    for (int r = 0; r <= R; ++r) {
      const T w5 = qjs[(r + 5) * PP + c];
      const T al_c = ppm_al(w2, w3, w4, w5);
      const T fo = ppm_flux_from_al(w2, w3, al_a, al_b, al_c, cys[r * WQ + c + 3]);
      const T fy = mul_rn(mul_rn(T(0.5), add_rn(fo, fy2s[r * PP + c + 3])), yfs[r * WQ + c + 3]);
      const int j = j0 + r;
      if (col_ok && j <= P.nj && P.fyo.p) __stcs(P.fyo.at(i, j, k, b), fy);
      if (r > 0) {
        const int jc = j - 1;  // the cell closed by interfaces r-1 and r: tile row r + 2
        if (col_ok && jc < P.nj) {
          const T fxl = fxas[(r - 1) * PP + c], fxh = fxas[(r - 1) * PP + c + 1];
          const T ra = __ldg(P.rarea.at(i, jc, b));
          const T q0 = qs[(r + 2) * WQ + c + 3];
          __stcs(P.qout.at(i, jc, k, b), fma_rn(ra, add_rn(sub_rn(fxl, fxh), sub_rn(fy_lo, fy)), q0));
        }
      }
      fy_lo = fy;
      w0 = w1, w1 = w2, w2 = w3, w3 = w4, w4 = w5;
      al_a = al_b, al_b = al_c;
    }
  }
  // fx output: thread = x-interface, R rows
  if (P.fxo.p && tid <= tw) {
    const int i = i0 + tid;
    if (i <= P.ni) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (j0 + r < P.nj) __stcs(P.fxo.at(i, j0 + r, k, b), fxas[r * PP + tid]);
    }
  }
}

template <typename T, int TI, int R>
int launch_split(int ni, int nj, int nk, int nb, F3<const T> q, F3<const T> crx, F3<const T> xfx, F3<const T> cry,
                 F3<const T> yfx, F2<const T> area, F2<const T> rarea, const int* corner_flags, F3<T> q_out, F3<T> fxo,
                 F3<T> fyo, cudaStream_t s, bool* applicable) {
  using G = SplitTile<T, TI, R>;
  constexpr int V = G::V;
  *applicable = false;
  // every tensor map starts at the halo origin of its field
  const TmaField<T> fq = tma_field<T>(q.p - 3 - 3 * q.sj, q.sj, q.sk, q.sb, nk, nb);
  const TmaField<T> fcx = tma_field<T>(crx.p - 3 * crx.sj, crx.sj, crx.sk, crx.sb, nk, nb);
  const TmaField<T> fxx = tma_field<T>(xfx.p - 3 * xfx.sj, xfx.sj, xfx.sk, xfx.sb, nk, nb);
  const TmaField<T> fcy = tma_field<T>(cry.p - 3, cry.sj, cry.sk, cry.sb, nk, nb);
  const TmaField<T> fyx = tma_field<T>(yfx.p - 3, yfx.sj, yfx.sk, yfx.sb, nk, nb);
  if (!(fq.ok && fcx.ok && fxx.ok && fcy.ok && fyx.ok)) return B2S_OK;
  CUtensorMap mq, mcx, mxx, mcy, myx;
  const bool ok =
      make_map<T>(&mq, fq.base, q.sj, q.sk, q.sb, ni + 6 + fq.off, nj + 6, nk, nb, G::WQ, G::RQ) &&
      make_map<T>(&mcx, fcx.base, crx.sj, crx.sk, crx.sb, ni + 1 + fcx.off, nj + 6, nk, nb, G::WX, G::RQ) &&
      make_map<T>(&mxx, fxx.base, xfx.sj, xfx.sk, xfx.sb, ni + 1 + fxx.off, nj + 6, nk, nb, G::WX, G::RQ) &&
      make_map<T>(&mcy, fcy.base, cry.sj, cry.sk, cry.sb, ni + 6 + fcy.off, nj + 1, nk, nb, G::WQ, G::RY) &&
      make_map<T>(&myx, fyx.base, yfx.sj, yfx.sk, yfx.sb, ni + 6 + fyx.off, nj + 1, nk, nb, G::WQ, G::RY);
  if (!ok) return B2S_OK;
  auto kern = k_fv_split<T, TI, R>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return set_error((int)e, "fv_tp2d_split: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  SplitParams<T> P;
  P.ni = ni, P.nj = nj, P.nk = nk;
  P.nstrips = (ni + TI - 1) / TI;
  P.tw = round_up((ni + P.nstrips - 1) / P.nstrips, V);  // equal strips, 16-byte multiples (<= TI since TI % V == 0)
  P.njblk = (nj + R - 1) / R;
  const int64_t nitems = (int64_t)P.nstrips * P.njblk * nk * nb;
  if (nitems > 0x7fffffffLL) return B2S_OK;
  static_assert(TI % V == 0, "tile width must keep the box start alignment from strip to strip");
  // a box must start on a 16-byte boundary: start at the aligned column at or before the needed one
  P.s_q = fq.off % V, P.c_q = fq.off - P.s_q;
  P.s_crx = fcx.off % V, P.c_crx = fcx.off - P.s_crx;
  P.s_xfx = fxx.off % V, P.c_xfx = fxx.off - P.s_xfx;
  P.s_cry = fcy.off % V, P.c_cry = fcy.off - P.s_cry;
  P.s_yfx = fyx.off % V, P.c_yfx = fyx.off - P.s_yfx;
  P.area = area, P.rarea = rarea;
  P.q = q, P.corner_flags = corner_flags;
  P.qout = q_out, P.fxo = fxo, P.fyo = fyo;
  *applicable = true;
  kern<<<(unsigned)nitems, G::THREADS, G::SMEM_BYTES, s>>>(mq, mcx, mxx, mcy, myx, P);
  return check_launch("fv_tp2d_split");
}

}  // namespace

// streaming variant (k_fv_split_stream.cu)
template <typename T>
int fv_tp2d_split_stream(int ti, int ni, int nj, int nk, int nb, F3<const T> q, F3<const T> crx, F3<const T> xfx,
                         F3<const T> cry, F3<const T> yfx, F2<const T> area, F2<const T> rarea, const int* corner_flags,
                         F3<T> q_out, F3<T> fx_out, F3<T> fy_out, cudaStream_t s, bool* applicable);

// b2s_set_option("fv_split_variant", 0 auto | 1 tile kernel | 2 streaming kernel);
// b2s_set_option("fv_split_ti", 0 | 56 | 120): maximum tile / strip width (0 = 56);
// TI + 6 and TI + 1 threads must fit a whole number of warps (64 and 128 threads)
template <typename T>
int fv_tp2d_split(int ni, int nj, int nk, int nb, F3<const T> q, F3<const T> crx, F3<const T> xfx, F3<const T> cry,
                  F3<const T> yfx, F2<const T> area, F2<const T> rarea, const int* corner_flags, F3<T> q_out,
                  F3<T> fx_out, F3<T> fy_out, cudaStream_t s) {
  B2S_ARGCHECK(ni > 0 && nj > 0 && nk > 0 && nb > 0, "fv_tp2d_split: empty domain %dx%dx%dx%d", ni, nj, nk, nb);
  B2S_ARGCHECK(q.p && crx.p && xfx.p && cry.p && yfx.p && area.p && rarea.p && q_out.p, "fv_tp2d_split: null field");
  int ti = option("fv_split_ti", 0);
  if (ti != 56 && ti != 120) ti = 56;  // measured on C384x72: 1.77 ms (56 x 8 tiles) vs 1.92 ms (120 x 4)
  bool applicable = false;
  int rc;
  const int variant = option("fv_split_variant", 0);
  if (variant != 1) {
    rc = fv_tp2d_split_stream<T>(ti, ni, nj, nk, nb, q, crx, xfx, cry, yfx, area, rarea, corner_flags, q_out, fx_out, fy_out, s,
                                 &applicable);
    if (applicable) return rc;
  }
  if (ti == 120)
    rc = launch_split<T, 120, 4>(ni, nj, nk, nb, q, crx, xfx, cry, yfx, area, rarea, corner_flags, q_out, fx_out, fy_out, s, &applicable);
  else
    rc = launch_split<T, 56, 8>(ni, nj, nk, nb, q, crx, xfx, cry, yfx, area, rarea, corner_flags, q_out, fx_out, fy_out, s, &applicable);
  if (applicable) return rc;
  return set_error(B2S_EUNSUPPORTED,
                   "fv_tp2d_split: the fields do not meet the TMA rules (element-aligned pointers, row / level / batch strides "
                   "that are multiples of 16 bytes); allocate them with b200stencil.fields");
}

#define INSTANTIATE(T)                                                                                                   \
  template int fv_tp2d_split<T>(int, int, int, int, F3<const T>, F3<const T>, F3<const T>, F3<const T>, F3<const T>,    \
                                F2<const T>, F2<const T>, const int*, F3<T>, F3<T>, F3<T>, cudaStream_t);
INSTANTIATE(double)
INSTANTIATE(float)

}  // namespace impl
}  // namespace b2s
