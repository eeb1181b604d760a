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
__device__ __forceinline__ float rcp_fast_(float x) { return __frcp_rn(x); }  // kept IEEE: the tile kernel must match bit for bit

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
  static constexpr int P = TI + 16;                                     // pitch of the exchange rows (idle lanes read up to column THREADS + 5)
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

template <typename T, int TI, bool FLUX_OUT, int RP>
__global__ void __launch_bounds__(StreamTile<T, TI>::THREADS, (sizeof(T) == 8 ? 512 : 704) / StreamTile<T, TI>::THREADS) k_fv_split_stream(
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

  // Roles of this thread.  EVERY thread runs EVERY phase of every row, whatever its role: lanes without a role
  // in a phase compute on whatever sits in shared memory at their (in-bounds) index and their results are never
  // read or stored.  The first version guarded each phase with `if (role)`: the merges cost 47 register moves, 30
  // branch / reconvergence instructions and 9 zero-fills per warp-row, a quarter of all issued instructions
  // (profiles/r01_next_rows.md).  Only global stores and loads carry predicates.
  const int c = tid;                   // tile column (compute column i0 - 3 + c), x-interface i0 + c
  const int cc = c - 3;                // compute column of the tile, valid for 3 <= c < tw + 3
  const int ccr = max(cc, 0);          // shared-row index of the compute-column phases (idle lanes read column 0)
  const bool ccol = c >= 3 && c < tw + 3;
  const int ig = i0 + cc;              // global compute column
  const bool store_col = ccol && ig < P.ni;
  const bool fx_col = FLUX_OUT && c <= tw && i0 + c <= P.ni && P.fxo.p != nullptr;
  const bool fy_col = FLUX_OUT && store_col && P.fyo.p != nullptr;
  const int flags = P.corner_flags ? __ldg(P.corner_flags + b) : 0;
  const int ia = i0 - 3 + c;
  const bool patch = flags != 0 && c < tw + 6 && (ia < 0 || (ia >= P.ni && ia < P.ni + 3));
  const int nrows = jc1 - jc0;

  // running pointers of the rows touched at iteration n (advanced once per row, dereferenced under the row guards):
  // q_out and fx_out row r - 3 = jc0 - 6 + n, fy_out row r - 2 = jc0 - 5 + n
  T* qo_p = P.qout.at(ig, jc0 - 6, k, b);
  const T* ra_p = P.rarea.at(ig, jc0 - 6 + RP, b);  // row stored at iteration n + RP
  T* fx_p = FLUX_OUT ? P.fxo.at(i0 + c, jc0 - 6, k, b) : nullptr;
  T* fy_p = FLUX_OUT ? P.fyo.at(ig, jc0 - 5, k, b) : nullptr;
  const int64_t qo_sj = P.qout.sj, ra_sj = P.rarea.sj, fx_sj = P.fxo.sj, fy_sj = P.fyo.sj;

  // register state
  T qw2 = T(0), qw3 = T(0), qw4 = T(0), qw5 = T(0);  // q rows r-3 .. r (y-sweep view)
  T jw2 = T(0), jw3 = T(0), jw4 = T(0), jw5 = T(0);  // q_j rows r-3 .. r
  T qal_a = T(0), qal_b = T(0), jal_a = T(0), jal_b = T(0);                  // carried interface values of the two windows
  T q_m3 = T(0), q_m2 = T(0), q_m1 = T(0);                                   // q rows r-3 .. r-1 as stored (direction-1 corners)
  T fyy_prev = T(0), yf_prev = T(0), fya_prev = T(0);                        // interface r-3: yfx*fy2, yfx, averaged flux
  T cx_ring[RB], xf_ring[RB], fx2_ring[RB], ar_ring[RB];                     // rows r-3 .. r of this thread's interface / column
#pragma unroll
  for (int u = 0; u < RB; ++u) cx_ring[u] = xf_ring[u] = fx2_ring[u] = ar_ring[u] = T(0);
  T q_st[RP], cxj[RP], xfj[RP], f2j[RP];  // per row of a group: q of row r-3 as stored, interface values of row r-3

  static_assert(RB % 2 == 0 && RB % RP == 0 && (RP == 1 || RP == 2), "exchange rows are indexed by the parity of the chunk row");
  T ra_row[RP];  // rarea of the rows stored by the NEXT group (in flight across one group)
#pragma unroll
  for (int h = 0; h < RP; ++h) ra_row[h] = T(0);
  for (int m = 0; m < nchunk; ++m) {
    mbar_wait(&full[m & 1], (m >> 1) & 1);
    const unsigned char* st = smem + (m & 1) * G::STAGE_BYTES;
    const T* Qs = reinterpret_cast<const T*>(st + G::Q_OFF) + P.s_q + c;     // [RB][WQ] (row, column i0-3+c)
    const T* ARs = reinterpret_cast<const T*>(st + G::AR_OFF) + P.s_ar + c;
    const T* CXs = reinterpret_cast<const T*>(st + G::CX_OFF) + P.s_cx + c;  // [RB][WX] (row, interface i0+c)
    const T* XFs = reinterpret_cast<const T*>(st + G::XF_OFF) + P.s_xf;
    const T* CYs = reinterpret_cast<const T*>(st + G::CY_OFF) + P.s_cy + c;  // [RB][WQ] (interface r-2, column i0-3+c)
    const T* YFs = reinterpret_cast<const T*>(st + G::YF_OFF) + P.s_yf + c;
    // rows past the last one of the block (the chunk is always run in full) read TMA zero-fill or the next block's
    // rows; nothing of them is stored.
    // A group = RP consecutive rows taken through the five phases together (two barriers per group): with RP = 2
    // the two rows' flux chains are independent, which doubles the instruction-level parallelism between barriers.
#pragma unroll
    for (int g = 0; g < RB / RP; ++g) {
      T q_new[RP], ar_new[RP], cy[RP], yf[RP], fy2[RP], fya[RP];
      // ---- phases 1 + 2 ----
#pragma unroll
      for (int h = 0; h < RP; ++h) {
        const int rr = g * RP + h;
        const int r = r0 + m * RB + rr;
        const int par = rr & 1;  // == iteration parity
        // 1. inner x-sweep of row r, thread = interface
        const T* row = Qs + rr * WQ;
        const T x0 = row[0], x1 = row[1], x2 = row[2], x3 = row[3], x4 = row[4], x5 = row[5];
        const T cx = CXs[rr * WX];
        const T xf = XFs[rr * WX + c];
        const T fx2 = ppm_flux_from_al(x2, x3, ppm_al(x0, x1, x2, x3), ppm_al(x1, x2, x3, x4), ppm_al(x2, x3, x4, x5), cx);
        fx2row[par * PP + c] = fx2;
        // 2. inner y-sweep: interface r-2, then q_i of row r-3; thread = tile column
        q_new[h] = x0;
        ar_new[h] = ARs[rr * WQ];
        T qy = x0;
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
        qw2 = qw3, qw3 = qw4, qw4 = qw5, qw5 = qy;  // q rows r-3 .. r (y-sweep view)
        const T qal_c = ppm_al(qw2, qw3, qw4, qw5);
        cy[h] = CYs[rr * WQ];
        yf[h] = YFs[rr * WQ];
        fy2[h] = ppm_flux_from_al(qw2, qw3, qal_a, qal_b, qal_c, cy[h]);  // interface r-2: between rows r-3 (qw2) and r-2 (qw3)
        qal_a = qal_b, qal_b = qal_c;
        const T fyy = mul_rn(yf[h], fy2[h]);
        // q_i of row r-3: interfaces r-3 (previous row) and r-2; q of row r-3 as stored, its area from the ring
        const T arj = ar_ring[(rr + 1) % RB];
        const T ra = add_rn(arj, sub_rn(yf_prev, yf[h]));
        qirow[par * PP + c] = mul_rn(fma_rn(q_m3, arj, sub_rn(fyy_prev, fyy)), rcp_fast_(ra));
        fyy_prev = fyy, yf_prev = yf[h];
        q_st[h] = q_m3;                            // q of row r-3 as stored: the update of phase 5 starts from it
        q_m3 = q_m2, q_m2 = q_m1, q_m1 = x0;
        // rings: row r's values replace row r-4's (slot rr is not read again before that)
        cxj[h] = cx_ring[(rr + 1) % RB], xfj[h] = xf_ring[(rr + 1) % RB], f2j[h] = fx2_ring[(rr + 1) % RB];
        cx_ring[rr] = cx, xf_ring[rr] = xf, fx2_ring[rr] = fx2, ar_ring[rr] = ar_new[h];
      }
      __syncthreads();
      // ---- phases 3 + 4 ----
#pragma unroll
      for (int h = 0; h < RP; ++h) {
        const int rr = g * RP + h;
        const int n = m * RB + rr;
        const int par = rr & 1;
        // 3. q_j of row r, outer y-sweep at interface r-2; thread = compute column
        {
          const T xl = XFs[rr * WX + ccr], xh = XFs[rr * WX + ccr + 1];
          const T ra = add_rn(ar_new[h], sub_rn(xl, xh));
          const T num = fma_rn(q_new[h], ar_new[h], sub_rn(mul_rn(xl, fx2row[par * PP + ccr]), mul_rn(xh, fx2row[par * PP + ccr + 1])));
          const T qj = mul_rn(num, rcp_fast_(ra));
          jw2 = jw3, jw3 = jw4, jw4 = jw5, jw5 = qj;  // q_j rows r-3 .. r
          const T jal_c = ppm_al(jw2, jw3, jw4, jw5);
          const T fo = ppm_flux_from_al(jw2, jw3, jal_a, jal_b, jal_c, cy[h]);
          jal_a = jal_b, jal_b = jal_c;
          fya[h] = mul_rn(mul_rn(T(0.5), add_rn(fo, fy2[h])), yf[h]);  // averaged y-flux at interface r-2
          if (FLUX_OUT && fy_col && n >= 5 && n <= 5 + nrows) __stcs(fy_p, fya[h]);  // interfaces jc0 .. jc1
        }
        // 4. outer x-sweep of row r-3 on q_i; thread = interface
        {
          const T* qi = qirow + par * PP + c;
          const T y0 = qi[0], y1 = qi[1], y2 = qi[2], y3 = qi[3], y4 = qi[4], y5 = qi[5];
          const T fo = ppm_flux_from_al(y2, y3, ppm_al(y0, y1, y2, y3), ppm_al(y1, y2, y3, y4), ppm_al(y2, y3, y4, y5), cxj[h]);
          const T fxa = mul_rn(mul_rn(T(0.5), add_rn(fo, f2j[h])), xfj[h]);
          fxarow[par * PP + c] = fxa;
          if (FLUX_OUT && fx_col && (unsigned)(n - 6) < (unsigned)nrows) __stcs(fx_p, fxa);
        }
        if (FLUX_OUT) fx_p += fx_sj, fy_p += fy_sj;
      }
      __syncthreads();
      // ---- phase 5: update of rows r-3; thread = compute column ----
#pragma unroll
      for (int h = 0; h < RP; ++h) {
        const int rr = g * RP + h;
        const int n = m * RB + rr;
        const int par = rr & 1;
        const T fxl = fxarow[par * PP + ccr], fxh = fxarow[par * PP + ccr + 1];
        if (store_col && (unsigned)(n - 6) < (unsigned)nrows)
          __stcs(qo_p, fma_rn(ra_row[h], add_rn(sub_rn(fxl, fxh), sub_rn(fya_prev, fya[h])), q_st[h]));
        fya_prev = fya[h];
        qo_p += qo_sj;
      }
      // rarea of the rows the next group stores (rows r-3 of iterations n+RP .. n+2RP-1), in flight across one group
#pragma unroll
      for (int h = 0; h < RP; ++h) {
        const int n2 = m * RB + g * RP + h + RP;
        if (store_col && (unsigned)(n2 - 6) < (unsigned)nrows) ra_row[h] = __ldg(ra_p);
        ra_p += ra_sj;
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
  const bool flux_out = fxo.p != nullptr || fyo.p != nullptr;
  // rows per barrier pair: b2s_set_option("fv_split_rp", 1 | 2); the flux-output variant stays at 1 (registers)
  const int rp = flux_out ? 1 : (option("fv_split_rp", 0) == 1 ? 1 : 2);
  auto kern = flux_out ? k_fv_split_stream<T, TI, true, 1> : (rp == 2 ? k_fv_split_stream<T, TI, false, 2> : k_fv_split_stream<T, TI, false, 1>);
  static bool configured[3] = {false, false, false};
  const int slot = flux_out ? 2 : rp - 1;
  if (!configured[slot]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return set_error((int)e, "fv_tp2d_split(stream): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured[slot] = true;
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
