// K6d remap_ppm: FV3-style PPM vertical remap ("PPM-limited map_single remap (kord)").
// Spec: oracle/numpy_oracle.py ppm_profile + remap_ppm_column ([recalled] FV3 fv_mapz.F90 ppm_profile,
// ppm_limiters, map1_ppm; no source in /root/reference).
//
// Same slab design as k_remap_slab.cu (the dependent chain of the marching integration runs against
// shared memory, HBM sees each element once, in bulk):
//   * CTA = COLS (32, or 16 for tall fp64 columns) adjacent columns of one (j, b) row, 256 threads;
//   * two TMA tile loads (box COLS x 1 x levels; cp.async per element when TMA cannot address the
//     field) bring the block's pe1 and q1 columns into shared memory;
//   * the PPM profile is built IN shared memory by all 256 threads (thread = column x level range):
//       1. monotonised slopes dc[k]                      -> slab A
//       2. 4th-order interface values, interior         -> slab B (km+1 interfaces)
//       3. area-preserving cubics at the top and the surface (one thread per column)
//       4. ppm_limiters per layer: AL -> slab A (over dc), AR -> slab B (over the interface values;
//          every thread saves the one interface value it shares with the next level range first)
//     A6 is not stored: A6 = 3 (2 q - (AL + AR)) holds after every limiter branch;
//   * the nk2 target levels are split into chunks of CH levels (lane = column, chunk = warp, or
//     half-warp for 16 columns); target edges are loaded into registers before the wait on the slab;
//     the first source layer of a chunk comes from a binary search in the slab, then the integration
//     marches like map1_ppm.
// Divisions: fp64 uses MUFU.RCP64H + two Newton steps (<= 1 ulp) -- with IEEE divisions the profile
// alone costs ~9 slow-path-checked divisions per layer and the kernel becomes issue-bound; results
// agree with the oracle to a few ulp (tests: 1e-12 fp64 / 1e-5 fp32), not bit for bit.
// Algorithmic bytes/point: 24 R + 8 W.
#include "impl.cuh"
#include "tma.cuh"

namespace b2s {
namespace impl {

namespace {

constexpr int kThreads = 256;

template <typename T>
struct PpmParams {
  int ni, nj, nk1, nk2, ntile_i, kord, iv;
  int off_e, off_q;
  F3<const T> pe1, q1, pe2;
  F3<T> q2;
};

__device__ __forceinline__ double rcp_(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = __fma_rn(-x, r, 1.0);
  r = __fma_rn(r, e, r);
  e = __fma_rn(-x, r, 1.0);
  return __fma_rn(r, e, r);
}
__device__ __forceinline__ float rcp_(float x) {  // MUFU.RCP, <= 1 ulp: the IEEE-rounded __frcp_rn costs ~8 instructions
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// fp64 fmax / fmin compile to seven instructions each (DSETP.MAX + the NaN-propagation selects); a compare and a
// select is three.  Operands here are never NaN (finite inputs, reciprocals of positive thicknesses).
__device__ __forceinline__ double max_(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double min_(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }

template <typename T>
__device__ __forceinline__ T sgn_(T a, T b) {  // Fortran SIGN(a, b)
  return b >= T(0) ? fabs(a) : -fabs(a);
}

// ppm_limiters of fv_mapz.F90 on one layer
template <typename T>
__device__ __forceinline__ void ppm_limit(T dm, T q, T& al, T& ar, int lmt) {
  if (lmt == 3) return;
  T a6 = T(3) * (T(2) * q - (al + ar));
  if (lmt == 0) {
    if (dm == T(0)) {
      al = q, ar = q;
    } else {
      const T da1 = ar - al, da2 = da1 * da1, a6da = a6 * da1;
      if (a6da < -da2) {
        a6 = T(3) * (al - q);
        ar = al - a6;
      } else if (a6da > da2) {
        a6 = T(3) * (ar - q);
        al = ar - a6;
      }
    }
  } else if (lmt == 1) {
    const T qmp = T(2) * dm;
    al = q - sgn_(min_(fabs(qmp), fabs(al - q)), qmp);
    ar = q + sgn_(min_(fabs(qmp), fabs(ar - q)), qmp);
  } else {
    if (fabs(ar - al) < -a6) {
      const T d = ar - al;
      const T fm = q + T(0.25) * (d * d) * rcp_(a6) + a6 * T(1.0 / 12.0);
      if (fm < T(0)) {
        if (q < ar && q < al) {
          ar = q, al = q;
        } else if (ar > al) {
          a6 = T(3) * (al - q);
          ar = al - a6;
        } else {
          a6 = T(3) * (ar - q);
          al = ar - a6;
        }
      }
    }
  }
}

// 4th-order interface value between layers k-1 and k (2 <= k <= km-2):
//   d_m2 .. d_p1 = delp[k-2 .. k+1], qm = q[k-1], q0 = q[k], dcm = dc[k-1], dc0 = dc[k]
template <typename T>
__device__ __forceinline__ T ppm_iface(T d_m2, T dpm, T dp0, T d_p1, T qm, T q0, T dcm, T dc0) {
  const T d4m = d_m2 + dpm, d4k = dpm + dp0, d4p = dp0 + d_p1;  // d4[k-1], d4[k], d4[k+1]
  const T c1 = (q0 - qm) * dpm * rcp_(d4k);
  const T a1 = d4m * rcp_(d4k + dpm);
  const T a2 = d4p * rcp_(d4k + dp0);
  return qm + c1 + T(2) * rcp_(d4m + d4p) * (dp0 * (c1 * (a1 - a2) + a2 * dcm) - dpm * a1 * dc0);
}

struct PpmLayout {
  int e_off, q_off, a_off, b_off, bar_off, total;
  __host__ __device__ PpmLayout(int nk1, int rowb) {
    e_off = 0;
    q_off = ((nk1 + 1) * rowb + 127) / 128 * 128;
    a_off = (q_off + nk1 * rowb + 127) / 128 * 128;
    b_off = a_off + nk1 * rowb;
    bar_off = (b_off + (nk1 + 1) * rowb + 15) / 16 * 16;
    total = bar_off + 16;
  }
};

// LOADER: 0 cp.async, 1 TMA, 2 TMA with boxes starting at the aligned column before a misaligned field
template <typename T, int COLS, int CH, int LOADER, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_remap_ppm(const __grid_constant__ CUtensorMap tm_e,
                                                              const __grid_constant__ CUtensorMap tm_q,
                                                              const PpmParams<T> P) {
  constexpr int PITCH = COLS + (LOADER == 2 ? 16 / (int)sizeof(T) : 0);  // slab row pitch in elements
  constexpr int SUB = 32 / COLS;                                         // chunks per warp
  constexpr int NCHUNK = (kThreads / 32) * SUB;
  constexpr int NKL = kThreads / COLS;                                   // level ranges of the profile phases
  extern __shared__ __align__(1024) unsigned char smem[];
  const int km = P.nk1, nk2 = P.nk2;
  const PpmLayout L(km, PITCH * (int)sizeof(T));
  T* E = reinterpret_cast<T*>(smem + L.e_off);   // [km+1] source edges
  T* Q = reinterpret_cast<T*>(smem + L.q_off);   // [km]   source means
  T* A = reinterpret_cast<T*>(smem + L.a_off);   // [km]   dc, then AL
  T* B = reinterpret_cast<T*>(smem + L.b_off);   // [km+1] interface values, then AR
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);

  int t = blockIdx.x;
  const int ti = t % P.ntile_i;
  t /= P.ntile_i;
  const int j = t % P.nj;
  const int b = t / P.nj;

  // ---- 1. slab loads ----
  if (LOADER != 0) {
    if (threadIdx.x == 0) {
      mbar_init(&bar[0], 1);
      fence_barrier_init();
      mbar_arrive_expect_tx(&bar[0], (uint32_t)((2 * km + 1) * PITCH * sizeof(T)));
      tma_load_4d(E, &tm_e, &bar[0], ti * COLS, j, 0, b);
      tma_load_4d(Q, &tm_q, &bar[0], ti * COLS, j, 0, b);
    }
  } else {
    const int col = threadIdx.x % COLS, kl = threadIdx.x / COLS;
    const int i = ti * COLS + col;
    if (i < P.ni) {
      const T* ep = P.pe1.at(i, j, 0, b);
      const T* qp = P.q1.at(i, j, 0, b);
      for (int k = kl; k <= km; k += NKL) cp_async<sizeof(T)>(E + k * PITCH + col, ep + (int64_t)k * P.pe1.sk);
      for (int k = kl; k < km; k += NKL) cp_async<sizeof(T)>(Q + k * PITCH + col, qp + (int64_t)k * P.q1.sk);
    }
    cp_async_commit();
  }

  // ---- 2. target edges of this thread's first chunk (in flight while the slab arrives) ----
  const int lane = threadIdx.x & 31;
  const int mcol = lane % COLS;                                   // column during the marching phase
  const int chunk = (threadIdx.x >> 5) * SUB + lane / COLS;
  const int mi = ti * COLS + mcol;
  const bool mvalid = mi < P.ni;
  int k2b = chunk * CH;
  const int64_t sk2 = P.pe2.sk, sko = P.q2.sk;
  const T* e2 = P.pe2.at(mvalid ? mi : 0, j, k2b, b);
  T* o2 = P.q2.at(mvalid ? mi : 0, j, k2b, b);
  int nlev = mvalid ? min(CH, nk2 - k2b) : -1;  // < 0: no target edge is loaded either
  T tgt[CH + 1];
  {
    const T* p = e2;
#pragma unroll
    for (int u = 0; u <= CH; ++u) {
      tgt[u] = u <= nlev ? __ldg(p) : T(0);
      p += sk2;
    }
  }

  if (LOADER != 0) {
    __syncthreads();
    mbar_wait(&bar[0], 0);
  } else {
    cp_async_wait_all();
    __syncthreads();
  }

  // ---- 3. PPM profile in shared memory: thread = (column, contiguous level range [ka, kb)) ----
  // Two marching sweeps with the windows in registers (one new edge, one new mean per level; rcp(d4[k+1]) of a
  // slope is rcp(d4[k]) of the next one):  3.1 slopes dc -> slab A;  3.2 interface value k+1 by the 4th-order
  // formula, then ppm_limiters of layer k, AL -> slab A (over dc[k], no longer needed), AR -> slab B.  The two top
  // and two bottom layers (area-preserving cubics, standard constraint) are built by the level ranges 0 and 1
  // from the slabs.  The first version kept dc and the interface values in slabs between four phases with
  // strided level ownership: 7 reciprocals and ~280 instructions per layer (profiles/r01_next_rows.md).
  {
    const int col = threadIdx.x % COLS, kl = threadIdx.x / COLS;
    const T* Ec = E + col + (LOADER == 2 ? P.off_e : 0);
    const T* Qc = Q + col + (LOADER == 2 ? P.off_q : 0);
    T* Ac = A + col;
    T* Bc = B + col;
    const int per = (km + NKL - 1) / NKL;
    const int ka = min(km, kl * per), kb = min(km, ka + per);
    // 3.1 monotonised slopes dc[k], k in [ka, kb) and 1 .. km-2
    {
      const int ks = max(ka, 1), ke = min(kb, km - 1);
      if (ks < ke) {
        const T* ep = Ec + (ks - 1) * PITCH;
        const T* qp_ = Qc + (ks - 1) * PITCH;
        T* ap = Ac + ks * PITCH;
        T e2_ = ep[2 * PITCH];
        T dm = ep[PITCH] - ep[0], d0 = e2_ - ep[PITCH];  // delp[k-1], delp[k]
        T qm = qp_[0], q0 = qp_[PITCH];
        T r4k = rcp_(dm + d0);
        ep += 3 * PITCH, qp_ += 2 * PITCH;
#pragma unroll 3
        for (int k = ks; k < ke; ++k) {
          const T e3 = ep[0], qp = qp_[0];
          const T dp = e3 - e2_;  // delp[k+1]
          const T d4k = dm + d0, d4p = d0 + dp;
          const T r4p = rcp_(d4p);
          const T c1 = (dm + T(0.5) * d0) * r4p;
          const T c2 = (dp + T(0.5) * d0) * r4k;
          const T df2 = d0 * (c1 * (qp - q0) + c2 * (q0 - qm)) * rcp_(d4k + dp);
          const T qmax = max_(max_(qm, q0), qp) - q0;
          const T qmin = q0 - min_(min_(qm, q0), qp);
          ap[0] = sgn_(min_(min_(fabs(df2), qmax), qmin), df2);
          dm = d0, d0 = dp, e2_ = e3, qm = q0, q0 = qp, r4k = r4p;
          ep += PITCH, qp_ += PITCH, ap += PITCH;
        }
      }
    }
    __syncthreads();
    // interface value k (2 <= k <= km-2) from the slabs: the start of a march, and the cubics' anchor
    auto iface = [&](int k) -> T {
      const T e0 = Ec[(k - 2) * PITCH], e1 = Ec[(k - 1) * PITCH], e2_ = Ec[k * PITCH], e3 = Ec[(k + 1) * PITCH], e4 = Ec[(k + 2) * PITCH];
      return ppm_iface<T>(e1 - e0, e2_ - e1, e3 - e2_, e4 - e3, Qc[(k - 1) * PITCH], Qc[k * PITCH], Ac[(k - 1) * PITCH], Ac[k * PITCH]);
    };
    int lmt = max(0, P.kord - 3);
    if (P.iv == 0) lmt = min(2, lmt);
    // everything that reads a dc another thread will overwrite with AL comes first ...
    const int ga = max(ka, 2), gb = min(kb, km - 2);  // generic layers of this range: both interfaces interior
    T a_k = T(0), dc_k = T(0), dc_end = T(0);
    if (ga < gb) {
      a_k = iface(ga);             // reads dc[ga-1] (the previous range's when ga == ka)
      dc_k = Ac[ga * PITCH];
      dc_end = Ac[gb * PITCH];     // dc of the next range's first layer (gb <= km-2: a slope exists)
    }
    T bl0 = T(0), br0 = T(0), bl1 = T(0), br1 = T(0);  // (AL, AR) of the two boundary layers this thread builds
    if (kl == 0) {
      // top: area-preserving cubic with zero second derivative at the boundary
      const T d1 = Ec[PITCH] - Ec[0], d2 = Ec[2 * PITCH] - Ec[PITCH];
      const T q0 = Qc[0], q1 = Qc[PITCH];
      const T a2 = iface(2);
      const T r12 = rcp_(d1 + d2);
      const T qm = (d2 * q0 + d1 * q1) * r12;
      const T dq = T(2) * (q1 - q0) * r12;
      const T c1 = T(4) * (a2 - qm - d2 * dq) * rcp_(d2 * (T(2) * d2 * d2 + d1 * (d2 + T(3) * d1)));
      const T c3 = dq - T(0.5) * c1 * (d2 * (T(5) * d1 + d2) - T(3) * d1 * d1);
      T al1 = qm - T(0.25) * c1 * d1 * d2 * (d2 + T(3) * d1);
      T al0 = d1 * (T(2) * c1 * d1 * d1 - c3) + al1;
      al1 = max_(al1, min_(q0, q1));
      al1 = min_(al1, max_(q0, q1));
      const T dc0 = T(0.5) * (al1 - q0);
      if (P.iv == 0) al0 = max_(T(0), al0), al1 = max_(T(0), al1);
      bl0 = al0, br0 = al1;
      ppm_limit<T>(dc0, q0, bl0, br0, 0);
      bl1 = al1, br1 = a2;
      ppm_limit<T>(Ac[PITCH], q1, bl1, br1, 0);
    } else if (kl == 1) {
      // surface: the same cubic; bl0/br0 = layer km-2, bl1/br1 = layer km-1
      const T d1 = Ec[km * PITCH] - Ec[(km - 1) * PITCH], d2 = Ec[(km - 1) * PITCH] - Ec[(km - 2) * PITCH];
      const T q0 = Qc[(km - 1) * PITCH], q1 = Qc[(km - 2) * PITCH];
      const T a2 = iface(km - 2);
      const T r12 = rcp_(d1 + d2);
      const T qm = (d2 * q0 + d1 * q1) * r12;
      const T dq = T(2) * (q1 - q0) * r12;
      const T c1 = (a2 - qm - d2 * dq) * rcp_(d2 * (T(2) * d2 * d2 + d1 * (d2 + T(3) * d1)));
      const T c3 = dq - T(2) * c1 * (d2 * (T(5) * d1 + d2) - T(3) * d1 * d1);
      T alb = qm - c1 * d1 * d2 * (d2 + T(3) * d1);
      T arb = d1 * (T(8) * c1 * d1 * d1 - c3) + alb;
      alb = max_(alb, min_(q0, q1));
      alb = min_(alb, max_(q0, q1));
      const T dcb = T(0.5) * (q0 - alb);
      if (P.iv == 0) alb = max_(T(0), alb), arb = max_(T(0), arb);
      bl0 = a2, br0 = alb;
      ppm_limit<T>(Ac[(km - 2) * PITCH], q1, bl0, br0, 0);
      bl1 = alb, br1 = arb;
      ppm_limit<T>(dcb, q0, bl1, br1, 0);
    }
    __syncthreads();
    // ... then the stores
    if (kl == 0) {
      Ac[0] = bl0, Bc[0] = br0, Ac[PITCH] = bl1, Bc[PITCH] = br1;
    } else if (kl == 1) {
      Ac[(km - 2) * PITCH] = bl0, Bc[(km - 2) * PITCH] = br0, Ac[(km - 1) * PITCH] = bl1, Bc[(km - 1) * PITCH] = br1;
    }
    // 3.2 march over the generic layers: interface k+1, then the limiter of layer k
    if (ga < gb) {
      const T* ep = Ec + (ga - 1) * PITCH;
      const T* qp_ = Qc + ga * PITCH;
      T* ap = Ac + ga * PITCH;
      T* bp = Bc + ga * PITCH;
      T e_n = ep[3 * PITCH];                                          // edge k+2
      T dA = ep[PITCH] - ep[0], dB = ep[2 * PITCH] - ep[PITCH], dC = e_n - ep[2 * PITCH];  // delp[k-1], delp[k], delp[k+1]
      T q_k = qp_[0];
      ep += 4 * PITCH, qp_ += PITCH;
#pragma unroll 3
      for (int k = ga; k < gb; ++k) {
        const T e_nn = ep[0];             // edge k+3
        const T q_1 = qp_[0];             // q[k+1]
        const T dc_1 = k + 1 < gb ? ap[PITCH] : dc_end;  // dc[k+1]; slab A holds dc above this thread's stores
        const T dD = e_nn - e_n;          // delp[k+2]
        const T a_1 = ppm_iface<T>(dA, dB, dC, dD, q_k, q_1, dc_k, dc_1);
        T al = a_k, ar = a_1;
        ppm_limit<T>(dc_k, q_k, al, ar, lmt);
        ap[0] = al, bp[0] = ar;
        dA = dB, dB = dC, dC = dD, e_n = e_nn, q_k = q_1, dc_k = dc_1, a_k = a_1;
        ep += PITCH, qp_ += PITCH, ap += PITCH, bp += PITCH;
      }
    }
    __syncthreads();
  }

  // ---- 4. map1_ppm: integrate the parabolas over the target layers of this thread's chunks ----
  // The source layer under the top edge (edges, AL, AR, mean, A6, 1/dp) is carried in registers from one target
  // layer to the next: a target layer that ends in source layer m leaves exactly that state behind.
  const T* El = E + mcol + (LOADER == 2 ? P.off_e : 0);
  const T* Ql = Q + mcol + (LOADER == 2 ? P.off_q : 0);
  const T* Al = A + mcol;
  const T* Bl = B + mcol;
  const T r3 = T(1.0 / 3.0), r23 = T(2.0 / 3.0), half = T(0.5), one = T(1);
  while (nlev > 0) {
    // first source layer l with pe1[l+1] >= top edge (map1_ppm's `pe2 >= pe1(l) .and. pe2 <= pe1(l+1)`)
    int l;
    {
      const T x = tgt[0];
      int a = 0, z = km - 1;
      while (a < z) {
        const int m = (a + z) >> 1;
        if (El[(m + 1) * PITCH] >= x) z = m; else a = m + 1;
      }
      l = a;
    }
    T e0 = El[l * PITCH], e1 = El[(l + 1) * PITCH];
    T al = Al[l * PITCH], ar = Bl[l * PITCH];
    T a6, rdp;
    {
      const T qv = Ql[l * PITCH];
      a6 = T(3) * (T(2) * qv - (al + ar));
      rdp = rcp_(e1 - e0);
    }
    T* op = o2;
#pragma unroll
    for (int u = 0; u < CH; ++u) {
      if (u < nlev) {
        const T top = tgt[u], bot = tgt[u + 1];
        const T pl = (top - e0) * rdp;
        T res;
        if (bot <= e1) {  // the target layer lies inside source layer l
          const T pr = (bot - e0) * rdp;
          res = al + half * (a6 + ar - al) * (pr + pl) - a6 * r3 * (pr * (pr + pl) + pl * pl);
        } else {
          T qsum = (e1 - top) * (al + half * (a6 + ar - al) * (one + pl) - a6 * (r3 * (one + pl * (one + pl))));
          T em = e1;
          const T* ep = El + (l + 2) * PITCH;
          for (int m = l + 1; m < km; ++m) {
            const T em1 = ep[0];
            const T qm = Ql[m * PITCH];
            if (bot > em1) {  // whole layer
              qsum = qsum + (em1 - em) * qm;
              em = em1;
              ep += PITCH;
            } else {
              const T alm = Al[m * PITCH], arm = Bl[m * PITCH];
              const T a6m = T(3) * (T(2) * qm - (alm + arm));
              const T dp = bot - em;
              const T rdm = rcp_(em1 - em);
              const T esl = dp * rdm;
              qsum = qsum + dp * (alm + half * esl * (arm - alm + a6m * (one - r23 * esl)));
              l = m, e0 = em, e1 = em1, al = alm, ar = arm, a6 = a6m, rdp = rdm;
              break;
            }
          }
          res = qsum * rcp_(bot - top);
        }
        __stcs(op, res);
        op += sko;
      }
    }
    k2b += NCHUNK * CH;
    nlev = min(CH, nk2 - k2b);
    if (nlev <= 0) break;
    e2 += (int64_t)(NCHUNK * CH) * sk2;
    o2 += (int64_t)(NCHUNK * CH) * sko;
    {
      const T* p = e2;
#pragma unroll
      for (int u = 0; u <= CH; ++u) {
        tgt[u] = u <= nlev ? __ldg(p) : T(0);
        p += sk2;
      }
    }
  }
}

template <typename T, int COLS, int CH, int LOADER, int MINB>
int launch_ppm(const CUtensorMap& me, const CUtensorMap& mq, const PpmParams<T>& P, int nj, int nb, cudaStream_t s) {
  auto kern = k_remap_ppm<T, COLS, CH, LOADER, MINB>;
  constexpr int PITCH = COLS + (LOADER == 2 ? 16 / (int)sizeof(T) : 0);
  const size_t smem = (size_t)PpmLayout(P.nk1, PITCH * (int)sizeof(T)).total;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return set_error((int)e, "remap_ppm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  const int64_t grid = (int64_t)P.ntile_i * nj * nb;
  kern<<<(unsigned)grid, kThreads, smem, s>>>(me, mq, P);
  return check_launch("remap_ppm");
}

template <typename T, int COLS>
int dispatch_ppm(int loader, const CUtensorMap& me, const CUtensorMap& mq, const PpmParams<T>& P, int nj, int nb,
                 cudaStream_t s) {
  // chunk length CH: the smallest of 5 | 9 | 18 whose NCHUNK chunks cover the column in one round
  // (16-column CTAs have 16 chunks: CH = 5 for nk2 <= 80, 9 for nk2 <= 144; 32-column CTAs have 8)
  constexpr int NCHUNK = 8 * (32 / COLS);
  const int ch = P.nk2 <= NCHUNK * 5 ? 5 : (P.nk2 <= NCHUNK * 9 ? 9 : 18);
  // three CTAs per SM whenever their slabs fit (registers capped at 85 by the launch bounds), else two
  const bool three = 3 * ((size_t)PpmLayout(P.nk1, (COLS + (loader == 2 ? 16 / (int)sizeof(T) : 0)) * (int)sizeof(T)).total + 1024) <= (size_t)227 * 1024;
#define B2S_PPM(CH, LOADER) (three ? launch_ppm<T, COLS, CH, LOADER, 3>(me, mq, P, nj, nb, s) : launch_ppm<T, COLS, CH, LOADER, 2>(me, mq, P, nj, nb, s))
#define B2S_PPM_CH(LOADER) (ch == 5 ? B2S_PPM(5, LOADER) : (ch == 9 ? B2S_PPM(9, LOADER) : B2S_PPM(18, LOADER)))
  if (loader == 0) return B2S_PPM_CH(0);
  if (loader == 2) return B2S_PPM_CH(2);
  return B2S_PPM_CH(1);
#undef B2S_PPM_CH
#undef B2S_PPM
}

}  // namespace

// b2s_set_option("remap_ppm_cols", 0 | 16 | 32): columns per CTA (0: 32 unless three CTAs per SM only
// fit with 16); b2s_set_option("remap_ppm_loader", 0 auto | 1 cp.async | 2 TMA).
template <typename T>
int remap_ppm(int ni, int nj, int nk1, int nk2, int nb, int kord, int iv, F3<const T> pe1, F3<const T> q1,
              F3<const T> pe2, F3<T> q2, cudaStream_t s) {
  B2S_ARGCHECK(ni > 0 && nj > 0 && nk1 >= 4 && nk2 > 0 && nb > 0,
               "remap_ppm: empty domain or fewer than 4 source layers: %dx%dx(%d->%d)x%d", ni, nj, nk1, nk2, nb);
  B2S_ARGCHECK(kord >= 4 && kord <= 6, "remap_ppm: kord %d not in {4 monotone, 5 positive definite, 6 unlimited interior}", kord);
  B2S_ARGCHECK(iv == 0 || iv == 1, "remap_ppm: iv %d not in {0 positive definite scalar, 1 other scalars}", iv);
  B2S_ARGCHECK(pe1.p && q1.p && pe2.p && q2.p, "remap_ppm: null field");
  constexpr int V = 16 / (int)sizeof(T);
  const int want = option("remap_ppm_loader", 0);
  TmaField<T> fe{}, fq{};
  bool tma = want != 1 && nk1 + 1 <= 256;
  if (tma) {
    fe = tma_field<T>(pe1.p, pe1.sj, pe1.sk, pe1.sb, nk1 + 1, nb);
    fq = tma_field<T>(q1.p, q1.sj, q1.sk, q1.sb, nk1, nb);
    tma = fe.ok && fq.ok;
  }
  const bool shifted = tma && (fe.off != 0 || fq.off != 0);
  auto bytes = [&](int c) { return (size_t)PpmLayout(nk1, (c + (shifted ? V : 0)) * (int)sizeof(T)).total + 1024; };
  int cols = option("remap_ppm_cols", 0);
  // measured (profiles/r01_remap_ppm.md): what matters is the chunk each thread marches, not the CTA count --
  // 32 columns (8 chunks) when 9-level chunks cover the target column, else 16 columns (16 chunks)
  if (cols != 16 && cols != 32) cols = (nk2 <= 72 && 2 * bytes(32) <= (size_t)227 * 1024) ? 32 : 16;
  if (bytes(cols) > (size_t)227 * 1024 && cols == 32) cols = 16;
  if (bytes(cols) > (size_t)227 * 1024)
    return set_error(B2S_EUNSUPPORTED, "remap_ppm: %d source levels do not fit the shared-memory slabs (limit %d)", nk1,
                     (int)(((size_t)226 * 1024 / ((16 + V) * sizeof(T)) - 2) / 4));
  PpmParams<T> P;
  P.ni = ni, P.nj = nj, P.nk1 = nk1, P.nk2 = nk2, P.kord = kord, P.iv = iv;
  P.ntile_i = (ni + cols - 1) / cols;
  P.off_e = P.off_q = 0;
  P.pe1 = pe1, P.q1 = q1, P.pe2 = pe2, P.q2 = q2;
  if ((int64_t)P.ntile_i * nj * nb > 0x7fffffffLL) return set_error(B2S_EUNSUPPORTED, "remap_ppm: grid too large");
  CUtensorMap me, mq;
  int loader = 0;
  if (tma) {
    const int box = cols + (shifted ? V : 0);
    if (make_map<T>(&me, fe.base, pe1.sj, pe1.sk, pe1.sb, ni + fe.off, nj, nk1 + 1, nb, box, 1, nk1 + 1) &&
        make_map<T>(&mq, fq.base, q1.sj, q1.sk, q1.sb, ni + fq.off, nj, nk1, nb, box, 1, nk1)) {
      loader = shifted ? 2 : 1;
      P.off_e = fe.off, P.off_q = fq.off;
    }
  }
  if (loader == 0) {
    if (want == 2) return set_error(B2S_EUNSUPPORTED, "remap_ppm: remap_ppm_loader=2 forced but the fields do not meet the TMA rules");
    memset(&me, 0, sizeof(me));
    memset(&mq, 0, sizeof(mq));
  }
  return cols == 32 ? dispatch_ppm<T, 32>(loader, me, mq, P, nj, nb, s) : dispatch_ppm<T, 16>(loader, me, mq, P, nj, nb, s);
}

template int remap_ppm<double>(int, int, int, int, int, int, int, F3<const double>, F3<const double>, F3<const double>,
                               F3<double>, cudaStream_t);
template int remap_ppm<float>(int, int, int, int, int, int, int, F3<const float>, F3<const float>, F3<const float>,
                              F3<float>, cudaStream_t);

}  // namespace impl
}  // namespace b2s
