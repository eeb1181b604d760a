// Library-owned multi-GPU halo exchange: b2s_halo_init / alloc / plan / exchange_start / exchange_wait / finalize.
//
// What it replaces.  The reference's bridge passes the communicator THROUGH the C boundary (type `MPI`,
// /root/reference/src/tcn/py_ftn_interface/argument.py:54-86; MPI_Comm_f2c in templates/interface.c.jinja2 via
// base.py:78-96) and has an init / run / finalize triple (example_def_dycore.yaml:4,21,71); the halo update itself is
// NDSL's HaloUpdater over mpi4py (not vendored).  Here the whole exchange lives behind the C-ABI so that a C or Fortran
// caller gets the multi-GPU step without Python:
//   * rendezvous: ranks of ONE node (one process -- or one thread -- per GPU) meet in a POSIX shared-memory segment
//     named by a caller-supplied session string (the analogue of an ncclUniqueId / MPI communicator);
//   * b2s_halo_alloc: symmetric allocation -- every rank cudaMalloc's the same size and maps every peer's buffer
//     (cudaIpc handles across processes, the raw pointer between threads of one process), so loads and stores on a
//     peer address travel over NVLink / NVSwitch;
//   * b2s_halo_plan: an affine link table (built on the host from the cubed-sphere connectivity) bound to a field;
//   * b2s_halo_exchange: a one-block handshake kernel (announce, await the neighbours) + the flat-grid strip copies;
//   * b2s_halo_exchange_start / _wait: ONE kernel per halo update (k_halo_exchange, csrc/k_halo.cu) that carries the
//     neighbour handshake inside (release/acquire flags in peer memory, device-resident epoch, bounded waits), forked
//     onto the context's high-priority stream so it overlaps the caller's interior compute; both calls can be captured
//     into a CUDA graph;
//   * the gate: with gated != 0 the exchange kernel raises a device flag when the last halo cell has landed; gated
//     stencil launches (b2s_fv_tp2d_gated) compute the cells that read no halo first and acquire the flag before the
//     first TMA load that touches a halo cell.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/b200stencil.h"
#include "impl.cuh"

#include "halo_device.cuh"

namespace b2s {
namespace impl {
int halo_exchange_launch(int elem_size, int nb, int max_strip, const HaloXchg& X, bool narrow, cudaStream_t s);
int halo_exchange3_launch(int elem_size, int max_strip, const HaloXchg3& X, cudaStream_t s);
template <typename T>
int fv_tp2d_fused(int ni, int nj, int nk, int nb, F3<const T> q, F3<const T> crx, F3<const T> xfx, F3<const T> cry,
                  F3<const T> yfx, F2<const T> rarea, F3<T> q_out, const HaloXchg& xchg, cudaStream_t s);
}
}  // namespace b2s

using namespace b2s;

namespace {

constexpr int kMaxRanks = 64;
constexpr uint32_t kMagic = 0xB2005A10u;
constexpr int kStateWords = 512;  // device state: [0] epoch [1] blocks done [2] status; [32 + b] strips done of sub-domain b; [320 + r] strips pushed to rank r
constexpr int kGateOffset = 128;  // gate words: [b] halos of sub-domain b ready, [64] consumer CTAs done, [65] consumer status
constexpr int kGateSlots = 64;

struct Slot {
  cudaIpcMemHandle_t handle;
  int64_t pid;
  uint64_t ptr;
  int64_t nbytes;
  int device;
  int pad;
};

struct Segment {
  std::atomic<uint32_t> magic;
  std::atomic<uint32_t> attached;
  std::atomic<uint32_t> bar_count;
  std::atomic<uint32_t> bar_gen;
  std::atomic<uint32_t> failed;
  uint32_t world;
  Slot slots[kMaxRanks];
};

struct Allocation {
  void* local = nullptr;
  int64_t nbytes = 0;
  std::vector<void*> peers;        // address of every rank's buffer in this process (peers[rank] == local)
  std::vector<bool> opened;        // true: mapped with cudaIpcOpenMemHandle (close at free)
};

struct Plan {
  int64_t* links_dev = nullptr;  // [nlinks, 12], sorted by destination sub-domain
  int* b_total_dev = nullptr;    // [kGateSlots] exchange work units (link, level) per destination sub-domain
  int nlinks = 0, nk = 0, max_strip = 0, elem_size = 0, nb = 0;
  void* field = nullptr;
  int64_t remote_bytes = 0;
  unsigned long long peers = 0;  // ranks whose field some link reads
  bool narrow = true;            // every strip offset relative to its (link, level chunk) origin fits 32 bits (k_halo_exchange2 / 3)
  // the mixed table (k_halo_exchange3), present when the caller's table carries outgoing links: same-GPU strips pulled,
  // strips that cross NVLink pushed by the rank that owns the source
  int64_t* rows_dev = nullptr;   // [nrows, 14]
  int* push_total_dev = nullptr; // [world]
  int nrows = 0, nrows1 = 0;     // rows [0, nrows1) before the deliveries are awaited, the rest after
  unsigned long long wait_a = 0, wait_d = 0;
  bool mixed = false;
};

struct HaloCtx {
  std::string session, shm_name;
  int rank = 0, world = 1, device = 0;
  Segment* seg = nullptr;
  double timeout_s = 60.0;
  std::vector<Allocation> allocs;
  std::vector<Plan> plans;
  int* flags = nullptr;          // symmetric: int32[world] announcements written by the peers
  int64_t* peer_flags_dev = nullptr;
  int* state = nullptr;          // device int32[kStateWords]
  cudaStream_t comm = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool pending = false;
};

struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

void nap() {
  timespec ts{0, 20000};
  nanosleep(&ts, nullptr);
}

// live contexts: a handle is only dereferenced if it is in here (a finalized or made-up handle is refused)
std::mutex g_live_mu;
std::set<HaloCtx*> g_live;

HaloCtx* as_ctx(int64_t h) {
  HaloCtx* c = reinterpret_cast<HaloCtx*>(static_cast<intptr_t>(h));
  std::lock_guard<std::mutex> lk(g_live_mu);
  return g_live.count(c) ? c : nullptr;
}

#define B2S_CTX(c, h, what)                                                                  \
  HaloCtx* c = as_ctx(h);                                                                    \
  if (!c) return set_error(B2S_EINVAL, "%s: not a live halo context (b2s_halo_init first)", what)

#define B2S_CUDA(expr, what)                                                                  \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) return set_error((int)e_, "%s: %s", what, cudaGetErrorString(e_)); \
  } while (0)

// host barrier of the session (sense-reversing counter in the shared segment); bounded
int rdv_barrier(HaloCtx* c, const char* what) {
  if (c->world == 1) return B2S_OK;
  Segment* s = c->seg;
  const uint32_t gen = s->bar_gen.load(std::memory_order_acquire);
  if (s->bar_count.fetch_add(1, std::memory_order_acq_rel) + 1 == (uint32_t)c->world) {
    s->bar_count.store(0, std::memory_order_relaxed);
    s->bar_gen.store(gen + 1, std::memory_order_release);
    return B2S_OK;
  }
  const double t0 = now_s();
  while (s->bar_gen.load(std::memory_order_acquire) == gen) {
    if (s->failed.load(std::memory_order_relaxed))
      return set_error(B2S_ENOTINIT, "%s: another rank of session '%s' reported a failure", what, c->session.c_str());
    if (now_s() - t0 > c->timeout_s) {
      s->failed.store(1, std::memory_order_relaxed);
      return set_error(B2S_ENOTINIT, "%s: rank %d waited %.0f s for the other ranks of session '%s' (B2S_RDV_TIMEOUT)", what,
                       c->rank, c->timeout_s, c->session.c_str());
    }
    nap();
  }
  return B2S_OK;
}

int fail_all(HaloCtx* c, int rc) {
  if (c->seg) c->seg->failed.store(1, std::memory_order_relaxed);
  return rc;
}

// collective symmetric allocation; *out = local buffer
int sym_alloc(HaloCtx* c, int64_t nbytes, Allocation* out) {
  Allocation a;
  a.nbytes = nbytes;
  a.peers.assign(c->world, nullptr);
  a.opened.assign(c->world, false);
  cudaError_t e = cudaMalloc(&a.local, (size_t)nbytes);
  if (e != cudaSuccess) return fail_all(c, set_error((int)e, "b2s_halo_alloc: cudaMalloc(%lld): %s", (long long)nbytes, cudaGetErrorString(e)));
  a.peers[c->rank] = a.local;
  if (c->world > 1) {
    Slot& mine = c->seg->slots[c->rank];
    memset(&mine, 0, sizeof(Slot));
    e = cudaIpcGetMemHandle(&mine.handle, a.local);
    if (e != cudaSuccess) return fail_all(c, set_error((int)e, "b2s_halo_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e)));
    mine.pid = (int64_t)getpid();
    mine.ptr = reinterpret_cast<uint64_t>(a.local);
    mine.nbytes = nbytes;
    mine.device = c->device;
    int rc = rdv_barrier(c, "b2s_halo_alloc");
    if (rc) return rc;
    for (int r = 0; r < c->world; ++r) {
      if (r == c->rank) continue;
      const Slot& s = c->seg->slots[r];
      if (s.nbytes != nbytes)
        return fail_all(c, set_error(B2S_EINVAL, "b2s_halo_alloc: rank %d asked for %lld bytes, rank %d for %lld (the allocation is symmetric)",
                                     r, (long long)s.nbytes, c->rank, (long long)nbytes));
      if (s.pid == (int64_t)getpid()) {  // a thread of this process (virtual ranks, one-process multi-GPU)
        if (s.device != c->device) {
          int can = 0;
          cudaDeviceCanAccessPeer(&can, c->device, s.device);
          if (!can) return fail_all(c, set_error(B2S_EUNSUPPORTED, "b2s_halo_alloc: device %d cannot access device %d", c->device, s.device));
          e = cudaDeviceEnablePeerAccess(s.device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return fail_all(c, set_error((int)e, "cudaDeviceEnablePeerAccess(%d): %s", s.device, cudaGetErrorString(e)));
          cudaGetLastError();
        }
        a.peers[r] = reinterpret_cast<void*>(s.ptr);
      } else {
        void* p = nullptr;
        e = cudaIpcOpenMemHandle(&p, s.handle, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
          return fail_all(c, set_error((int)e, "b2s_halo_alloc: cudaIpcOpenMemHandle(rank %d, device %d): %s", r, s.device, cudaGetErrorString(e)));
        a.peers[r] = p;
        a.opened[r] = true;
      }
    }
    rc = rdv_barrier(c, "b2s_halo_alloc");  // every rank has read the slots: they may be reused
    if (rc) return rc;
  }
  *out = a;
  return B2S_OK;
}

void sym_free(HaloCtx* c, Allocation& a) {
  for (int r = 0; r < (int)a.peers.size(); ++r)
    if (a.opened[r] && a.peers[r]) cudaIpcCloseMemHandle(a.peers[r]);
  if (a.local) cudaFree(a.local);
  a.local = nullptr;
  (void)c;
}

Allocation* find_alloc(HaloCtx* c, const void* p) {
  const char* q = static_cast<const char*>(p);
  for (auto& a : c->allocs) {
    const char* b = static_cast<const char*>(a.local);
    if (b && q >= b && q < b + a.nbytes) return &a;
  }
  return nullptr;
}

}  // namespace

extern "C" int b2s_halo_init(const char* session, int rank, int world, int device, int64_t* ctx_out) {
  if (!ctx_out) return set_error(B2S_EINVAL, "b2s_halo_init: ctx_out is NULL");
  *ctx_out = 0;
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world)
    return set_error(B2S_EINVAL, "b2s_halo_init: rank %d of %d (1 <= world <= %d)", rank, world, kMaxRanks);
  if (world > 1 && (!session || !*session || strlen(session) > 200 || strchr(session, '/')))
    return set_error(B2S_EINVAL, "b2s_halo_init: world > 1 needs a session name (no '/', <= 200 chars) shared by all ranks");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return set_error(e == cudaSuccess ? B2S_ENOTINIT : (int)e, "b2s_halo_init: no CUDA device; libb200stencil has no CPU fallback");
  if (device < 0 || device >= ndev) return set_error(B2S_EINVAL, "b2s_halo_init: device %d out of range [0,%d)", device, ndev);
  DeviceGuard guard(device);
  HaloCtx* c = new HaloCtx;
  c->session = session ? session : "";
  c->rank = rank, c->world = world, c->device = device;
  if (const char* t = getenv("B2S_RDV_TIMEOUT")) c->timeout_s = atof(t) > 0 ? atof(t) : c->timeout_s;
  auto bail = [&](int rc) {
    if (c->seg) {
      c->seg->failed.store(1, std::memory_order_relaxed);
      munmap(c->seg, sizeof(Segment));
    }
    delete c;
    return rc;
  };
  if (world > 1) {
    c->shm_name = "/b2s_" + c->session;
    int fd = shm_open(c->shm_name.c_str(), O_CREAT | O_RDWR, 0600);
    if (fd < 0) return bail(set_error(B2S_ENOTINIT, "b2s_halo_init: shm_open(%s): %s", c->shm_name.c_str(), strerror(errno)));
    if (ftruncate(fd, sizeof(Segment)) != 0) {
      close(fd);
      return bail(set_error(B2S_ENOTINIT, "b2s_halo_init: ftruncate(%s): %s", c->shm_name.c_str(), strerror(errno)));
    }
    void* m = mmap(nullptr, sizeof(Segment), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return bail(set_error(B2S_ENOTINIT, "b2s_halo_init: mmap(%s): %s", c->shm_name.c_str(), strerror(errno)));
    c->seg = static_cast<Segment*>(m);  // a fresh segment is zero-filled, which is the initial state of every field
    uint32_t seen = 0;
    if (c->seg->magic.compare_exchange_strong(seen, kMagic, std::memory_order_acq_rel))
      c->seg->world = (uint32_t)world;  // the first rank to arrive stamps the segment
    else if (seen != kMagic)
      return bail(set_error(B2S_ENOTINIT, "b2s_halo_init: shared-memory object %s exists and is not a b2s rendezvous segment", c->shm_name.c_str()));
    c->seg->attached.fetch_add(1, std::memory_order_acq_rel);
    int rc = rdv_barrier(c, "b2s_halo_init");
    if (rc) return bail(rc);
    if (c->seg->world != (uint32_t)world)
      return bail(set_error(B2S_EINVAL, "b2s_halo_init: session '%s' was opened for %u ranks, rank %d says %d", c->session.c_str(), c->seg->world, rank, world));
    if (rank == 0) shm_unlink(c->shm_name.c_str());  // every rank holds a mapping now; the name can go
  }
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  e = cudaStreamCreateWithPriority(&c->comm, cudaStreamNonBlocking, hi);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&c->state, kStateWords * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(c->state, 0, kStateWords * sizeof(int));
  if (e != cudaSuccess) return bail(fail_all(c, set_error((int)e, "b2s_halo_init: %s", cudaGetErrorString(e))));
  // Load every kernel that can take part in a device-side wait NOW.  CUDA loads kernels lazily, at their first launch,
  // and a load may have to wait for the context to go idle (and holds a driver lock meanwhile): a first launch issued
  // while an exchange kernel spins on a neighbour -- whose own launch, in another thread of this process, needs that
  // lock -- would stall until the spin times out.  After the final barrier below nothing of the path is left to load.
  int rc = impl::halo_kernels_preload();
  if (!rc) rc = impl::fv_tma_preload();
  if (!rc) rc = impl::fv_stream_preload();
  if (rc) return bail(fail_all(c, rc));
  // announcement flags: one int32 per rank, in peer-mapped memory, zero before anyone announces
  Allocation fl;
  rc = sym_alloc(c, (int64_t)sizeof(int) * 2 * kMaxRanks, &fl);  // A (announced) flags, then D (delivered) flags
  if (rc) return bail(rc);
  c->allocs.push_back(fl);
  c->flags = static_cast<int*>(fl.local);
  e = cudaMemset(c->flags, 0, sizeof(int) * 2 * kMaxRanks);
  std::vector<int64_t> pf(world);
  for (int r = 0; r < world; ++r) pf[r] = (int64_t) reinterpret_cast<intptr_t>(fl.peers[r]);
  if (e == cudaSuccess) e = cudaMalloc(&c->peer_flags_dev, sizeof(int64_t) * world);
  if (e == cudaSuccess) e = cudaMemcpy(c->peer_flags_dev, pf.data(), sizeof(int64_t) * world, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return bail(fail_all(c, set_error((int)e, "b2s_halo_init: %s", cudaGetErrorString(e))));
  rc = rdv_barrier(c, "b2s_halo_init");  // nobody announces into flags that are not zeroed yet
  if (rc) return bail(rc);
  {
    std::lock_guard<std::mutex> lk(g_live_mu);
    g_live.insert(c);
  }
  *ctx_out = (int64_t) reinterpret_cast<intptr_t>(c);
  return B2S_OK;
}

extern "C" int b2s_halo_finalize(int64_t ctx) {
  B2S_CTX(c, ctx, "b2s_halo_finalize");
  DeviceGuard guard(c->device);
  cudaDeviceSynchronize();
  int rc = B2S_OK;
  if (c->seg && !c->seg->failed.load()) rc = rdv_barrier(c, "b2s_halo_finalize");  // no peer still pulls from my buffers
  for (auto& p : c->plans) {
    if (p.links_dev) cudaFree(p.links_dev);
    if (p.b_total_dev) cudaFree(p.b_total_dev);
    if (p.rows_dev) cudaFree(p.rows_dev);
    if (p.push_total_dev) cudaFree(p.push_total_dev);
  }
  for (auto& a : c->allocs) sym_free(c, a);
  if (c->peer_flags_dev) cudaFree(c->peer_flags_dev);
  if (c->state) cudaFree(c->state);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->comm) cudaStreamDestroy(c->comm);
  if (c->seg) munmap(c->seg, sizeof(Segment));
  {
    std::lock_guard<std::mutex> lk(g_live_mu);
    g_live.erase(c);
  }
  delete c;
  return rc;
}

extern "C" int b2s_halo_rank(int64_t ctx) {
  HaloCtx* c = as_ctx(ctx);
  return c ? c->rank : -1;
}

extern "C" int b2s_halo_world(int64_t ctx) {
  HaloCtx* c = as_ctx(ctx);
  return c ? c->world : -1;
}

extern "C" int b2s_halo_barrier(int64_t ctx) {
  B2S_CTX(c, ctx, "b2s_halo_barrier");
  return rdv_barrier(c, "b2s_halo_barrier");
}

extern "C" int b2s_halo_alloc(int64_t ctx, int64_t nbytes, void** ptr) {
  B2S_CTX(c, ctx, "b2s_halo_alloc");
  if (!ptr || nbytes <= 0) return set_error(B2S_EINVAL, "b2s_halo_alloc: nbytes=%lld ptr=%p", (long long)nbytes, (void*)ptr);
  DeviceGuard guard(c->device);
  Allocation a;
  int rc = sym_alloc(c, nbytes, &a);
  if (rc) return rc;
  c->allocs.push_back(a);
  *ptr = a.local;
  return B2S_OK;
}

extern "C" int b2s_halo_free(int64_t ctx, void* ptr) {
  B2S_CTX(c, ctx, "b2s_halo_free");
  DeviceGuard guard(c->device);
  for (size_t n = 1; n < c->allocs.size(); ++n)  // allocation 0 holds the announcement flags
    if (c->allocs[n].local == ptr) {
      cudaDeviceSynchronize();
      int rc = rdv_barrier(c, "b2s_halo_free");  // peers have stopped reading it
      sym_free(c, c->allocs[n]);
      c->allocs.erase(c->allocs.begin() + n);
      return rc;
    }
  return set_error(B2S_EINVAL, "b2s_halo_free: %p was not returned by b2s_halo_alloc on this context", ptr);
}

extern "C" int b2s_halo_peer_ptr(int64_t ctx, const void* ptr, int peer, void** peer_ptr) {
  B2S_CTX(c, ctx, "b2s_halo_peer_ptr");
  if (!peer_ptr || peer < 0 || peer >= c->world) return set_error(B2S_EINVAL, "b2s_halo_peer_ptr: peer %d of %d", peer, c->world);
  Allocation* a = find_alloc(c, ptr);
  if (!a) return set_error(B2S_EINVAL, "b2s_halo_peer_ptr: %p is not inside a b2s_halo_alloc buffer of this context", ptr);
  *peer_ptr = static_cast<char*>(a->peers[peer]) + (static_cast<const char*>(ptr) - static_cast<const char*>(a->local));
  return B2S_OK;
}

// links: HOST array [nlinks, 12] int64 -- words 0..9 as for b2s_halo_move (offsets in elements relative to `field`, the
// same on every rank: the allocation is symmetric), [10] = rank that owns the source sub-domain, [11] = destination
// sub-domain (batch index of the field, < 64; 0 when the field is not batched) in its low 16 bits.
// Optional, for the push path of the ungated exchange (all ranks must mark the two ends of a strip consistently):
//   [11] bit 32 (B2S_HALO_LINK_OUT)     the row is an OUTGOING strip: its source is on this rank, [10] is the rank that owns
//                                       the DESTINATION sub-domain (whose batch index is in the low 16 bits);
//   [11] bit 33 (B2S_HALO_LINK_PUSHED)  an incoming strip its owner pushes (the owner's table has the matching OUT row);
//   [11] bit 34 (B2S_HALO_LINK_STAGED)  a same-rank copy that must run AFTER the deliveries of rank [10] have arrived: the
//                                       owner pushed the strip, packed, into a staging area of this rank's allocation (the
//                                       OUT row's destination) and this row unpacks it into the halo.  Only the ungated
//                                       exchange uses it; it comes in addition to the PUSHED row of the same strip.
// Gated and fused exchanges always pull every incoming strip in place, marked or not.
extern "C" int b2s_halo_plan(int64_t ctx, const void* field, int elem_size, int nk, int nlinks, const int64_t* links, int* plan_out) {
  B2S_CTX(c, ctx, "b2s_halo_plan");
  if (!plan_out || !field || (elem_size != 4 && elem_size != 8) || nk <= 0 || nk > 65535 || nlinks < 0 || nlinks > 65535 || (nlinks && !links))
    return set_error(B2S_EINVAL, "b2s_halo_plan: bad arguments (elem_size=%d nk=%d nlinks=%d)", elem_size, nk, nlinks);
  Allocation* a = find_alloc(c, field);
  if (!a && c->world > 1)
    return set_error(B2S_EINVAL, "b2s_halo_plan: the field must live in a b2s_halo_alloc buffer (peers read it over NVLink)");
  const int64_t off = a ? static_cast<const char*>(field) - static_cast<const char*>(a->local) : 0;
  DeviceGuard guard(c->device);
  constexpr int64_t kOut = (int64_t)1 << 32, kPushed = (int64_t)1 << 33, kStaged = (int64_t)1 << 34;
  auto base_on = [&](int64_t rank) -> int64_t {
    const char* base = a ? static_cast<const char*>(a->peers[rank]) + off : static_cast<const char*>(field);
    return (int64_t) reinterpret_cast<intptr_t>(base);
  };
  Plan p;
  p.nk = nk, p.elem_size = elem_size, p.field = const_cast<void*>(field);
  // incoming strips: stable sort by destination sub-domain (the gates of a gated exchange open in that order), and
  // inside a sub-domain the strips read from this GPU before those read from peers (they need no announcement)
  std::vector<int> in, out, staged;
  for (int n = 0; n < nlinks; ++n) {
    const int64_t* L = links + (size_t)n * 12;
    const int64_t peer = L[10], dst_b = L[11] & 0xffff;
    if (peer < 0 || peer >= c->world) return set_error(B2S_EINVAL, "b2s_halo_plan: link %d names rank %lld of %d", n, (long long)peer, c->world);
    if (dst_b >= kGateSlots) return set_error(B2S_EINVAL, "b2s_halo_plan: link %d names destination sub-domain %lld (0 .. %d)", n, (long long)dst_b, kGateSlots - 1);
    if (L[8] <= 0 || L[9] <= 0) return set_error(B2S_EINVAL, "b2s_halo_plan: link %d has an empty strip", n);
    if ((L[11] & kOut) && peer == c->rank) return set_error(B2S_EINVAL, "b2s_halo_plan: link %d is marked outgoing but stays on rank %d", n, c->rank);
    if ((L[11] & kPushed) && peer == c->rank) return set_error(B2S_EINVAL, "b2s_halo_plan: link %d is marked pushed but its source is on this rank", n);
    for (int side = 0; side < 2; ++side) {
      const int64_t* S = L + 4 * side;  // [0] offset [1] depth stride [2] edge stride [3] level stride
      const int64_t reach = L[8] * std::llabs(S[1]) + L[9] * std::llabs(S[2]) + impl::kMaxLevelsPerUnit * std::llabs(S[3]);
      if (reach >= ((int64_t)1 << 31)) p.narrow = false;
    }
    if (L[8] * L[9] > p.max_strip) p.max_strip = (int)(L[8] * L[9]);
    ((L[11] & kOut) ? out : ((L[11] & kStaged) ? staged : in)).push_back(n);
    if (L[11] & (kOut | kPushed | kStaged)) p.mixed = true;
  }
  auto key = [&](int x) { return (links[(size_t)x * 12 + 11] & 0xffff) * 2 + (links[(size_t)x * 12 + 10] != c->rank ? 1 : 0); };
  std::stable_sort(in.begin(), in.end(), [&](int x, int y) { return key(x) < key(y); });
  p.nlinks = (int)in.size();
  std::vector<int64_t> tbl(in.size() * 12);
  std::vector<int> per_b(kGateSlots, 0);
  for (size_t n = 0; n < in.size(); ++n) {
    const int64_t* L = links + (size_t)in[n] * 12;
    const int64_t owner = L[10], dst_b = L[11] & 0xffff;
    int64_t* D = tbl.data() + n * 12;
    memcpy(D, L, 10 * sizeof(int64_t));
    D[10] = base_on(owner);
    D[11] = (owner == c->rank ? 0 : owner + 1) | (dst_b << 16);
    if (owner != c->rank) p.remote_bytes += L[8] * L[9] * nk * elem_size, p.peers |= 1ull << owner;
    per_b[dst_b] += 1;
    if (dst_b + 1 > p.nb) p.nb = (int)dst_b + 1;
  }
  for (int& v : per_b) v *= nk;  // work units (link, level) per destination sub-domain
  // host table -> device; a failure frees what this plan has allocated so far
  auto upload = [&](auto** dev, const auto* host, size_t count) -> int {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(dev), count * sizeof(**dev));
    if (e == cudaSuccess) e = cudaMemcpy(*dev, host, count * sizeof(**dev), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) return B2S_OK;
    for (void* q : {(void*)p.b_total_dev, (void*)p.links_dev, (void*)p.push_total_dev, (void*)p.rows_dev})
      if (q) cudaFree(q);
    return set_error((int)e, "b2s_halo_plan: %s", cudaGetErrorString(e));
  };
  if (int rc = upload(&p.b_total_dev, per_b.data(), (size_t)kGateSlots)) return rc;
  if (!tbl.empty())
    if (int rc = upload(&p.links_dev, tbl.data(), tbl.size())) return rc;
  if (p.mixed) {
    // mixed table: incoming strips that are not pushed (pulled: same-GPU ones first, they need no announcement), then the
    // outgoing strips (pushed), grouped by destination rank so that the delivery flags go out one rank after the other
    std::vector<int64_t> rows;
    std::vector<int> push_total(c->world, 0);
    auto add_row = [&](const int64_t* L, int64_t src_base, int64_t dst_base, int64_t peer, bool push) {
      const size_t at = rows.size();
      rows.resize(at + impl::kMixedWords, 0);
      memcpy(rows.data() + at, L, 10 * sizeof(int64_t));
      rows[at + 10] = src_base;
      rows[at + 11] = (peer == c->rank ? 0 : peer + 1) | ((L[11] & 0xffff) << 16) | ((int64_t)(push ? 1 : 0) << 40);
      rows[at + 12] = dst_base;
    };
    for (int pass = 0; pass < 2; ++pass)
      for (int n : in) {
        const int64_t* L = links + (size_t)n * 12;
        if (L[11] & kPushed) {
          if (pass == 0) p.wait_d |= 1ull << L[10];
          continue;
        }
        if ((L[10] != c->rank) != (pass == 1)) continue;
        add_row(L, base_on(L[10]), base_on(c->rank), L[10], false);
        if (L[10] != c->rank) p.wait_a |= 1ull << L[10];
      }
    std::stable_sort(out.begin(), out.end(), [&](int x, int y) { return links[(size_t)x * 12 + 10] < links[(size_t)y * 12 + 10]; });
    for (int n : out) {
      const int64_t* L = links + (size_t)n * 12;
      add_row(L, base_on(c->rank), base_on(L[10]), L[10], true);
      push_total[L[10]] += nk;
      p.wait_a |= 1ull << L[10];
    }
    p.nrows1 = (int)(rows.size() / impl::kMixedWords);
    for (int n : staged) {  // second phase: unpack what the peers delivered into this rank's staging area
      const int64_t* L = links + (size_t)n * 12;
      add_row(L, base_on(c->rank), base_on(c->rank), c->rank, false);
      p.wait_d |= 1ull << L[10];
    }
    p.nrows = (int)(rows.size() / impl::kMixedWords);
    if (int rc = upload(&p.push_total_dev, push_total.data(), (size_t)c->world)) return rc;
    if (!rows.empty())
      if (int rc = upload(&p.rows_dev, rows.data(), rows.size())) return rc;
  }
  c->plans.push_back(p);
  *plan_out = (int)c->plans.size() - 1;
  return B2S_OK;
}

extern "C" int64_t b2s_halo_plan_remote_bytes(int64_t ctx, int plan) {
  HaloCtx* c = as_ctx(ctx);
  return (c && plan >= 0 && plan < (int)c->plans.size()) ? c->plans[plan].remote_bytes : -1;
}

static impl::HaloXchg xchg_of(const HaloCtx* c, const Plan& p, int gated) {
  impl::HaloXchg X;
  X.links = p.links_dev, X.peer_flags = c->peer_flags_dev, X.b_total = p.b_total_dev, X.state = c->state, X.dst = p.field;
  X.peers = p.peers;
  X.nlinks = p.nlinks, X.nk = p.nk, X.my_rank = c->rank, X.world = c->world, X.gated = gated;
  X.single_wait = option("halo_handshake", 0) == 1 ? 1 : 0;
  return X;
}

static int launch_exchange(HaloCtx* c, int plan, int gated, cudaStream_t s) {
  if (plan < 0 || plan >= (int)c->plans.size()) return set_error(B2S_EINVAL, "b2s_halo_exchange: plan %d of %d", plan, (int)c->plans.size());
  const Plan& p = c->plans[plan];
  if (!gated && p.mixed && option("halo_push", 1) != 0) {
    // the choice must be the same on every rank (a receiver waits for the deliveries its table announces): it depends
    // only on the caller's table and on an option the caller sets on all ranks alike
    if (!p.narrow) return set_error(B2S_EUNSUPPORTED, "b2s_halo_exchange: strides beyond 32 bits; plan without outgoing links (pull only) for this field");
    impl::HaloXchg3 X;
    X.rows = p.rows_dev, X.peer_flags = c->peer_flags_dev, X.push_total = p.push_total_dev, X.state = c->state;
    X.wait_a = p.wait_a, X.wait_d = p.wait_d;
    X.nrows = p.nrows, X.nrows1 = p.nrows1, X.nk = p.nk, X.my_rank = c->rank, X.world = c->world;
    return impl::halo_exchange3_launch(p.elem_size, p.max_strip, X, s);
  }
  return impl::halo_exchange_launch(p.elem_size, p.nb, p.max_strip, xchg_of(c, p, gated), p.narrow, s);
}

// The transport step as ONE launch: halo update of the plan's field (neighbour handshake + strip copies over peer
// memory, taken in shares by the CTAs of the stencil grid) fused with fv_tp2d on the whole batch behind per-sub-domain
// gates.  q must be the plan's field (its compute cell (0,0,0): the pointer b2s_fv_tp2d takes).
template <typename T>
int b2s::impl::halo_fv_tp2d(int64_t ctx, int plan, int ni, int nj, int nk, int nb, F3<const T> crx, F3<const T> xfx, F3<const T> cry,
                            F3<const T> yfx, F2<const T> rarea, F3<T> q, F3<T> q_out, cudaStream_t s) {
  B2S_CTX(c, ctx, "b2s_halo_fv_tp2d");
  if (plan < 0 || plan >= (int)c->plans.size()) return set_error(B2S_EINVAL, "b2s_halo_fv_tp2d: plan %d of %d", plan, (int)c->plans.size());
  const Plan& p = c->plans[plan];
  B2S_ARGCHECK(q.p != nullptr && p.elem_size == (int)sizeof(T) && p.nk == nk && p.nb <= nb,
               "b2s_halo_fv_tp2d: the plan is for %d-byte elements, %d levels, %d sub-domains; the call has %d, %d, %d", p.elem_size,
               p.nk, p.nb, (int)sizeof(T), nk, nb);
  B2S_ARGCHECK(static_cast<const void*>(q.p - 3 - 3 * q.sj) == p.field, "b2s_halo_fv_tp2d: q is not the field the plan was made for");
  DeviceGuard guard(c->device);
  if (p.nlinks == 0) return set_error(B2S_EINVAL, "b2s_halo_fv_tp2d: the plan has no links; call b2s_fv_tp2d");
  return impl::fv_tp2d_fused<T>(ni, nj, nk, nb, F3<const T>{q.p, q.sj, q.sk, q.sb}, crx, xfx, cry, yfx, rarea, q_out, xchg_of(c, p, 1), s);
}
template int b2s::impl::halo_fv_tp2d<double>(int64_t, int, int, int, int, int, F3<const double>, F3<const double>, F3<const double>,
                                             F3<const double>, F2<const double>, F3<double>, F3<double>, cudaStream_t);
template int b2s::impl::halo_fv_tp2d<float>(int64_t, int, int, int, int, int, F3<const float>, F3<const float>, F3<const float>,
                                            F3<const float>, F2<const float>, F3<float>, F3<float>, cudaStream_t);

// Halo update of the plan's field on `stream` itself (no fork): handshake kernel + pull kernel (halo_variant 1 | 2: one kernel).
extern "C" int b2s_halo_exchange(int64_t ctx, int plan, void* stream) {
  B2S_CTX(c, ctx, "b2s_halo_exchange");
  DeviceGuard guard(c->device);
  return launch_exchange(c, plan, 0, static_cast<cudaStream_t>(stream));
}

// Fork: the exchange kernel runs on the context's own high-priority stream, ordered after everything already
// enqueued on `stream`; work the caller enqueues on `stream` before b2s_halo_exchange_wait runs concurrently with it.
// gated != 0: the kernel raises the gate flag (b2s_halo_gate) when the halos are complete; exactly one gated stencil
// launch must consume it before the next gated exchange.
extern "C" int b2s_halo_exchange_start(int64_t ctx, int plan, int gated, void* stream) {
  B2S_CTX(c, ctx, "b2s_halo_exchange_start");
  if (c->pending) return set_error(B2S_EINVAL, "b2s_halo_exchange_start: called twice without b2s_halo_exchange_wait");
  DeviceGuard guard(c->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  B2S_CUDA(cudaEventRecord(c->ev_fork, s), "b2s_halo_exchange_start: cudaEventRecord");
  B2S_CUDA(cudaStreamWaitEvent(c->comm, c->ev_fork, 0), "b2s_halo_exchange_start: cudaStreamWaitEvent");
  int rc = launch_exchange(c, plan, gated, c->comm);
  // the join event is recorded even after a failed launch so that a capturing stream can always be re-joined
  cudaError_t e = cudaEventRecord(c->ev_join, c->comm);
  c->pending = true;
  if (rc) return rc;
  if (e != cudaSuccess) return set_error((int)e, "b2s_halo_exchange_start: cudaEventRecord(join): %s", cudaGetErrorString(e));
  return B2S_OK;
}

extern "C" int b2s_halo_exchange_wait(int64_t ctx, void* stream) {
  B2S_CTX(c, ctx, "b2s_halo_exchange_wait");
  if (!c->pending) return B2S_OK;
  DeviceGuard guard(c->device);
  c->pending = false;
  B2S_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), c->ev_join, 0), "b2s_halo_exchange_wait: cudaStreamWaitEvent");
  return B2S_OK;
}

// Device address of the gate words (int32[66]: [b] halos of sub-domain b ready, [64] consumer CTAs done, [65] status).
extern "C" int b2s_halo_gate(int64_t ctx, int** gate) {
  B2S_CTX(c, ctx, "b2s_halo_gate");
  if (!gate) return set_error(B2S_EINVAL, "b2s_halo_gate: NULL");
  *gate = c->state + kGateOffset;
  return B2S_OK;
}

// Diagnostics, host-synchronising: device timeline (globaltimer, ns) of the LAST exchange and gated stencil --
// out[0] exchange start, [1] exchange end, [2] gate 0 opened, [3] first stencil CTA started, [4] stencil CTA 0 passed
// gate 0, [5] last stencil CTA finished.  Slots never written are 0.
extern "C" int b2s_halo_trace(int64_t ctx, int64_t* out6) {
  B2S_CTX(c, ctx, "b2s_halo_trace");
  if (!out6) return set_error(B2S_EINVAL, "b2s_halo_trace: NULL");
  DeviceGuard guard(c->device);
  B2S_CUDA(cudaDeviceSynchronize(), "b2s_halo_trace: cudaDeviceSynchronize");
  B2S_CUDA(cudaMemcpy(out6, c->state + 200, 6 * sizeof(int64_t), cudaMemcpyDeviceToHost), "b2s_halo_trace: cudaMemcpy");
  return B2S_OK;
}

// Host-synchronising: epochs completed so far; *status != 0 if a device-side wait gave up (a neighbour's announcement
// or the gate did not arrive within the bounded spin), i.e. some halo cells of an earlier exchange are wrong.
extern "C" int b2s_halo_status(int64_t ctx, int* epoch, int* status) {
  B2S_CTX(c, ctx, "b2s_halo_status");
  DeviceGuard guard(c->device);
  int h[kStateWords];
  B2S_CUDA(cudaDeviceSynchronize(), "b2s_halo_status: cudaDeviceSynchronize");
  B2S_CUDA(cudaMemcpy(h, c->state, sizeof(h), cudaMemcpyDeviceToHost), "b2s_halo_status: cudaMemcpy");
  if (epoch) *epoch = h[0];
  if (status) *status = (h[2] ? 1 : 0) | (h[kGateOffset + kGateSlots + 1] ? 2 : 0);
  return B2S_OK;
}
