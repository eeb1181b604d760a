// Runtime half of the C-ABI: init/finalize, thread-local error channel, options, launch counter.
// Replaces nothing arithmetic in the reference; it is the error channel and lifecycle the
// reference's void bridge lacks (SURVEY.md 8b "Return / errors", "Lifecycle").
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "../../include/b200stencil.h"
#include "common.cuh"

namespace b2s {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static int g_device = -1;  // device bound by b2s_init (-1 = not initialised)
static int g_sms = 0;
static std::mutex g_mu;
static std::map<std::string, int> g_options = {
    {"fv_variant", 0},      // 0 auto (3, else 2, else 1), 1 direct (L1/L2) kernel, 2 TMA-pipelined tile kernel, 3 TMA streaming kernel (k_fv_stream.cu)
    {"fv_small_points", 0}, // auto choice: launches below this many points use the tile kernel, 0 = 12 000 000
    {"fv_jb", 0},           // streaming kernel: rows per work item, 0 auto (column height, halved until >= 8 items per SM)
    {"fv_ti", 0},           // TMA tile width: 0 auto, 32 | 64 | 96 | 128 | 192 (k_fv_tma.cu)
    {"fv_rows", 0},         // rows per stage: 0 auto, 4 | 8
    {"fv_stages", 0},       // mbarrier ring depth: 0 auto, 2 | 3
    {"while_variant", 0},   // while_in_function: 0 auto, 1 column scan, 2 k-split (k_patterns.cu)
    {"remap_variant", 0},   // 0 auto, 1 nested (thread per column), 2 slab + cp.async, 3 slab + TMA (k_remap_slab.cu)
    {"remap_nw", 0},        // slab kernel: warps per column group, 0 auto, 8 | 16
    {"remap_cg", 0},        // slab kernel: 32-column groups per CTA, 0 auto (1 for fp64, 2 for fp32), 1 | 2
    {"remap_ppm_cols", 0},  // remap_ppm: columns per CTA, 0 auto, 16 | 32 (k_remap_ppm.cu)
    {"remap_ppm_loader", 0},// remap_ppm: 0 auto, 1 cp.async, 2 TMA
    {"fv_split_variant", 0},// fv_tp2d_split: 0 auto, 1 tile kernel, 2 streaming kernel (k_fv_split_stream.cu)
    {"fv_split_jb", 0},     // streaming kernel: rows per CTA, 0 auto (128)
    {"fv_split_rp", 0},     // streaming kernel: rows per barrier pair, 0 auto (2), 1 | 2
    {"halo_levels", 0},     // halo_move / halo_pull: levels per thread, 0 auto (1), 1 | 4 | 8 (k_halo.cu)
    {"halo_blocks_per_sm", 0}, // k_halo_exchange: resident blocks per SM of the persistent grid, 0 auto (2), 1 .. 8
    {"halo_variant", 0},       // pull exchange: 0 auto (4 alone on the stream, 1 beside a gated stencil), 1 first one-kernel version (table in global memory, 4 loads in flight), 2 (table in shared memory, 8 loads in flight), 4 = one-block handshake kernel + flat-grid pull (two launches)
    {"halo_levels_per_unit", 0}, // k_halo_exchange2 / 3: levels per work unit, 0 auto
    {"halo_handshake", 0},     // wait for the neighbours' announcements: 0 = block 0 polls the peers' flags and relays through a local word, 1 = every block polls them (first form)
    {"halo_push", 1},          // ungated exchange of a plan that carries outgoing links: 1 = strips that cross NVLink are pushed by their owner (k_halo_exchange3), 0 = everything pulled; must be the same on every rank
    {"fv_split_ti", 0},     // fv_tp2d_split tile width: 0 auto, 64 | 128 (k_fv_split.cu)
    {"sat_unroll", 0},      // levels per load batch of k_saturation_adjust: 0 auto, 1 | 2 | 4
    {"sat_kchunk", 0},      // levels per thread of k_saturation_adjust: 0 auto (8)
};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "%s: %s", what, cudaGetErrorString(e));
  return B2S_OK;
}

int sm_count() {
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    g_sms = n;
  }
  return g_sms;
}

int option(const char* name, int fallback) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_options.find(name);
  return it == g_options.end() ? fallback : it->second;
}

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return set_error(e == cudaSuccess ? B2S_ENOTINIT : static_cast<int>(e),
                     "b2s_init: no CUDA device (%s); libb200stencil has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return set_error(B2S_EINVAL, "b2s_init: device %d out of range [0,%d)", device, n);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return set_error(B2S_EUNSUPPORTED, "b2s_init: device is sm_%d%d; this library only ships sm_100a code", prop.major,
                     prop.minor);
  g_device = device;
  g_sms = prop.multiProcessorCount;
  return B2S_OK;
}

extern "C" int b2s_finalize(void) {
  g_device = -1;
  return B2S_OK;
}

extern "C" const char* b2s_last_error(void) { return g_err; }
extern "C" int b2s_device(void) { return g_device; }
extern "C" int b2s_abi_version(void) { return B2S_ABI_VERSION; }
extern "C" int b2s_sm_count(void) { return sm_count(); }
extern "C" int64_t b2s_launch_count(void) { return g_launches.load(); }

extern "C" int b2s_set_option(const char* name, int value) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_options.find(name ? name : "");
  if (it == g_options.end()) return -1;
  it->second = value;
  return 0;
}

extern "C" int b2s_get_option(const char* name) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_options.find(name ? name : "");
  return it == g_options.end() ? -1 : it->second;
}
