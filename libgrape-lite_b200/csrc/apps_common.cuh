// apps_common.cuh — launch helpers and small kernels shared by the apps.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdlib>

#include "app_base.h"

namespace gl {

// persistent-grid size for a kernel: SMs x resident CTAs (queried once)
template <typename K>
int persistent_grid(K kernel, int sm_count, int threads = kTB) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  return per_sm * sm_count;
}

// GL_TRACE=1 in the environment: synchronise after every launch and print its
// wall time (debugging aid; never enabled in measurements)
inline bool trace_on() {
  static int on = -1;
  if (on < 0) on = getenv("GL_TRACE") ? 1 : 0;
  return on == 1;
}
#define GL_LAUNCH(kernel, grid, block, stream, ...)                                   \
  do {                                                                                \
    if (::gl::trace_on()) {                                                           \
      cudaStreamSynchronize(stream);                                                  \
      auto t0__ = std::chrono::steady_clock::now();                                   \
      kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);                          \
      cudaStreamSynchronize(stream);                                                  \
      fprintf(stderr, "[gl-trace] %-60s %8.1f us\n", #kernel,                         \
              std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0__).count()); \
    } else {                                                                          \
      kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);                          \
    }                                                                                 \
    GL_COUNT_LAUNCH();                                                                \
    GL_CUDA(cudaGetLastError());                                                      \
  } while (0)

// Runs one frontier-driven edge scan (tile kernel + hub kernel) on `er`.
template <class Op>
int run_frontier_scan(Engine& eng, const uint32_t* frontier, uint32_t nverts,
                      EdgeRange er, const Op& op) {
  static thread_local int g1 = 0, g2 = 0;
  if (!g1) g1 = persistent_grid(k_frontier_scan<Op>, eng.sm_count);
  if (!g2) g2 = persistent_grid(k_hub_scan<Op>, eng.sm_count);
  uint32_t ntiles = (nverts + kTileV - 1) / kTileV;
  int grid1 = (int) std::min<uint32_t>((uint32_t) g1, std::max<uint32_t>(ntiles, 1));
  GL_LAUNCH(k_frontier_scan<Op>, grid1, kTB, eng.stream, frontier, nverts, er,
            op, eng.ctrl, eng.hubs, eng.hub_cap, eng.hub_deg);
  GL_LAUNCH(k_hub_scan<Op>, g2, kTB, eng.stream, er, op, eng.ctrl, eng.hubs,
            eng.hub_cap);
  return GL_OK;
}

#ifdef __CUDACC__
// owner fragment of an outer vertex from its gid
GL_DEV uint32_t gid_fid(uint32_t gid, int fid_offset) { return gid >> fid_offset; }

// Generic producer: for every set bit v in `remote` over the outer range
// [ivnum, tvnum) send Item{lid_at_owner, payload(v)} to the owner.
// Replaces the "ForEach over outer vertices + SyncStateOnOuterVertexWarpOpt"
// idiom (e.g. cuda/sssp/sssp.h:295-304).
template <typename Item, class Payload>
__global__ void __launch_bounds__(kTB)
k_pack_outer(const uint32_t* __restrict__ remote, uint32_t ivnum, uint32_t ovnum,
             const uint32_t* __restrict__ ovgid, MsgView mv, Payload pay,
             int clear_bits, uint32_t* remote_rw) {
  // word-level scan of the outer part of the bitmap: a warp takes 32 words
  // (1024 outer copies), skips all-zero groups, and for every non-zero word
  // lets lane b test bit b, so msg_send runs warp-converged.
  const uint32_t w_lo = ivnum >> 5;
  const uint32_t w_hi = (ivnum + ovnum + 31) >> 5;          // exclusive
  const uint32_t nwords = w_hi > w_lo ? w_hi - w_lo : 0;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t groups = (nwords + 31) >> 5;
  for (uint32_t grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; grp < groups; grp += warps) {
    const uint32_t wi = w_lo + (grp << 5) + lane_id();
    const uint32_t word = wi < w_hi ? remote[wi] : 0u;
    uint32_t nzmask = __ballot_sync(0xffffffffu, word != 0);
    while (nzmask) {
      const uint32_t src = __ffs(nzmask) - 1;
      nzmask &= nzmask - 1;
      const uint32_t w = __shfl_sync(0xffffffffu, word, src);
      const uint32_t v = ((w_lo + (grp << 5) + src) << 5) + lane_id();
      bool pred = ((w >> lane_id()) & 1u) && v >= ivnum && v < ivnum + ovnum;
      uint32_t dst = 0;
      Item it;
      if (pred) {
        const uint32_t gid = ovgid[v - ivnum];
        dst = gid >> mv.fid_offset;
        it = pay(v, gid & mv.id_mask);
      }
      msg_send<Item>(mv, pred, dst, it);
    }
  }
  (void) clear_bits;
  (void) remote_rw;
}

// Generic consumer: apply every received item (ParallelProcess,
// gpu_message_manager.h:362-393 + message_kernels.h:28-127).
template <typename Item, class Apply>
__global__ void __launch_bounds__(kTB)
k_unpack(MsgView mv, Apply apply, ScanCtrl* ctrl) {
  ScanAcc acc;
  for (uint32_t src = 0; src < mv.fnum; ++src) {
    if (src == mv.fid) continue;
    const uint32_t n = mv.recv_count[src];
    const Item* items = (const Item*) mv.recv_slot[src];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += gridDim.x * blockDim.x)
      apply(items[i], acc);
  }
  flush_acc(acc, ctrl);
}
#endif

struct ItemU32 {
  uint32_t lid;
};
struct ItemU32U32 {
  uint32_t lid, val;
};
struct ItemU32F32 {
  uint32_t lid;
  float val;
};
struct ItemU32F64 {
  uint32_t lid, pad;
  double val;
};
struct ItemU32I64 {
  uint32_t lid, pad;
  int64_t val;
};

inline size_t bm_words(uint64_t bits) { return (size_t) ((bits + 31) / 32); }

}  // namespace gl
