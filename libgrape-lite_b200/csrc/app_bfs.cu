// app_bfs.cu — direction-optimising level-synchronous BFS.
//
// Behaviour follows examples/analytical_apps/cuda/bfs/bfs.h:25-273 (PEval
// :86-133 seeds the source, IncEval :135-271 expands one level, push or pull)
// and the CPU app bfs/bfs.h:44-212.  Result: int64 depth per inner vertex,
// unreachable = INT64_MAX (bfs.h:31,40-45).  Levels are order-independent, so
// any push/pull schedule yields bit-identical output.
//
// B200 re-design: levels kept as u32 + a visited bitmap that stays L2
// resident (2 MB at scale 24); the frontier is a bitmap consumed by the fused
// tile kernel (no O(V) compaction pass, no per-round Count() syncs: one
// control-block read per superstep); the pull step builds its result words in
// shared memory and writes each bitmap word once.
#include "apps_common.cuh"

namespace gl {
namespace {

struct OpBfsPush {
  using Meta = uint32_t;
  using W = float;
  static constexpr bool kWeighted = false;
  uint32_t* level;
  uint32_t* vis;
  uint32_t* nxt;
  uint32_t* remote;
  const uint64_t* rp;
  uint32_t ivnum;
  uint32_t next_depth;
  GL_DEV Meta assign(uint32_t) const { return 0; }
  GL_DEV void edge(uint32_t, Meta, uint32_t v, W, ScanAcc& acc) const {
    if (bit_test(vis, v)) return;        // plain (possibly stale) read first
    if (!bit_set_atomic(vis, v)) return; // somebody else won
    level[v] = next_depth;
    acc.touched++;
    if (v < ivnum) {
      bit_set_atomic(nxt, v);
      acc.next_count++;
      acc.next_edges += rp[v + 1] - rp[v];
    } else {
      bit_set_atomic(remote, v);
      acc.remote++;
    }
  }
};

__global__ void k_bfs_seed(uint32_t src, uint32_t* level, uint32_t* cur,
                           uint32_t* vis) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    level[src] = 0;
    cur[src >> 5] |= 1u << (src & 31);
    vis[src >> 5] |= 1u << (src & 31);
  }
}

// Pull step over the inner vertices (bfs.h:239-259): every unvisited,
// non-isolated inner vertex looks for a parent in the current frontier among
// its inner neighbours [rp[v], row_end[v]).
__global__ void __launch_bounds__(kTB)
k_bfs_pull(const uint64_t* __restrict__ rp, const uint64_t* __restrict__ row_end,
           const uint32_t* __restrict__ col, uint32_t ivnum,
           const uint32_t* __restrict__ cur, uint32_t* vis, uint32_t* nxt,
           const uint32_t* __restrict__ nz, uint32_t* level,
           uint32_t next_depth, ScanCtrl* ctrl) {
  __shared__ uint32_t s_v[kTileV];
  __shared__ uint32_t s_found[kTileV / 32];
  __shared__ uint32_t s_warp[kTB / 32 + 1];
  __shared__ uint32_t s_tile;
  const uint32_t ntiles = (ivnum + kTileV - 1) / kTileV;
  const uint32_t nwords = (ivnum + 31) / 32;
  ScanAcc acc;
  uint64_t scanned = 0;
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(&ctrl->tile_ticket, 1u);
    if (threadIdx.x < kTileV / 32) s_found[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= ntiles) break;
    const uint32_t widx = tile * (kTileV / 32) + (threadIdx.x >> 3);
    uint32_t word = widx < nwords ? (nz[widx] & ~vis[widx]) : 0u;
    uint32_t nib = (word >> ((threadIdx.x & 7) * 4)) & 0xFu;
    uint32_t nc;
    uint32_t off = block_excl_scan(__popc(nib), s_warp, &nc);
    if (nc == 0) continue;
    const uint32_t vbase = tile * kTileV + threadIdx.x * 4;
    while (nib) {
      uint32_t b = __ffs(nib) - 1;
      nib &= nib - 1;
      s_v[off++] = vbase + b;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nc; i += kTB) {
      const uint32_t v = s_v[i];
      const uint64_t b = rp[v];
      const uint64_t e = row_end[v];
      bool found = false;
      uint64_t p = b;
      for (; p < e; ++p) {
        uint32_t u = col[p];
        if (bit_test(cur, u)) {
          found = true;
          ++p;
          break;
        }
      }
      scanned += p - b;
      if (found) {
        level[v] = next_depth;
        atomicOr(&s_found[(v & (kTileV - 1)) >> 5], 1u << (v & 31));
        acc.next_count++;
        acc.touched++;
        acc.next_edges += rp[v + 1] - b;
      }
    }
    __syncthreads();
    if (threadIdx.x < kTileV / 32) {
      uint32_t w = s_found[threadIdx.x];
      uint32_t wi = tile * (kTileV / 32) + threadIdx.x;
      if (w && wi < nwords) {
        vis[wi] |= w;  // the word is owned by this tile during the pull
        nxt[wi] = w;
      }
    }
    __syncthreads();
  }
  flush_acc(acc, ctrl);
  unsigned long long s = warp_sum((unsigned long long) scanned);
  if (lane_id() == 0 && s) atomicAdd(&ctrl->scanned, s);
}

// Pull step over the outer vertices (bfs.h:210-223): an unvisited outer vertex
// whose reverse adjacency holds a frontier vertex takes next_depth and is
// reported to its owner.
__global__ void __launch_bounds__(kTB)
k_bfs_pull_outer(const uint64_t* __restrict__ orp, const uint32_t* __restrict__ ocol,
                 uint32_t ivnum, uint32_t ovnum, const uint32_t* __restrict__ cur,
                 uint32_t* vis, uint32_t* remote, uint32_t* level,
                 uint32_t next_depth, ScanCtrl* ctrl) {
  ScanAcc acc;
  uint64_t scanned = 0;
  for (uint32_t o = blockIdx.x * blockDim.x + threadIdx.x; o < ovnum;
       o += gridDim.x * blockDim.x) {
    const uint32_t v = ivnum + o;
    if (bit_test(vis, v)) continue;
    uint64_t b = orp[o], e = orp[o + 1], p = b;
    bool found = false;
    for (; p < e; ++p) {
      if (bit_test(cur, ocol[p])) {
        found = true;
        ++p;
        break;
      }
    }
    scanned += p - b;
    if (found) {
      level[v] = next_depth;
      bit_set_atomic(vis, v);
      bit_set_atomic(remote, v);
      acc.remote++;
      acc.touched++;
    }
  }
  flush_acc(acc, ctrl);
  unsigned long long s = warp_sum((unsigned long long) scanned);
  if (lane_id() == 0 && s) atomicAdd(&ctrl->scanned, s);
}

struct BfsPayload {
  GL_DEV ItemU32 operator()(uint32_t, uint32_t lid) const { return ItemU32{lid}; }
};
struct BfsApply {
  uint32_t* level;
  uint32_t* cur;
  uint32_t* vis;
  const uint64_t* rp;
  uint32_t depth;
  GL_DEV void operator()(const ItemU32& it, ScanAcc& acc) const {
    uint32_t v = it.lid;
    if (bit_set_atomic(vis, v)) {  // bfs.h:158-166 (curr_depth < depth[v])
      level[v] = depth;
      bit_set_atomic(cur, v);
      acc.aux++;
      acc.next_edges += rp[v + 1] - rp[v];
    }
  }
};

__global__ void k_level_to_depth(const uint32_t* level, uint32_t n, int64_t* out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint32_t l = level[i];
    out[i] = l == kInfU32 ? INT64_MAX : (int64_t) l;
  }
}

struct BfsApp : gl_app {
  uint32_t *level = nullptr, *cur = nullptr, *nxt = nullptr, *vis = nullptr,
           *remote = nullptr;
  int64_t* out64 = nullptr;
  size_t words = 0;
  uint32_t tvnum = 0;
  uint32_t curr_depth = 0;
  // frontier statistics driving the push/pull switch
  uint64_t n_f = 0, m_f = 0, visited_edges = 0, visited_cnt = 0;
  bool pulling = false;

  ~BfsApp() override {
    cudaFree(level);
    cudaFree(cur);
    cudaFree(nxt);
    cudaFree(vis);
    cudaFree(remote);
    cudaFree(out64);
  }
  size_t ResultElemBytes() const override { return sizeof(int64_t); }

  int Setup() override {
    tvnum = fv.ivnum + fv.ovnum;
    words = bm_words(tvnum) + 1;
    GL_CUDA(cudaMalloc(&level, sizeof(uint32_t) * std::max<uint32_t>(tvnum, 1)));
    GL_CUDA(cudaMalloc(&cur, sizeof(uint32_t) * words));
    GL_CUDA(cudaMalloc(&nxt, sizeof(uint32_t) * words));
    GL_CUDA(cudaMalloc(&vis, sizeof(uint32_t) * words));
    GL_CUDA(cudaMalloc(&remote, sizeof(uint32_t) * words));
    GL_CUDA(cudaMalloc(&out64, sizeof(int64_t) * std::max<uint32_t>(fv.ivnum, 1)));
    // message = bare lid (bfs.h:50-51: sizeof(vid_t) per outer vertex)
    return mm.Init(comm, fv, sizeof(ItemU32));
  }

  int Init() override {
    cudaStream_t s = eng.stream;
    GL_CUDA(cudaMemsetAsync(level, 0xFF, sizeof(uint32_t) * tvnum, s));
    GL_CUDA(cudaMemsetAsync(cur, 0, sizeof(uint32_t) * words, s));
    GL_CUDA(cudaMemsetAsync(nxt, 0, sizeof(uint32_t) * words, s));
    GL_CUDA(cudaMemsetAsync(vis, 0, sizeof(uint32_t) * words, s));
    GL_CUDA(cudaMemsetAsync(remote, 0, sizeof(uint32_t) * words, s));
    curr_depth = 0;
    n_f = m_f = visited_edges = visited_cnt = 0;
    pulling = false;
    return GL_OK;
  }

  int PEval() override {
    uint32_t src;
    if (gl_frag_oid2lid(frag, cfg.source_oid, &src) == GL_OK) {
      GL_LAUNCH(k_bfs_seed, 1, 32, eng.stream, src, level, cur, vis);
      uint64_t rp2[2];
      GL_CUDA(cudaMemcpyAsync(rp2, fv.oe_rp + src, sizeof(rp2), cudaMemcpyDeviceToHost, eng.stream));
      GL_CUDA(cudaStreamSynchronize(eng.stream));
      n_f = 1;
      m_f = rp2[1] - rp2[0];
      visited_cnt = 1;
      visited_edges = m_f;
    }
    mm.ForceContinue();
    return GL_OK;
  }

  int IncEval() override {
    cudaStream_t s = eng.stream;
    const uint32_t next_depth = curr_depth + 1;
    GL_TRY(eng.reset_ctrl());
    if (fv.fnum > 1) {
      // ParallelProcess (bfs.h:158-166)
      MsgView mv = mm.view();
      BfsApply ap{level, cur, vis, fv.oe_rp, curr_depth};
      GL_LAUNCH((k_unpack<ItemU32, BfsApply>), eng.sm_count * 4, kTB, s, mv, ap, eng.ctrl);
      GL_TRY(eng.fetch_ctrl());
      n_f += eng.h_ctrl->aux;
      m_f += eng.h_ctrl->next_edges;
      visited_cnt += eng.h_ctrl->aux;
      visited_edges += eng.h_ctrl->next_edges;
      GL_TRY(eng.reset_ctrl());
      GL_CUDA(cudaMemsetAsync(remote + (fv.ivnum >> 5), 0, sizeof(uint32_t) * (words - (fv.ivnum >> 5)), s));
    }
    // direction choice (Beamer-style on edge counts; the reference uses vertex
    // ratios, bfs.h:171-180 — both only affect speed, never the levels)
    bool use_pull = false;
    if (cfg.direction_opt && n_f > 0) {
      const uint64_t m_total = frag->oe.entries;
      const uint64_t m_u = m_total > visited_edges ? m_total - visited_edges : 0;
      if (!pulling) use_pull = m_f > m_u / 14;
      else use_pull = n_f >= (uint64_t) fv.ivnum / 24;
    }
    pulling = use_pull;
    EdgeRange er{fv.oe_rp, fv.oe_col, nullptr};
    if (!use_pull) {
      OpBfsPush op{level, vis, nxt, remote, fv.oe_rp, fv.ivnum, next_depth};
      GL_TRY(run_frontier_scan(eng, cur, fv.ivnum, er, op));
    } else {
      static thread_local int gp = 0;
      if (!gp) gp = persistent_grid(k_bfs_pull, eng.sm_count);
      if (fv.ovnum) {
        GL_LAUNCH(k_bfs_pull_outer, eng.sm_count * 8, kTB, s, fv.ovie_rp, fv.ovie_col,
                  fv.ivnum, fv.ovnum, cur, vis, remote, level, next_depth, eng.ctrl);
      }
      const uint64_t* row_end = fv.fnum > 1 ? fv.oe_split : fv.oe_rp + 1;
      GL_LAUNCH(k_bfs_pull, gp, kTB, s, fv.oe_rp, row_end, fv.oe_col, fv.ivnum, cur,
                vis, nxt, frag->nonzero_deg, level, next_depth, eng.ctrl);
    }
    if (fv.fnum > 1) {
      MsgView mv = mm.view();
      GL_LAUNCH((k_pack_outer<ItemU32, BfsPayload>), eng.sm_count * 4, kTB, s, remote,
                fv.ivnum, fv.ovnum, fv.ovgid, mv, BfsPayload(), 0, nullptr);
    }
    GL_TRY(eng.fetch_ctrl());
    const ScanCtrl& c = *eng.h_ctrl;
    note_step(c.scanned, (uint32_t) std::min<uint64_t>(n_f, 0xFFFFFFFFu), use_pull ? 1 : 0);
    q_touched += c.touched;
    n_f = c.next_count;
    m_f = c.next_edges;
    visited_cnt += c.next_count;
    visited_edges += c.next_edges;
    if (c.next_count > 0) mm.ForceContinue();
    curr_depth = next_depth;
    std::swap(cur, nxt);
    GL_CUDA(cudaMemsetAsync(nxt, 0, sizeof(uint32_t) * words, s));
    return GL_OK;
  }

  int Result(void* host_out, size_t) override {
    if (fv.ivnum == 0) return GL_OK;
    GL_LAUNCH(k_level_to_depth, (fv.ivnum + 255) / 256, 256, eng.stream, level, fv.ivnum, out64);
    GL_CUDA(cudaMemcpyAsync(host_out, out64, sizeof(int64_t) * fv.ivnum, cudaMemcpyDeviceToHost, eng.stream));
    GL_CUDA(cudaStreamSynchronize(eng.stream));
    return GL_OK;
  }
};

}  // namespace

gl_app* make_bfs() { return new BfsApp; }

}  // namespace gl
