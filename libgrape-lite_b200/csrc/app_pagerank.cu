// app_pagerank.cu — PageRank, fixed number of rounds.
//
// Push formulation of examples/analytical_apps/cuda/pagerank/pagerank.h:37-252:
// PEval (:137-151) sets rank = 1/N; each IncEval (:153-249) first adds the
// partial sums received for its inner vertices, then (for max_iter rounds)
// computes the dangling mass (:183-193, all-reduced), initialises
//   next[v] = (1-d)/N + d*dangling/N                         (:196-200)
// pushes d*rank[u]/deg(u) along every out-edge with atomicAdd (:207-221) and
// ships the partial sums accumulated on outer vertices to their owners
// (:227-238).  Arithmetic is fp64 (the reference uses f32 for undirected
// graphs; the north-star tolerance of 1e-6 needs f64 accumulation).
// cfg.pr_pull = 1 selects a deterministic pull step (single fragment):
//   next[v] = base + d * sum_{u in N(v)} rank[u]/deg(u)
// which is the CPU app's formulation (pagerank/pagerank.h:102-154).
#include "apps_common.cuh"

namespace gl {
namespace {

struct OpPrPush {
  using Meta = double;
  using W = float;
  static constexpr bool kWeighted = false;
  const double* rank;
  double* next;
  const uint64_t* rp;
  double delta;
  GL_DEV Meta assign(uint32_t u) const {
    uint64_t dg = rp[u + 1] - rp[u];
    return dg ? delta * rank[u] / (double) dg : 0.0;
  }
  GL_DEV void edge(uint32_t, Meta m, uint32_t v, W, ScanAcc&) const {
    atomicAdd(next + v, m);
  }
};

__global__ void k_pr_init(double* rank, double* next, uint32_t ivnum,
                          uint32_t tvnum, double p) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < tvnum) {
    rank[i] = i < ivnum ? p : 0.0;
    next[i] = 0.0;
  }
}

__global__ void k_pr_dangling(const double* rank, const uint64_t* rp,
                              uint32_t ivnum, double* out) {
  double s = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ivnum;
       i += gridDim.x * blockDim.x)
    if (rp[i + 1] == rp[i]) s += rank[i];
  s = warp_sum(s);
  if (lane_id() == 0 && s != 0.0) atomicAdd(out, s);
}

__global__ void k_pr_base(double* next, uint32_t ivnum, double base) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ivnum) next[i] = base;
}

// contrib[u] = rank[u]/deg(u)
__global__ void k_pr_contrib(const double* rank, const uint64_t* rp,
                             uint32_t ivnum, double* contrib) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ivnum) {
    uint64_t dg = rp[i + 1] - rp[i];
    contrib[i] = dg ? rank[i] / (double) dg : 0.0;
  }
}

// pull: one warp per row chunk; lanes stride the row, fixed-order tree sum
__global__ void __launch_bounds__(kTB)
k_pr_pull(const uint64_t* __restrict__ rp, const uint32_t* __restrict__ col,
          const double* __restrict__ contrib, double* next, uint32_t ivnum,
          double base, double delta, ScanCtrl* ctrl) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  uint64_t scanned = 0;
  for (uint32_t v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < ivnum; v += warps) {
    uint64_t b = rp[v], e = rp[v + 1];
    double s = 0;
    for (uint64_t p = b + lane_id(); p < e; p += 32) s += contrib[col[p]];
    s = warp_sum(s);
    if (lane_id() == 0) {
      next[v] = base + delta * s;
      scanned += e - b;
    }
  }
  if (lane_id() == 0 && scanned) atomicAdd(&ctrl->scanned, (unsigned long long) scanned);
}

// ship partial sums of outer vertices (pagerank.h:227-238)
__global__ void __launch_bounds__(kTB)
k_pr_send(double* next, uint32_t ivnum, uint32_t ovnum,
          const uint32_t* __restrict__ ovgid, MsgView mv) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t rounds = (ovnum + stride - 1) / stride;
  uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  for (uint32_t r = 0; r < rounds; ++r, o += stride) {
    bool pred = false;
    uint32_t dst = 0;
    ItemU32F64 it{0, 0, 0.0};
    if (o < ovnum) {
      double x = next[ivnum + o];
      if (x > 0) {
        uint32_t gid = ovgid[o];
        dst = gid >> mv.fid_offset;
        it.lid = gid & mv.id_mask;
        it.val = x;
        next[ivnum + o] = 0.0;
        pred = true;
      }
    }
    msg_send<ItemU32F64>(mv, pred, dst, it);
  }
}
struct PrApply {
  double* rank;
  GL_DEV void operator()(const ItemU32F64& it, ScanAcc&) const {
    atomicAdd(rank + it.lid, it.val);
  }
};

__global__ void k_ones(uint32_t* bm, uint32_t nbits, uint32_t words) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= words) return;
  uint32_t lo = i * 32, w = 0;
  if (lo + 32 <= nbits) w = 0xFFFFFFFFu;
  else if (lo < nbits) w = (1u << (nbits - lo)) - 1u;
  bm[i] = w;
}

struct PageRankApp : gl_app {
  double *rank = nullptr, *next = nullptr, *contrib = nullptr, *d_dangling = nullptr;
  uint32_t* all_inner = nullptr;
  size_t words = 0;
  uint32_t tvnum = 0;
  int curr_iter = 0;

  ~PageRankApp() override {
    cudaFree(rank);
    cudaFree(next);
    cudaFree(contrib);
    cudaFree(d_dangling);
    cudaFree(all_inner);
  }
  size_t ResultElemBytes() const override { return sizeof(double); }

  int Setup() override {
    tvnum = fv.ivnum + fv.ovnum;
    words = bm_words(fv.ivnum) + 1;
    GL_CUDA(cudaMalloc(&rank, sizeof(double) * std::max<uint32_t>(tvnum, 1)));
    GL_CUDA(cudaMalloc(&next, sizeof(double) * std::max<uint32_t>(tvnum, 1)));
    GL_CUDA(cudaMalloc(&d_dangling, sizeof(double)));
    GL_CUDA(cudaMalloc(&all_inner, sizeof(uint32_t) * words));
    GL_LAUNCH(k_ones, (unsigned) ((words + 255) / 256), 256, eng.stream, all_inner, fv.ivnum, (uint32_t) words);
    if (cfg.pr_pull) {
      GL_CUDA(cudaMalloc(&contrib, sizeof(double) * std::max<uint32_t>(tvnum, 1)));
    }
    GL_TRY(mm.Init(comm, fv, sizeof(ItemU32F64)));
    if (cfg.pr_pull && fv.fnum > 1) GL_TRY(mm.BuildMirrorPlan(eng.stream, fv));
    return GL_OK;
  }

  int Init() override {
    curr_iter = 0;
    return GL_OK;
  }

  int PEval() override {
    double p = 1.0 / (double) fv.total_vnum;
    if (tvnum) GL_LAUNCH(k_pr_init, (tvnum + 255) / 256, 256, eng.stream, rank, next, fv.ivnum, tvnum, p);
    mm.ForceContinue();
    return GL_OK;
  }

  int IncEval() override {
    cudaStream_t s = eng.stream;
    if (fv.fnum > 1) {
      MsgView mv = mm.view();
      PrApply ap{rank};
      GL_LAUNCH((k_unpack<ItemU32F64, PrApply>), eng.sm_count * 4, kTB, s, mv, ap, eng.ctrl);
    }
    if (curr_iter++ >= cfg.max_round) return GL_OK;
    mm.ForceContinue();
    GL_TRY(eng.reset_ctrl());
    // dangling mass (pagerank.h:183-193)
    double dangling = 0;
    GL_CUDA(cudaMemsetAsync(d_dangling, 0, sizeof(double), s));
    if (fv.ivnum) GL_LAUNCH(k_pr_dangling, eng.sm_count * 4, 256, s, rank, fv.oe_rp, fv.ivnum, d_dangling);
    GL_CUDA(cudaMemcpyAsync(&dangling, d_dangling, sizeof(double), cudaMemcpyDeviceToHost, s));
    GL_CUDA(cudaStreamSynchronize(s));
    GL_TRY(mm.AllReduceF64(&dangling, 1, 0));
    const double N = (double) fv.total_vnum;
    const double base = (1.0 - cfg.pr_delta) / N + cfg.pr_delta * dangling / N;
    if (cfg.pr_pull) {
      static thread_local int gp = 0;
      if (!gp) gp = persistent_grid(k_pr_pull, eng.sm_count);
      if (fv.ivnum) {
        GL_LAUNCH(k_pr_contrib, (fv.ivnum + 255) / 256, 256, s, rank, fv.oe_rp, fv.ivnum, contrib);
      }
      // outer copies take their owner's contribution (dense mirror sync)
      if (fv.fnum > 1) GL_TRY(mm.SyncValuesToGhosts(s, contrib, 8));
      if (fv.ivnum) {
        GL_LAUNCH(k_pr_pull, gp, kTB, s, fv.oe_rp, fv.oe_col, contrib, next, fv.ivnum, base, cfg.pr_delta, eng.ctrl);
      }
    } else {
      if (fv.ivnum) GL_LAUNCH(k_pr_base, (fv.ivnum + 255) / 256, 256, s, next, fv.ivnum, base);
      OpPrPush op{rank, next, fv.oe_rp, cfg.pr_delta};
      EdgeRange er{fv.oe_rp, fv.oe_col, nullptr};
      GL_TRY(run_frontier_scan(eng, all_inner, fv.ivnum, er, op));
      if (fv.fnum > 1) {
        MsgView mv = mm.view();
        GL_LAUNCH(k_pr_send, eng.sm_count * 4, kTB, s, next, fv.ivnum, fv.ovnum, fv.ovgid, mv);
      }
    }
    GL_TRY(eng.fetch_ctrl());
    note_step(eng.h_ctrl->scanned, fv.ivnum, 2);
    q_touched += fv.ivnum;
    std::swap(rank, next);
    return GL_OK;
  }

  int Result(void* host_out, size_t) override {
    if (fv.ivnum == 0) return GL_OK;
    GL_CUDA(cudaMemcpyAsync(host_out, rank, sizeof(double) * fv.ivnum, cudaMemcpyDeviceToHost, eng.stream));
    GL_CUDA(cudaStreamSynchronize(eng.stream));
    return GL_OK;
  }
};

}  // namespace

gl_app* make_pagerank() { return new PageRankApp; }

}  // namespace gl
