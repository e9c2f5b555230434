// test_batch_shuffle.cu — a PageRank written against the reference's BatchShuffle app API
// (BatchShuffleAppBase / INSTALL_GPU_BATCH_SHUFFLE_WORKER / GPUBatchShuffleWorker /
// BatchShuffleMessageManager::SyncInnerVertices, grape/cuda/{app,worker,parallel}/*batch_shuffle*):
// every round each fragment computes rank/degree of its inner vertices, SyncInnerVertices pushes
// them to the outer copies, and every inner vertex PULLS over its adjacency.  The reference ships
// no GPU app on this API; this one exercises it on 1..N fragments and must reproduce
// dataset/p2p-31-PR (checked by tests/test_gpu_compat.py).
//
// usage: test_batch_shuffle --efile F --vfile F --out_prefix DIR [--pr_d D] [--pr_mr R]
#include <sys/stat.h>

#include <cstdio>
#include <fstream>
#include <iomanip>
#include <map>
#include <string>

#include "grape/grape.h"
#include "grape/fragment/loader.h"
#include "grape/cuda/app/batch_shuffle_app_base.h"
#include "grape/cuda/fragment/host_fragment.h"
#include "grape/cuda/parallel/parallel_engine.h"
#include "grape/cuda/worker/gpu_batch_shuffle_worker.h"

#include "cuda/app_config.h"

namespace gc = grape::cuda;

template <typename FRAG_T>
class PRContext : public grape::VoidContext<FRAG_T> {
 public:
  using vid_t = typename FRAG_T::vid_t;
  explicit PRContext(const FRAG_T& frag) : grape::VoidContext<FRAG_T>(frag) {}
  void Init(gc::BatchShuffleMessageManager&, gc::AppConfig cfg, double d, int mr) {
    auto& frag = this->fragment();
    delta = d;
    max_round = mr;
    lb = cfg.lb;
    contrib.Init(frag.Vertices(), 0.0);
    rank.Init(frag.InnerVertices(), 0.0);
    next.Init(frag.InnerVertices(), 0.0);
  }
  void Output(std::ostream& os) override {
    auto& frag = this->fragment();
    rank.D2H();
    for (auto v : frag.InnerVertices())
      os << frag.GetId(v) << " " << std::scientific << std::setprecision(15) << rank[v] << std::endl;
  }
  double delta = 0.85;
  int max_round = 10, step = 0;
  gc::LoadBalancing lb{};
  gc::VertexArray<double, vid_t> contrib, rank, next;
};

template <typename FRAG_T>
class PRBatchShuffle : public gc::BatchShuffleAppBase<FRAG_T, PRContext<FRAG_T>>,
                       public gc::ParallelEngine,
                       public gc::Communicator {
 public:
  INSTALL_GPU_BATCH_SHUFFLE_WORKER(PRBatchShuffle<FRAG_T>, PRContext<FRAG_T>, FRAG_T)
  using dev_fragment_t = typename fragment_t::device_t;
  using vertex_t = typename dev_fragment_t::vertex_t;
  using nbr_t = typename dev_fragment_t::nbr_t;

  void PEval(const fragment_t& frag, context_t& ctx, message_manager_t& messages) override {
    auto iv = frag.InnerVertices();
    auto d_rank = ctx.rank.DeviceObject();
    const double p = 1.0 / (double) frag.GetTotalVerticesNum();
    gc::WorkSourceRange<vertex_t> ws(*iv.begin(), iv.size());
    ForEach(messages.stream(), ws, [=] __device__(vertex_t v) mutable { d_rank[v] = p; });
    messages.stream().Sync();
    messages.ForceContinue();
  }

  void IncEval(const fragment_t& frag, context_t& ctx, message_manager_t& messages) override {
    if (ctx.step++ >= ctx.max_round) return;   // no sync this round -> ToTerminate
    auto d_frag = frag.DeviceObject();
    auto iv = frag.InnerVertices();
    auto d_rank = ctx.rank.DeviceObject();
    auto d_next = ctx.next.DeviceObject();
    auto d_contrib = ctx.contrib.DeviceObject();
    auto& stream = messages.stream();
    gc::WorkSourceRange<vertex_t> ws(*iv.begin(), iv.size());
    // dangling mass (all-reduced) and the inner vertices' contributions
    gc::SharedValue<double> dsum;
    dsum.set(0.0, stream);
    auto* d_sum = dsum.data();
    ForEach(stream, ws, [=] __device__(vertex_t v) mutable {
      const int dg = d_frag.GetLocalOutDegree(v);
      if (dg == 0) {
        atomicAdd(d_sum, d_rank[v]);
        d_contrib[v] = 0.0;
      } else {
        d_contrib[v] = d_rank[v] / (double) dg;
      }
    });
    double local = dsum.get(stream), dangling = 0;
    Sum(local, dangling);
    // owner state -> outer copies (the dense mirror sync)
    messages.SyncInnerVertices(frag, ctx.contrib);
    const double n = (double) frag.GetTotalVerticesNum(), delta = ctx.delta;
    const double base = (1.0 - delta) / n + delta * dangling / n;
    ForEach(stream, ws, [=] __device__(vertex_t v) mutable {
      double acc = 0.0;
      for (auto& e : d_frag.GetOutgoingAdjList(v)) acc += d_contrib[e.get_neighbor()];
      d_next[v] = base + delta * acc;
    });
    stream.Sync();
    ctx.rank.Swap(ctx.next);
  }
};

int main(int argc, char** argv) {
  std::map<std::string, std::string> o = {{"efile", ""}, {"vfile", ""}, {"out_prefix", ""}, {"pr_d", "0.85"}, {"pr_mr", "10"}};
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string k = argv[i];
    if (k.rfind("--", 0) != 0 || !o.count(k.substr(2))) {
      fprintf(stderr, "unknown option %s\n", k.c_str());
      return 2;
    }
    o[k.substr(2)] = argv[i + 1];
  }
  grape::InitMPIComm();
  {
    grape::CommSpec comm_spec;
    comm_spec.Init(MPI_COMM_WORLD);
    using FRAG_T = gc::HostFragment<int64_t, uint32_t, grape::EmptyType, grape::EmptyType, grape::LoadStrategy::kOnlyOut>;
    grape::LoadGraphSpec graph_spec = grape::DefaultLoadGraphSpec();
    graph_spec.set_directed(false);
    graph_spec.set_rebalance(false, 0);
    auto fragment = grape::LoadGraph<FRAG_T>(o["efile"], o["vfile"], comm_spec, graph_spec);
    using AppType = PRBatchShuffle<FRAG_T>;
    auto app = std::make_shared<AppType>();
    auto worker = AppType::CreateWorker(app, fragment);
    gc::AppConfig cfg;
    cfg.lb = gc::ParseLoadBalancing("cm");
    cfg.wl_alloc_factor_in = 0.4;
    cfg.wl_alloc_factor_out_local = 0.2;
    cfg.wl_alloc_factor_out_remote = 0.2;
    worker->Init(comm_spec, cfg, std::stod(o["pr_d"]), std::stoi(o["pr_mr"]));
    worker->Query();
    if (!o["out_prefix"].empty()) {
      mkdir(o["out_prefix"].c_str(), 0777);
      std::ofstream os(grape::GetResultFilename(o["out_prefix"], fragment->fid()));
      worker->Output(os);
    }
    worker->Finalize();
    printf("{\"app\": \"pagerank_batch_shuffle\", \"fnum\": %d, \"supersteps\": %d}\n", (int) comm_spec.fnum(), worker->supersteps());
  }
  grape::FinalizeMPIComm();
  return 0;
}
