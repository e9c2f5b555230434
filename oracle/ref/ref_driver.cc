// ref_driver.cc — runs the UNMODIFIED reference CPU apps (sources under
// $(REFERENCE), compiled against the shims in ./shims) on an in-memory graph.
// TEST / MEASUREMENT INFRASTRUCTURE ONLY.
//
// Mirrors examples/analytical_apps/run_app.h:103-131 (CreateAndQuery) and
// utils.h:56-82 (DoQuery): build an ImmutableEdgecutFragment through the
// reference's own BasicFragmentLoader (AddVertex / AddEdge / ConstructFragment,
// as EVFragmentLoader does, grape/fragment/ev_fragment_loader.h:118-196, minus
// the text parsing), create the worker, Init, then time worker->Query(...) —
// the reference's "run algorithm" interval — and write ctx.Output() text.
//
// usage: ref_driver --app NAME (--graph FILE.bin | --rmat SCALE,EDGEFACTOR,SEED,WEIGHTMODE)
//                   [--directed 0|1] [--source OID | --source maxdeg] [--pr_d D] [--mr R]
//                   [--threads T] [--repeat K] [--out FILE] [--opt 0|1] [--dump-edges FILE]
// --rmat generates bench.py's synthetic input here (oracle/rmat_gen.h) -- the
// CPU arm never loads the product library.  --source maxdeg picks the
// maximum-degree vertex (ties -> smallest oid) like the GPU arm.  --opt 1 runs
// the reference's tuned CPU apps (run_app_opt.h:346-420: BFSOpt, SSSPOpt,
// WCCOpt, PageRankOpt on segment partitioner + sorted-array idxer).  The JSON
// line carries "traversed_edges" (Graph500 numerator: input edges whose source
// endpoint was reached) computed from the app's own result array.
// graph file: "GRB1", int64 n, int64 m, int32 weighted, int32 has_oids,
//             [int64 oid[n]], int64 src[m], int64 dst[m], [double w[m]]
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <grape/grape.h>
#include <grape/fragment/basic_fragment_loader.h>
#include <grape/fragment/immutable_edgecut_fragment.h>

#include "../rmat_gen.h"
#include "bfs/bfs.h"
#include "bfs/bfs_opt.h"
#include "pagerank/pagerank_opt.h"
#include "sssp/sssp_opt.h"
#include "wcc/wcc_opt.h"
#include "cdlp/cdlp.h"
#include "lcc/lcc.h"
#include "pagerank/pagerank.h"
#include "pagerank/pagerank_parallel.h"
#include "pagerank/pagerank_push.h"
#include "sssp/sssp.h"
#include "wcc/wcc.h"

using namespace grape;  // NOLINT

struct InputGraph {
  int64_t n = 0, m = 0;
  int weighted = 0, has_oids = 0;
  std::vector<int64_t> oids, src, dst;
  std::vector<double> w;
};

static bool read_graph(const std::string& path, InputGraph& g) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  char magic[4];
  bool ok = fread(magic, 1, 4, f) == 4 && memcmp(magic, "GRB1", 4) == 0;
  ok = ok && fread(&g.n, 8, 1, f) == 1 && fread(&g.m, 8, 1, f) == 1;
  ok = ok && fread(&g.weighted, 4, 1, f) == 1 && fread(&g.has_oids, 4, 1, f) == 1;
  if (ok && g.has_oids) {
    g.oids.resize(g.n);
    ok = fread(g.oids.data(), 8, g.n, f) == (size_t) g.n;
  }
  if (ok) {
    g.src.resize(g.m);
    g.dst.resize(g.m);
    ok = fread(g.src.data(), 8, g.m, f) == (size_t) g.m && fread(g.dst.data(), 8, g.m, f) == (size_t) g.m;
  }
  if (ok && g.weighted) {
    g.w.resize(g.m);
    ok = fread(g.w.data(), 8, g.m, f) == (size_t) g.m;
  }
  fclose(f);
  return ok;
}

template <typename EDATA_T>
struct EdataOf {
  static EDATA_T get(const InputGraph& g, int64_t i) { return g.weighted ? (EDATA_T) g.w[i] : (EDATA_T) 1; }
};
template <>
struct EdataOf<EmptyType> {
  static EmptyType get(const InputGraph&, int64_t) { return EmptyType(); }
};

struct Options {
  std::string app, graph, out, rmat, dump;
  bool opt = false, source_maxdeg = false;
  bool directed = false;
  int64_t source = 0;
  double pr_d = 0.85;
  int mr = 10;
  int threads = 0;
  int repeat = 1;
};

// Graph500 TEPS numerator from the app's own result array (BFS / SSSP contexts
// expose partial_result; the other apps traverse every edge)
template <typename CTX_T, typename FRAG_T>
auto traversed_of(const CTX_T& ctx, const FRAG_T& frag, const InputGraph& g, int) -> decltype(ctx.partial_result, (long long) 0) {
  using val_t = typename std::decay<decltype(ctx.partial_result[typename FRAG_T::vertex_t()])>::type;
  long long cnt = 0;
#pragma omp parallel for reduction(+ : cnt)
  for (int64_t i = 0; i < g.m; ++i) {
    typename FRAG_T::vertex_t v;
    if (frag.GetInnerVertex(g.src[i], v) && ctx.partial_result[v] != std::numeric_limits<val_t>::max()) ++cnt;
  }
  return cnt;
}
template <typename CTX_T, typename FRAG_T>
long long traversed_of(const CTX_T&, const FRAG_T&, const InputGraph& g, long) {
  return (long long) g.m;
}

template <typename EDATA_T, LoadStrategy LS, template <class> class APP_T, typename... Args>
int RunApp(const CommSpec& comm_spec, const InputGraph& g, const Options& opt, Args... args) {
  using FRAG_T = ImmutableEdgecutFragment<int64_t, uint32_t, EmptyType, EDATA_T, LS>;
  auto t0 = std::chrono::steady_clock::now();
  LoadGraphSpec graph_spec = DefaultLoadGraphSpec();
  graph_spec.set_directed(opt.directed);
  graph_spec.set_rebalance(false, 0);
  // one fragment; the reference's default partitioner/idxer of run_app
  // (flags.cc:60-63: map partitioner, hashmap idxer)
  graph_spec.partitioner_type = PartitionerType::kMapPartitioner;
  graph_spec.idxer_type = IdxerType::kHashMapIdxer;
  if (opt.opt) {   // run_app_opt.h:346-362
    graph_spec.partitioner_type = PartitionerType::kSegmentedPartitioner;
    graph_spec.idxer_type = IdxerType::kSortedArrayIdxer;
  }
  std::shared_ptr<FRAG_T> fragment(nullptr);
  {
    BasicFragmentLoader<FRAG_T> loader(comm_spec, graph_spec);
    EmptyType vd;
    if (g.has_oids) {
      for (int64_t i = 0; i < g.n; ++i) loader.AddVertex(g.oids[i], vd);
    } else {
      for (int64_t i = 0; i < g.n; ++i) loader.AddVertex(i, vd);
    }
    loader.ConstructVertices();
    for (int64_t i = 0; i < g.m; ++i) loader.AddEdge(g.src[i], g.dst[i], EdataOf<EDATA_T>::get(g, i));
    loader.ConstructFragment(fragment);
  }
  double load_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  using AppType = APP_T<FRAG_T>;
  ParallelEngineSpec spec = MultiProcessSpec(comm_spec, false);  // run_app.h:186
  if (opt.threads > 0) spec.thread_num = opt.threads;
  std::vector<double> ms;
  long long traversed = 0;
  // The reference's Worker is single-shot (Query() ends with
  // messages_.Finalize()), so every repetition gets a fresh app + worker on
  // the SAME loaded fragment; only Query() is timed, as in run_app.h:188-196.
  for (int r = 0; r < opt.repeat; ++r) {
    auto app = std::make_shared<AppType>();
    auto worker = AppType::CreateWorker(app, fragment);
    worker->Init(comm_spec, spec);
    MPI_Barrier(comm_spec.comm());
    auto q0 = std::chrono::steady_clock::now();
    worker->Query(std::forward<Args>(args)...);   // == timer "run algorithm"
    ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - q0).count());
    if (r + 1 == opt.repeat) {
      traversed = traversed_of(*worker->GetContext(), *fragment, g, 0);
      if (!opt.out.empty()) {
        std::ofstream os(opt.out);
        worker->Output(os);
      }
    }
    worker->Finalize();
  }
  printf("{\"app\": \"%s\", \"opt\": %d, \"threads\": %u, \"hardware_concurrency\": %u, \"load_s\": %.3f, "
         "\"vertices\": %lld, \"edges\": %lld, \"source\": %lld, \"traversed_edges\": %lld, \"query_ms\": [",
         opt.app.c_str(), opt.opt ? 1 : 0, spec.thread_num, std::thread::hardware_concurrency(), load_s, (long long) g.n,
         (long long) g.m, (long long) opt.source, traversed);
  for (size_t i = 0; i < ms.size(); ++i) printf("%s%.4f", i ? ", " : "", ms[i]);
  printf("]}\n");
  fflush(stdout);
  return 0;
}

int main(int argc, char** argv) {
  Options opt;
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string k = argv[i], v = argv[i + 1];
    if (k == "--app") opt.app = v;
    else if (k == "--graph") opt.graph = v;
    else if (k == "--out") opt.out = v;
    else if (k == "--directed") opt.directed = atoi(v.c_str()) != 0;
    else if (k == "--source") { if (v == "maxdeg") opt.source_maxdeg = true; else opt.source = atoll(v.c_str()); }
    else if (k == "--rmat") opt.rmat = v;
    else if (k == "--opt") opt.opt = atoi(v.c_str()) != 0;
    else if (k == "--dump-edges") opt.dump = v;
    else if (k == "--pr_d") opt.pr_d = atof(v.c_str());
    else if (k == "--mr") opt.mr = atoi(v.c_str());
    else if (k == "--threads") opt.threads = atoi(v.c_str());
    else if (k == "--repeat") opt.repeat = atoi(v.c_str());
    else { fprintf(stderr, "unknown option %s\n", k.c_str()); return 2; }
  }
  InputGraph g;
  if (!opt.rmat.empty()) {
    int scale = 0, ef = 16, wmode = 0;
    unsigned long long seed = 1;
    if (sscanf(opt.rmat.c_str(), "%d,%d,%llu,%d", &scale, &ef, &seed, &wmode) < 1 || scale < 1 || scale > 30) {
      fprintf(stderr, "bad --rmat %s\n", opt.rmat.c_str());
      return 2;
    }
    g.n = 1ll << scale;
    g.m = (int64_t) ef << scale;
    g.weighted = wmode ? 1 : 0;
    g.src.resize(g.m);
    g.dst.resize(g.m);
    if (wmode) g.w.resize(g.m);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g.m; ++i) {
      uint64_t a, b;
      rmatdef::edge((uint64_t) i, scale, seed, &a, &b);
      g.src[i] = (int64_t) a;
      g.dst[i] = (int64_t) b;
      if (wmode) g.w[i] = (double) rmatdef::weight((uint64_t) i, seed, wmode);
    }
  } else if (!read_graph(opt.graph, g)) {
    fprintf(stderr, "cannot read graph %s\n", opt.graph.c_str());
    return 2;
  }
  if (!opt.dump.empty()) {   // tests/test_rmat_def.py: int64 src[m], int64 dst[m], double w[m]
    FILE* f = fopen(opt.dump.c_str(), "wb");
    if (!f) return 2;
    fwrite(g.src.data(), 8, g.m, f);
    fwrite(g.dst.data(), 8, g.m, f);
    if (g.weighted) fwrite(g.w.data(), 8, g.m, f);
    fclose(f);
    if (opt.app.empty()) return 0;
  }
  if (opt.source_maxdeg) {
    // degree = CSR entries of the undirected graph (a self loop counts twice), ties -> smallest id
    std::vector<uint32_t> deg((size_t) g.n, 0);
    for (int64_t i = 0; i < g.m; ++i) {
      ++deg[(size_t) g.src[i]];
      ++deg[(size_t) g.dst[i]];
    }
    int64_t best = 0;
    for (int64_t v = 1; v < g.n; ++v)
      if (deg[(size_t) v] > deg[(size_t) best]) best = v;
    opt.source = g.has_oids ? g.oids[(size_t) best] : best;
  }
  InitMPIComm();
  int rc = 0;
  {
    CommSpec comm_spec;
    comm_spec.Init(MPI_COMM_WORLD);
    const std::string& a = opt.app;
    // type choices = examples/analytical_apps/run_app.h:167-299
    if (opt.opt && a == "sssp") rc = RunApp<double, LoadStrategy::kOnlyOut, SSSPOpt, int64_t>(comm_spec, g, opt, opt.source);
    else if (opt.opt && a == "bfs") rc = RunApp<EmptyType, LoadStrategy::kOnlyOut, BFSOpt, int64_t>(comm_spec, g, opt, opt.source);
    else if (opt.opt && a == "wcc") rc = RunApp<EmptyType, LoadStrategy::kOnlyOut, WCCOpt>(comm_spec, g, opt);
    else if (opt.opt && a == "pagerank") rc = RunApp<EmptyType, LoadStrategy::kOnlyOut, PageRankOpt, double, int>(comm_spec, g, opt, opt.pr_d, opt.mr);
    else if (a == "sssp") rc = RunApp<double, LoadStrategy::kOnlyOut, SSSP, int64_t>(comm_spec, g, opt, opt.source);
    else if (a == "bfs") rc = RunApp<EmptyType, LoadStrategy::kOnlyOut, BFS, int64_t>(comm_spec, g, opt, opt.source);
    else if (a == "wcc") rc = RunApp<EmptyType, LoadStrategy::kOnlyOut, WCC>(comm_spec, g, opt);
    else if (a == "pagerank") rc = RunApp<EmptyType, LoadStrategy::kOnlyOut, PageRank, double, int>(comm_spec, g, opt, opt.pr_d, opt.mr);
    else if (a == "pagerank_push") rc = RunApp<EmptyType, LoadStrategy::kOnlyOut, PageRankPush, double, int>(comm_spec, g, opt, opt.pr_d, opt.mr);
    else if (a == "pagerank_parallel") rc = RunApp<EmptyType, LoadStrategy::kBothOutIn, PageRankParallel, double, int>(comm_spec, g, opt, opt.pr_d, opt.mr);
    else if (a == "cdlp") rc = RunApp<EmptyType, LoadStrategy::kOnlyOut, CDLP, int>(comm_spec, g, opt, opt.mr);
    else if (a == "lcc") rc = RunApp<EmptyType, LoadStrategy::kOnlyOut, LCC, int>(comm_spec, g, opt, std::numeric_limits<int>::max());
    else { fprintf(stderr, "unknown app %s\n", a.c_str()); rc = 2; }
  }
  FinalizeMPIComm();
  return rc;
}
