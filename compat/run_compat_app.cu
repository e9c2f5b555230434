// run_compat_app.cu — runs the reference's UNCHANGED GPU app sources
// (examples/analytical_apps/cuda/{bfs,sssp,wcc,pagerank}/*.h) on the B200
// engine through the drop-in grape/cuda/** headers of this directory.
// Mirrors examples/analytical_apps/run_cuda_app.h:110-138,182-317 (LoadGraph
// -> CreateWorker -> Init -> Query -> Output) without gflags.
//
// usage: run_compat_app --application bfs|sssp|wcc|wcc_opt|pagerank|cdlp|lcc|lcc_basic
//        (--efile F --vfile F | --rmat SCALE,EDGEFACTOR,SEED,WEIGHTMODE)
//        [--out_prefix DIR] [--directed 0|1] [--bfs_source N|maxdeg] [--sssp_source N|maxdeg]
//        [--pr_d D] [--pr_mr R] [--cdlp_mr R] [--lb none|cm|wm|cta|strict|cmold[,more...]] [--repeat K]
// --rmat builds the fragment from bench.py's synthetic input (oracle/rmat_gen.h) through the
// reference's BasicFragmentLoader; --lb takes a list and --repeat re-runs Query() on the loaded
// fragment (fresh app + worker each time, as the reference's worker is single-shot): one JSON line
// per (lb, repetition).
//
// The SAME source also builds oracle/_ref/ref_gpu_driver: compiled against the reference's OWN
// grape/cuda/** headers (a patched copy, oracle/ref/patch_gpu_reference.py) instead of compat/ --
// the "existing kernel" baseline of BASELINE.md, timed by bench.py --ref-gpu.
#include <sys/stat.h>

#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <string>

#include "grape/grape.h"
#include "grape/fragment/basic_fragment_loader.h"
#include "grape/fragment/loader.h"
#include "grape/cuda/fragment/host_fragment.h"
#include "grape/cuda/worker/gpu_worker.h"

#include "cuda/app_config.h"
#include "cuda/bfs/bfs.h"
#include "cuda/cdlp/cdlp.h"
#include "cuda/lcc/lcc.h"
#include "cuda/lcc/lcc_directed.h"
#include "cuda/lcc/lcc_directed_opt.h"
#include "cuda/lcc/lcc_directed_preprocess.h"
#include "cuda/lcc/lcc_opt.h"
#include "cuda/lcc/lcc_preprocess.h"
#include "cuda/pagerank/pagerank.h"
#include "cuda/sssp/sssp.h"
#include "cuda/wcc/wcc.h"
#include "cuda/wcc/wcc_opt.h"

#include "rmat_gen.h"   // oracle/rmat_gen.h (the synthetic-input definition)

namespace gc = grape::cuda;

#ifdef GRAPE_CUDA_B200_COMPAT_H_
#define WORKER_SUPERSTEPS(w) ((w)->supersteps())
static const char* kImpl = "b200-compat";
#else
#define WORKER_SUPERSTEPS(w) (-1)
static const char* kImpl = "reference-gpu";
#endif

static std::vector<std::string> split_list(const std::string& s) {
  std::vector<std::string> out;
  size_t b = 0;
  while (b <= s.size()) {
    size_t e = s.find(',', b);
    if (e == std::string::npos) e = s.size();
    if (e > b) out.push_back(s.substr(b, e - b));
    b = e + 1;
  }
  return out;
}

// synthetic input (bench.py's definition) or the files; returns the chosen source oid through *maxdeg
template <typename FRAG_T>
std::shared_ptr<FRAG_T> LoadFragment(const grape::CommSpec& comm_spec, const std::map<std::string, std::string>& o,
                                     int64_t* maxdeg) {
  grape::LoadGraphSpec graph_spec = grape::DefaultLoadGraphSpec();
  graph_spec.set_directed(o.at("directed") == "1");
  graph_spec.set_rebalance(false, 0);
  *maxdeg = 0;
  if (o.at("rmat").empty()) return grape::LoadGraph<FRAG_T>(o.at("efile"), o.at("vfile"), comm_spec, graph_spec);
  using edata_t = typename FRAG_T::edata_t;
  int scale = 0, ef = 16, wmode = 0;
  unsigned long long seed = 1;
  if (sscanf(o.at("rmat").c_str(), "%d,%d,%llu,%d", &scale, &ef, &seed, &wmode) < 1 || scale < 1 || scale > 30) {
    fprintf(stderr, "bad --rmat %s\n", o.at("rmat").c_str());
    exit(2);
  }
  const int64_t n = 1ll << scale, m = (int64_t) ef << scale;
  std::vector<int64_t> src((size_t) m), dst((size_t) m);
  std::vector<float> w(wmode ? (size_t) m : 0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < m; ++i) {
    uint64_t a, b;
    rmatdef::edge((uint64_t) i, scale, seed, &a, &b);
    src[(size_t) i] = (int64_t) a;
    dst[(size_t) i] = (int64_t) b;
    if (wmode) w[(size_t) i] = rmatdef::weight((uint64_t) i, seed, wmode);
  }
  {
    std::vector<uint32_t> deg((size_t) n, 0);
    for (int64_t i = 0; i < m; ++i) {
      ++deg[(size_t) src[(size_t) i]];
      ++deg[(size_t) dst[(size_t) i]];
    }
    int64_t best = 0;
    for (int64_t v = 1; v < n; ++v)
      if (deg[(size_t) v] > deg[(size_t) best]) best = v;
    *maxdeg = best;
  }
  graph_spec.partitioner_type = grape::PartitionerType::kMapPartitioner;
  graph_spec.idxer_type = grape::IdxerType::kHashMapIdxer;
  std::shared_ptr<FRAG_T> fragment(nullptr);
  grape::BasicFragmentLoader<FRAG_T> loader(comm_spec, graph_spec);
  grape::EmptyType vd;
  if (comm_spec.worker_id() == 0) {   // one rank feeds the loader, which shuffles to the owners
    for (int64_t v = 0; v < n; ++v) loader.AddVertex(v, vd);
  }
  loader.ConstructVertices();
  if (comm_spec.worker_id() == 0) {
    for (int64_t i = 0; i < m; ++i) {
      if constexpr (std::is_same<edata_t, grape::EmptyType>::value) loader.AddEdge(src[(size_t) i], dst[(size_t) i], grape::EmptyType());
      else loader.AddEdge(src[(size_t) i], dst[(size_t) i], (edata_t) (wmode ? w[(size_t) i] : 1.0f));
    }
  }
  loader.ConstructFragment(fragment);
  return fragment;
}

// "maxdeg" sources travel as this sentinel until the graph is loaded
static constexpr int64_t kSourceMaxDeg = INT64_MIN;
template <typename T>
T fix_arg(T a, int64_t) { return a; }
inline int64_t fix_arg(int64_t a, int64_t maxdeg) { return a == kSourceMaxDeg ? maxdeg : a; }

template <typename EDATA_T, grape::LoadStrategy LS, template <class> class APP_T, typename... Args>
int CreateAndQuery(const grape::CommSpec& comm_spec, const std::map<std::string, std::string>& o,
                   const gc::AppConfig& app_config_in, Args... args) {
  using FRAG_T = gc::HostFragment<int64_t, uint32_t, grape::EmptyType, EDATA_T, LS>;
  auto t0 = std::chrono::steady_clock::now();
  int64_t maxdeg = 0;
  std::shared_ptr<FRAG_T> fragment = LoadFragment<FRAG_T>(comm_spec, o, &maxdeg);
  double load_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  using AppType = APP_T<FRAG_T>;
  const int repeat = std::max(1, std::stoi(o.at("repeat")));
  const auto lbs = split_list(o.at("lb"));
  for (size_t li = 0; li < lbs.size(); ++li) {
    gc::AppConfig app_config = app_config_in;
    app_config.lb = gc::ParseLoadBalancing(lbs[li]);
    for (int r = 0; r < repeat; ++r) {
      auto app = std::make_shared<AppType>();
      auto worker = AppType::CreateWorker(app, fragment);
      worker->Init(comm_spec, app_config, fix_arg(args, maxdeg)...);
      MPI_Barrier(comm_spec.comm());
      auto q0 = std::chrono::steady_clock::now();
      worker->Query();
      double query_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - q0).count();
      if (li + 1 == lbs.size() && r + 1 == repeat && !o.at("out_prefix").empty()) {
        mkdir(o.at("out_prefix").c_str(), 0777);
        std::ofstream os(grape::GetResultFilename(o.at("out_prefix"), fragment->fid()));
        worker->Output(os);
      }
      printf("{\"impl\": \"%s\", \"app\": \"%s\", \"lb\": \"%s\", \"rep\": %d, \"fid\": %d, \"load_s\": %.3f, "
             "\"query_ms\": %.4f, \"supersteps\": %d, \"source\": %lld}\n",
             kImpl, o.at("application").c_str(), lbs[li].c_str(), r, (int) fragment->fid(), load_s, query_ms,
             WORKER_SUPERSTEPS(worker), (long long) maxdeg);
      fflush(stdout);
      worker->Finalize();
    }
  }
  return 0;
}


// CreateAndQueryWithPreprocess (run_cuda_app.h:139-183): a CPU app (PRE_T, the
// reference's own ParallelEngine / ParallelMessageManager) prepares host arrays
// that the GPU app's context takes as extra Init arguments.
template <typename EDATA_T, grape::LoadStrategy LS, template <class> class APP_T, template <class> class PRE_T,
          typename... Args>
int CreateAndQueryWithPreprocess(const grape::CommSpec& comm_spec, const std::map<std::string, std::string>& o,
                                 const gc::AppConfig& app_config, Args... args) {
  using FRAG_T = gc::HostFragment<int64_t, uint32_t, grape::EmptyType, EDATA_T, LS>;
  auto t0 = std::chrono::steady_clock::now();
  int64_t maxdeg = 0;
  std::shared_ptr<FRAG_T> fragment = LoadFragment<FRAG_T>(comm_spec, o, &maxdeg);
  double load_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  {
    // DoPreprocess (run_cuda_app.h:78-106)
    auto pre = std::make_shared<PRE_T<FRAG_T>>();
    auto spec = grape::MultiProcessSpec(comm_spec, false);
    auto worker = PRE_T<FRAG_T>::CreateWorker(pre, fragment);
    worker->Init(comm_spec, spec);
    MPI_Barrier(comm_spec.comm());
    worker->Query(app_config, args...);   // LCCPContext::Init(messages, app_config, sorted_col, row_offset)
    MPI_Barrier(comm_spec.comm());
    std::ofstream none;
    worker->Output(none);
    worker->Finalize();
  }
  using AppType = APP_T<FRAG_T>;
  auto app = std::make_shared<AppType>();
  auto worker = AppType::CreateWorker(app, fragment);
  worker->Init(comm_spec, app_config, args...);
  auto q0 = std::chrono::steady_clock::now();
  worker->Query();
  double query_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - q0).count();
  if (!o.at("out_prefix").empty()) {
    mkdir(o.at("out_prefix").c_str(), 0777);
    std::ofstream os(grape::GetResultFilename(o.at("out_prefix"), fragment->fid()));
    worker->Output(os);
  }
  worker->Finalize();
  printf("{\"impl\": \"%s\", \"app\": \"%s\", \"lb\": \"%s\", \"rep\": 0, \"fid\": %d, \"load_s\": %.3f, \"query_ms\": %.4f, \"supersteps\": %d}\n",
         kImpl, o.at("application").c_str(), o.at("lb").c_str(), (int) fragment->fid(), load_s, query_ms, WORKER_SUPERSTEPS(worker));
  return 0;
}

int main(int argc, char** argv) {
  std::map<std::string, std::string> o = {{"application", "bfs"}, {"efile", ""},   {"vfile", ""},
                                          {"out_prefix", ""},     {"directed", "0"}, {"bfs_source", "0"},
                                          {"sssp_source", "0"},   {"pr_d", "0.85"},  {"pr_mr", "10"},
                                          {"lb", "cta"},          {"cdlp_mr", "10"},
                                          {"rmat", ""},           {"repeat", "1"},
                                          {"wl_in", "0.4"},       {"wl_out_local", "0.2"}, {"wl_out_remote", "0.2"}};
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string k = argv[i];
    if (k.rfind("--", 0) != 0 || !o.count(k.substr(2))) {
      fprintf(stderr, "unknown option %s\n", k.c_str());
      return 2;
    }
    o[k.substr(2)] = argv[i + 1];
  }
  grape::InitMPIComm();
  int rc = 0;
  {
    grape::CommSpec comm_spec;
    comm_spec.Init(MPI_COMM_WORLD);
    // run_cuda_app.h:235-240
    gc::AppConfig app_config;
    app_config.lb = gc::ParseLoadBalancing(split_list(o["lb"]).empty() ? std::string("cta") : split_list(o["lb"])[0]);
    // flags.cc:64-69 defaults; a frontier of an R-MAT graph can exceed 40 % of the vertices, which overflows
    // the reference's work lists (its BFS then faults): --wl_in 1 --wl_out_local 1 --wl_out_remote 1
    app_config.wl_alloc_factor_in = std::stod(o["wl_in"]);
    app_config.wl_alloc_factor_out_local = std::stod(o["wl_out_local"]);
    app_config.wl_alloc_factor_out_remote = std::stod(o["wl_out_remote"]);
    const std::string& a = o["application"];
    const bool directed = o["directed"] == "1";
    using grape::LoadStrategy;
    // the type choices of run_cuda_app.h:243-312
    if (a == "bfs") {
      if (directed)
        rc = CreateAndQuery<grape::EmptyType, LoadStrategy::kBothOutIn, gc::BFS>(comm_spec, o, app_config, (o["bfs_source"] == "maxdeg" ? kSourceMaxDeg : (int64_t) std::stoll(o["bfs_source"])));
      else
        rc = CreateAndQuery<grape::EmptyType, LoadStrategy::kOnlyOut, gc::BFS>(comm_spec, o, app_config, (o["bfs_source"] == "maxdeg" ? kSourceMaxDeg : (int64_t) std::stoll(o["bfs_source"])));
    } else if (a == "sssp") {
      rc = CreateAndQuery<float, LoadStrategy::kOnlyOut, gc::SSSP>(comm_spec, o, app_config, (o["sssp_source"] == "maxdeg" ? kSourceMaxDeg : (int64_t) std::stoll(o["sssp_source"])), 0);
    } else if (a == "wcc") {
      if (directed)
        rc = CreateAndQuery<grape::EmptyType, LoadStrategy::kBothOutIn, gc::WCC>(comm_spec, o, app_config);
      else
        rc = CreateAndQuery<grape::EmptyType, LoadStrategy::kOnlyOut, gc::WCC>(comm_spec, o, app_config);
    } else if (a == "wcc_opt") {
      // run_cuda_app.h:271-274 (COO fragment + union-find, gathered on fragment 0)
      rc = CreateAndQuery<grape::EmptyType, LoadStrategy::kOnlyOut, gc::WCCOpt>(comm_spec, o, app_config);
    } else if (a == "pagerank") {
      if (directed)
        rc = CreateAndQuery<grape::EmptyType, LoadStrategy::kBothOutIn, gc::Pagerank>(comm_spec, o, app_config, std::stod(o["pr_d"]), std::stoi(o["pr_mr"]));
      else
        rc = CreateAndQuery<grape::EmptyType, LoadStrategy::kOnlyOut, gc::Pagerank>(comm_spec, o, app_config, (float) std::stod(o["pr_d"]), std::stoi(o["pr_mr"]));
    } else if (a == "lcc" && !directed) {
      // run_cuda_app.h:287-303 (undirected: LCCOPT + the CPU preprocess LCCP)
      uint32_t** col = reinterpret_cast<uint32_t**>(malloc(sizeof(uint32_t*)));
      size_t** row_offset = reinterpret_cast<size_t**>(malloc(sizeof(size_t*)));
      rc = CreateAndQueryWithPreprocess<grape::EmptyType, LoadStrategy::kOnlyOut, gc::LCCOPT, grape::LCCP>(
          comm_spec, o, app_config, col, row_offset);
    } else if (a == "lcc" && directed) {
      // run_cuda_app.h:290-299 (directed: LCCDOPT + the CPU preprocess LCCDP)
      uint32_t** col = reinterpret_cast<uint32_t**>(malloc(sizeof(uint32_t*)));
      size_t** row_offset = reinterpret_cast<size_t**>(malloc(sizeof(size_t*)));
      char** weight = reinterpret_cast<char**>(malloc(sizeof(char*)));
      size_t** true_degree = reinterpret_cast<size_t**>(malloc(sizeof(size_t*)));
      rc = CreateAndQueryWithPreprocess<grape::EmptyType, LoadStrategy::kBothOutIn, gc::LCCDOPT, grape::LCCDP>(
          comm_spec, o, app_config, col, row_offset, weight, true_degree);
    } else if (a == "lcc_basic") {
      // the non-"opt" GPU LCC apps (cuda/lcc/lcc.h, lcc_directed.h): message-driven neighbour exchange
      if (directed)
        rc = CreateAndQuery<grape::EmptyType, LoadStrategy::kBothOutIn, gc::LCCD>(comm_spec, o, app_config);
      else
        rc = CreateAndQuery<grape::EmptyType, LoadStrategy::kOnlyOut, gc::LCC>(comm_spec, o, app_config);
    } else if (a == "cdlp") {
      if (directed)
        rc = CreateAndQuery<grape::EmptyType, LoadStrategy::kBothOutIn, gc::CDLP>(comm_spec, o, app_config, std::stoi(o["cdlp_mr"]));
      else
        rc = CreateAndQuery<grape::EmptyType, LoadStrategy::kOnlyOut, gc::CDLP>(comm_spec, o, app_config, std::stoi(o["cdlp_mr"]));
    } else {
      fprintf(stderr, "unknown application %s\n", a.c_str());
      rc = 2;
    }
  }
  grape::FinalizeMPIComm();
  return rc;
}
