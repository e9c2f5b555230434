// app_lcc.cu — local clustering coefficient by oriented triangle counting.
//
// Semantics of examples/analytical_apps/lcc/lcc.h:48-233 + lcc_context.h:52-66
// (GPU apps: cuda/lcc/lcc_preprocess.h:96-188 + lcc_opt.h:183-316):
//   degree(v) = CSR entries of v's row (multi-edges count);
//   N+(v) = { u in N(v) : deg u < deg v, or deg u == deg v and gid(v) > gid(u) }
//   for every v, every u in N+(v) (with multiplicity), every w in N+(u) (with
//   multiplicity): if w is a member of N+(v) then tri[u], tri[v], tri[w] += 1;
//   lcc(v) = 2*tri / (d*(d-1)), 0 when d < 2.
// Integer counting => bit-exact; the final division is fp64 as in the reference.
//
// B200 re-design: the degree-ordered DAG is built on the device (the reference
// builds it on the CPU with one std::vector per vertex and ships it through
// MPI, lcc_preprocess.h), then one warp per oriented edge (v,u) intersects
// N+(u) against the sorted N+(v) by binary search; counts use 64-bit atomics.
#include <cub/cub.cuh>

#include "apps_common.cuh"

namespace gl {
namespace {

GL_DEV bool keep_edge(uint64_t dv, uint64_t du, uint32_t v, uint32_t u) {
  return du < dv || (du == dv && v > u);
}

__global__ void k_lcc_count(const uint64_t* __restrict__ rp, const uint32_t* __restrict__ col,
                            uint32_t ivnum, uint64_t* cnt) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v <= ivnum; v += warps) {
    if (v == ivnum) {
      if (lane_id() == 0) cnt[v] = 0;
      continue;
    }
    const uint64_t b = rp[v], e = rp[v + 1], dv = e - b;
    uint32_t k = 0;
    for (uint64_t p = b + lane_id(); p < e; p += 32) {
      uint32_t u = col[p];
      uint64_t du = rp[u + 1] - rp[u];
      k += keep_edge(dv, du, v, u);
    }
    k = warp_sum(k);
    if (lane_id() == 0) cnt[v] = k;
  }
}

// order-preserving fill (rows stay sorted by lid)
__global__ void k_lcc_fill(const uint64_t* __restrict__ rp, const uint32_t* __restrict__ col,
                           uint32_t ivnum, const uint64_t* __restrict__ orp,
                           uint32_t* ocol, uint32_t* osrc) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < ivnum; v += warps) {
    const uint64_t b = rp[v], e = rp[v + 1], dv = e - b;
    uint64_t o = orp[v];
    for (uint64_t p0 = b; p0 < e; p0 += 32) {
      uint64_t p = p0 + lane_id();
      bool k = false;
      uint32_t u = 0;
      if (p < e) {
        u = col[p];
        k = keep_edge(dv, rp[u + 1] - rp[u], v, u);
      }
      uint32_t mask = __ballot_sync(0xffffffffu, k);
      if (k) {
        uint64_t at = o + __popc(mask & ((1u << lane_id()) - 1));
        ocol[at] = u;
        osrc[at] = v;
      }
      o += __popc(mask);
    }
  }
}

// one warp per oriented edge (v,u)
__global__ void __launch_bounds__(256)
k_lcc_tri(const uint64_t* __restrict__ orp, const uint32_t* __restrict__ ocol,
          const uint32_t* __restrict__ osrc, uint64_t om,
          unsigned long long* tri, ScanCtrl* ctrl) {
  const uint64_t warps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
  uint64_t scanned = 0;
  for (uint64_t ei = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5; ei < om; ei += warps) {
    const uint32_t v = osrc[ei], u = ocol[ei];
    const uint64_t vb = orp[v], ve = orp[v + 1];
    const uint64_t ub = orp[u], ue = orp[u + 1];
    uint32_t found = 0;
    for (uint64_t p = ub + lane_id(); p < ue; p += 32) {
      const uint32_t w = ocol[p];
      uint64_t lo = vb, hi = ve;
      while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (ocol[mid] < w) lo = mid + 1; else hi = mid;
      }
      if (lo < ve && ocol[lo] == w) {
        ++found;
        atomicAdd(tri + w, 1ull);
      }
    }
    found = warp_sum(found);
    if (lane_id() == 0) {
      if (found) {
        atomicAdd(tri + u, (unsigned long long) found);
        atomicAdd(tri + v, (unsigned long long) found);
      }
      scanned += ue - ub;
    }
  }
  if (lane_id() == 0 && scanned) atomicAdd(&ctrl->scanned, (unsigned long long) scanned);
}

__global__ void k_lcc_out(const uint64_t* rp, const unsigned long long* tri,
                          uint32_t ivnum, double* out) {
  uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= ivnum) return;
  int64_t d = (int64_t) (rp[v + 1] - rp[v]);
  out[v] = d < 2 ? 0.0 : 2.0 * (double) tri[v] / (double) (d * (d - 1));
}

struct LccApp : gl_app {
  uint64_t *cnt = nullptr, *orp = nullptr;
  uint32_t *ocol = nullptr, *osrc = nullptr;
  unsigned long long* tri = nullptr;
  double* out64 = nullptr;
  uint64_t om = 0;
  int stage = 0;

  ~LccApp() override {
    cudaFree(cnt);
    cudaFree(orp);
    cudaFree(ocol);
    cudaFree(osrc);
    cudaFree(tri);
    cudaFree(out64);
  }
  size_t ResultElemBytes() const override { return sizeof(double); }

  int Setup() override {
    if (fv.fnum > 1) {
      set_error("LCC on fnum > 1 needs the neighbour-list exchange of lcc.h:96-140 (next row); run it on one fragment");
      return GL_ERR_STATE;
    }
    GL_CUDA(cudaMalloc(&cnt, sizeof(uint64_t) * ((size_t) fv.ivnum + 1)));
    GL_CUDA(cudaMalloc(&orp, sizeof(uint64_t) * ((size_t) fv.ivnum + 1)));
    GL_CUDA(cudaMalloc(&tri, sizeof(unsigned long long) * std::max<uint32_t>(fv.ivnum, 1)));
    GL_CUDA(cudaMalloc(&out64, sizeof(double) * std::max<uint32_t>(fv.ivnum, 1)));
    return mm.Init(comm, fv, sizeof(ItemU32U32));
  }

  int Init() override {
    stage = 0;
    GL_CUDA(cudaMemsetAsync(tri, 0, sizeof(unsigned long long) * std::max<uint32_t>(fv.ivnum, 1), eng.stream));
    return GL_OK;
  }

  // PEval: degrees + degree-ordered DAG (lcc.h:48-74 and stage 0, :84-132)
  int PEval() override {
    cudaStream_t s = eng.stream;
    GL_LAUNCH(k_lcc_count, eng.sm_count * 8, 256, s, fv.oe_rp, fv.oe_col, fv.ivnum, cnt);
    size_t tb = 0;
    GL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, orp, (int) (fv.ivnum + 1), s));
    void* tmp = nullptr;
    GL_CUDA(cudaMalloc(&tmp, std::max<size_t>(tb, 16)));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, orp, (int) (fv.ivnum + 1), s);
    uint64_t total = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, orp + fv.ivnum, 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(tmp);
    if (e != cudaSuccess) {
      set_error("LCC scan: %s", cudaGetErrorString(e));
      return GL_ERR_CUDA;
    }
    if (total != om || !ocol) {
      cudaFree(ocol);
      cudaFree(osrc);
      ocol = osrc = nullptr;
      om = total;
      GL_CUDA(cudaMalloc(&ocol, sizeof(uint32_t) * std::max<uint64_t>(om, 4)));
      GL_CUDA(cudaMalloc(&osrc, sizeof(uint32_t) * std::max<uint64_t>(om, 4)));
    }
    if (fv.ivnum) GL_LAUNCH(k_lcc_fill, eng.sm_count * 8, 256, s, fv.oe_rp, fv.oe_col, fv.ivnum, orp, ocol, osrc);
    mm.ForceContinue();
    return GL_OK;
  }

  // IncEval: triangle counting (stage 1, lcc.h:133-196), then idle (stage 2)
  int IncEval() override {
    if (stage == 0) {
      stage = 1;
      GL_TRY(eng.reset_ctrl());
      if (om) GL_LAUNCH(k_lcc_tri, eng.sm_count * 8, 256, eng.stream, orp, ocol, osrc, om, tri, eng.ctrl);
      GL_TRY(eng.fetch_ctrl());
      note_step(eng.h_ctrl->scanned, fv.ivnum, 2);
      q_touched += fv.ivnum;
      mm.ForceContinue();
    } else {
      stage = 2;
    }
    return GL_OK;
  }

  int Result(void* host_out, size_t) override {
    if (fv.ivnum == 0) return GL_OK;
    GL_LAUNCH(k_lcc_out, (fv.ivnum + 255) / 256, 256, eng.stream, fv.oe_rp, tri, fv.ivnum, out64);
    GL_CUDA(cudaMemcpyAsync(host_out, out64, sizeof(double) * fv.ivnum, cudaMemcpyDeviceToHost, eng.stream));
    GL_CUDA(cudaStreamSynchronize(eng.stream));
    return GL_OK;
  }
};

}  // namespace

gl_app* make_lcc() { return new LccApp; }

}  // namespace gl
