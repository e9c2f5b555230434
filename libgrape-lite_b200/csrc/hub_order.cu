// hub_order.cu — degree-ordered ("hub-first") relabelling helpers.
//
// B200's L2 (126 MB) cannot hold the per-vertex state of a scale-24 graph next
// to the 2 GB edge stream, and an SM can only keep a bounded number of random
// sector misses in flight.  Ordering vertices by descending degree makes the
// state of the few thousand hubs — the target of most of the 5e8 random
// accesses of a sweep — contiguous, so those accesses hit L1/L2.  The apps
// use the order internally and translate at their result boundary; the
// fragment's public layout (lid = ascending oid) is unchanged.
#include <cub/cub.cuh>

#include "apps_common.cuh"

namespace gl {
namespace {
__global__ void k_degkey(const uint64_t* rp, uint32_t n, uint32_t* key, uint32_t* val) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint64_t dg = rp[i + 1] - rp[i];
    if (dg > 0xFFFFFFFFull) dg = 0xFFFFFFFFull;
    key[i] = 0xFFFFFFFFu - (uint32_t) dg;   // ascending key = descending degree (stable: ties by lid)
    val[i] = i;
  }
}
__global__ void k_invert(const uint32_t* order, uint32_t n, uint32_t* perm) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) perm[order[i]] = i;
}
__global__ void k_perm_deg(const uint64_t* rp, const uint32_t* order, uint32_t n, uint64_t* deg) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) deg[i] = rp[order[i] + 1] - rp[order[i]];
  if (i == n) deg[i] = 0;
}
// one warp per new row: copy the row of order[i], relabelling the neighbours
__global__ void __launch_bounds__(256)
k_perm_rows(const uint64_t* __restrict__ rp, const uint32_t* __restrict__ col,
            const uint32_t* __restrict__ order, const uint32_t* __restrict__ perm,
            uint32_t n, const uint64_t* __restrict__ rp_p, uint32_t* col_p,
            const uint32_t* __restrict__ w, uint32_t* w_p) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const uint32_t u = order[i];
    const uint64_t b = rp[u], e = rp[u + 1], o = rp_p[i];
    for (uint64_t p = b + lane_id(); p < e; p += 32) {
      const uint32_t c = col[p];
      col_p[o + (p - b)] = c < n ? perm[c] : c;
      if (w) w_p[o + (p - b)] = w[p];   // 4-byte edge data moves with its edge
    }
  }
}
}  // namespace

int build_hub_order(cudaStream_t s, const uint64_t* rp, uint32_t n, uint32_t** perm_out, uint32_t** order_out) {
  uint32_t *key = nullptr, *key2 = nullptr, *val = nullptr, *val2 = nullptr;
  GL_CUDA(cudaMalloc(&key, 4ull * std::max<uint32_t>(n, 1)));
  GL_CUDA(cudaMalloc(&key2, 4ull * std::max<uint32_t>(n, 1)));
  GL_CUDA(cudaMalloc(&val, 4ull * std::max<uint32_t>(n, 1)));
  GL_CUDA(cudaMalloc(&val2, 4ull * std::max<uint32_t>(n, 1)));
  if (n) GL_LAUNCH(k_degkey, (n + 255) / 256, 256, s, rp, n, key, val);
  size_t tb = 0;
  cub::DoubleBuffer<uint32_t> kb(key, key2), vb(val, val2);
  GL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, kb, vb, (int) n, 0, 32, s));
  void* tmp = nullptr;
  GL_CUDA(cudaMalloc(&tmp, std::max<size_t>(tb, 16)));
  if (n) GL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, kb, vb, (int) n, 0, 32, s));
  uint32_t* perm = nullptr;
  GL_CUDA(cudaMalloc(&perm, 4ull * std::max<uint32_t>(n, 1)));
  if (n) GL_LAUNCH(k_invert, (n + 255) / 256, 256, s, vb.Current(), n, perm);
  GL_CUDA(cudaStreamSynchronize(s));
  uint32_t* order = vb.Current();
  cudaFree(key);
  cudaFree(key2);
  cudaFree(vb.Alternate());
  cudaFree(tmp);
  *perm_out = perm;
  if (order_out) *order_out = order; else cudaFree(order);
  return GL_OK;
}

// rows reordered by `order`, neighbours relabelled by `perm`
int build_permuted_csr(cudaStream_t s, const uint64_t* rp, const uint32_t* col, uint64_t m, uint32_t n,
                       const uint32_t* order, const uint32_t* perm, uint64_t** rp_out, uint32_t** col_out,
                       const void* w4, void** w4_out, bool sort_rows) {
  uint64_t *deg = nullptr, *rp_p = nullptr;
  GL_CUDA(cudaMalloc(&deg, 8ull * ((size_t) n + 1)));
  GL_CUDA(cudaMalloc(&rp_p, 8ull * ((size_t) n + 1)));
  GL_LAUNCH(k_perm_deg, (n + 1 + 255) / 256, 256, s, rp, order, n, deg);
  size_t tb = 0;
  GL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, deg, rp_p, (int) (n + 1), s));
  void* tmp = nullptr;
  GL_CUDA(cudaMalloc(&tmp, std::max<size_t>(tb, 16)));
  GL_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, deg, rp_p, (int) (n + 1), s));
  uint32_t* col_p = nullptr;
  GL_CUDA(cudaMalloc(&col_p, 4ull * (m + 16)));
  uint32_t* w_p = nullptr;
  if (w4) GL_CUDA(cudaMalloc(&w_p, 4ull * (m + 16)));
  if (n) GL_LAUNCH(k_perm_rows, 148 * 8, 256, s, rp, col, order, perm, n, rp_p, col_p, (const uint32_t*) w4, w_p);
  // rows sorted by the new ids = by descending degree of the neighbour: the
  // most promising parents come first in a pull scan, and the entries of a
  // row that fall into the same bitmap word are adjacent
  if (n && m && sort_rows) {
    uint32_t* sorted = nullptr;
    GL_CUDA(cudaMalloc(&sorted, 4ull * (m + 16)));
    size_t sb = 0;
    void* st = nullptr;
    if (!w_p) {
      GL_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, sb, col_p, sorted, (int64_t) m, (int64_t) n, rp_p, rp_p + 1, s));
      GL_CUDA(cudaMalloc(&st, std::max<size_t>(sb, 16)));
      GL_CUDA(cub::DeviceSegmentedSort::SortKeys(st, sb, col_p, sorted, (int64_t) m, (int64_t) n, rp_p, rp_p + 1, s));
    } else {
      uint32_t* w_sorted = nullptr;
      GL_CUDA(cudaMalloc(&w_sorted, 4ull * (m + 16)));
      GL_CUDA(cub::DeviceSegmentedSort::SortPairs(nullptr, sb, col_p, sorted, w_p, w_sorted, (int64_t) m, (int64_t) n, rp_p, rp_p + 1, s));
      GL_CUDA(cudaMalloc(&st, std::max<size_t>(sb, 16)));
      GL_CUDA(cub::DeviceSegmentedSort::SortPairs(st, sb, col_p, sorted, w_p, w_sorted, (int64_t) m, (int64_t) n, rp_p, rp_p + 1, s));
      GL_CUDA(cudaStreamSynchronize(s));
      cudaFree(w_p);
      w_p = w_sorted;
    }
    GL_CUDA(cudaStreamSynchronize(s));
    cudaFree(st);
    cudaFree(col_p);
    col_p = sorted;
  }
  GL_CUDA(cudaStreamSynchronize(s));
  cudaFree(deg);
  cudaFree(tmp);
  *rp_out = rp_p;
  *col_out = col_p;
  if (w4_out) *w4_out = w_p;
  return GL_OK;
}

int sort_csr_rows(cudaStream_t s, const uint64_t* rp, uint32_t n, uint64_t m, uint32_t** col) {
  if (!n || !m) return GL_OK;
  uint32_t* sorted = nullptr;
  GL_CUDA(cudaMalloc(&sorted, 4ull * (m + 16)));
  size_t sb = 0;
  void* st = nullptr;
  GL_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, sb, *col, sorted, (int64_t) m, (int64_t) n, rp, rp + 1, s));
  GL_CUDA(cudaMalloc(&st, std::max<size_t>(sb, 16)));
  GL_CUDA(cub::DeviceSegmentedSort::SortKeys(st, sb, *col, sorted, (int64_t) m, (int64_t) n, rp, rp + 1, s));
  GL_CUDA(cudaStreamSynchronize(s));
  cudaFree(st);
  cudaFree(*col);
  *col = sorted;
  return GL_OK;
}

}  // namespace gl
