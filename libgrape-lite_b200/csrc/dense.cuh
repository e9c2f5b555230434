// dense.cuh — edge-balanced dense sweep: "for every row, reduce a value over
// its CSR entries" (the ForEachEdge of a WorkSourceRange over ALL vertices with
// strict load balancing, parallel_engine.h:881-979, 1311-1374).
//
// B200 design
//  * The column-index array is cut into fixed tiles of kDenseTile entries
//    (16 KB) — every CTA gets the same number of edges whatever the degree
//    distribution (a 10^5-entry R-MAT hub simply spans many tiles).
//  * Tiles (and their weight tiles) are staged into shared memory by the TMA
//    engine: 1-D bulk async copies (cp.async.bulk.shared::cluster.global with
//    mbarrier complete_tx, SASS UBLKCP), double buffered, so the next tile
//    streams in while the current one is reduced.
//  * tile_row[t] (first row touching tile t) is computed once per fragment —
//    no per-sweep prefix sums, sorted search or allocations as in the
//    reference's strict mode.
//  * A thread owns 16 consecutive entries of the tile: all 16 gathers are
//    issued first, one binary search finds its first row, and the warp
//    combines partial results of the same row with a segmented shuffle scan;
//    only segment tails touch global memory.
#pragma once
#include "engine.cuh"

namespace gl {

constexpr int kDenseTile = 4096;      // entries per tile
constexpr int kDenseRows = 1024;      // row pointers staged per tile (8 KB)

template <typename T>
struct DenseSmemW {
  T w[2][kDenseTile];
};
template <>
struct DenseSmemW<void> {};

#ifdef __CUDACC__

// tile_row[t] = row containing entry t*kDenseTile (largest r with rp[r] <= e,
// skipping empty rows); tile_row[ntiles] = nrows.
static __global__ void k_dense_tile_rows(const uint64_t* __restrict__ rp, uint32_t nrows,
                                         uint64_t m, uint32_t ntiles, uint32_t* tile_row) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  if (t == ntiles) {
    tile_row[t] = nrows;
    return;
  }
  const uint64_t e = (uint64_t) t * kDenseTile;
  // first r with rp[r] > e
  uint32_t lo = 0, hi = nrows + 1;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (rp[mid] <= e) lo = mid + 1; else hi = mid;
  }
  tile_row[t] = lo ? lo - 1 : 0;
  (void) m;
}

// Op requirements:
//   using Val = ...;  using W = float|double;  static constexpr bool kWeighted;
//   Val identity() const;
//   Val entry(uint32_t v, W w) const;                    // value of one CSR entry (row independent)
//   Val combine(Val a, Val b) const;
//   void flush(uint32_t row, Val partial, ScanAcc&) const; // atomically fold a partial into the row's state
template <class Op>
__global__ void __launch_bounds__(kTB)
k_dense_pull(const uint64_t* __restrict__ rp, const uint32_t* __restrict__ col,
             const void* __restrict__ wgt, const uint32_t* __restrict__ tile_row,
             uint32_t ntiles, uint32_t nrows, uint64_t m, Op op, ScanCtrl* ctrl) {
  using Val = typename Op::Val;
  using W = typename Op::W;
  static_assert(!Op::kWeighted || sizeof(typename Op::W) == 0,
                "weighted dense sweeps need the dynamic shared-memory variant");
  __shared__ __align__(128) uint32_t s_col[2][kDenseTile];
  __shared__ __align__(128) W s_w[Op::kWeighted ? 2 : 1][Op::kWeighted ? kDenseTile : 4];
  __shared__ uint64_t s_rp[kDenseRows + 2];
  __shared__ __align__(8) uint64_t s_bar[2];

  if (threadIdx.x == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](uint32_t tile, int stage) {
    const uint64_t e0 = (uint64_t) tile * kDenseTile;
    uint32_t n = (uint32_t) ((m - e0) < (uint64_t) kDenseTile ? (m - e0) : (uint64_t) kDenseTile);
    uint32_t bytes = ((n * 4u) + 15u) & ~15u;
    uint32_t wbytes = Op::kWeighted ? (((n * (uint32_t) sizeof(W)) + 15u) & ~15u) : 0u;
    mbar_expect_tx(&s_bar[stage], bytes + wbytes);
    tma_load_1d(&s_col[stage][0], col + e0, bytes, &s_bar[stage]);
    if (Op::kWeighted) tma_load_1d(&s_w[Op::kWeighted ? stage : 0][0], (const W*) wgt + e0, wbytes, &s_bar[stage]);
  };

  ScanAcc acc;
  uint32_t it = 0;
  uint32_t tile = blockIdx.x;
  if (tile < ntiles && threadIdx.x == 0) issue(tile, 0);
  for (; tile < ntiles; tile += gridDim.x, ++it) {
    const int stage = it & 1;
    const uint32_t next = tile + gridDim.x;
    if (next < ntiles && threadIdx.x == 0) issue(next, stage ^ 1);   // prefetch

    const uint64_t e0 = (uint64_t) tile * kDenseTile;
    const uint32_t n = (uint32_t) ((m - e0) < (uint64_t) kDenseTile ? (m - e0) : (uint64_t) kDenseTile);
    const uint32_t r0 = tile_row[tile];
    uint32_t r1 = tile_row[tile + 1];                 // row of the next tile's first entry (or nrows)
    if (r1 >= nrows) r1 = nrows - 1;
    // rows r0..r1 may own entries of this tile; stage rp[r0 .. r0+nr] (nr+1 values)
    uint32_t nr = r1 - r0 + 1;
    const bool fits = nr <= (uint32_t) kDenseRows;
    if (fits) {
      for (uint32_t i = threadIdx.x; i <= nr; i += kTB) s_rp[i] = rp[r0 + i];
    }
    __syncthreads();
    mbar_wait_parity(&s_bar[stage], (it >> 1) & 1);

    constexpr int kEPT = kDenseTile / kTB;   // 16 consecutive entries per thread
    const uint32_t le = threadIdx.x * kEPT;  // local index of this thread's first entry
    // phase A: issue every gather of this thread's 16 entries up front (16
    // independent random loads in flight per thread)
    Val ev[kEPT];
#pragma unroll
    for (int q = 0; q < kEPT / 4; ++q) {
      const uint4 c = *(const uint4*) &s_col[stage][le + 4 * q];
      const uint32_t cs[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t idx = le + 4 * q + k;
        W wv = (W) 1;
        if (Op::kWeighted) wv = s_w[Op::kWeighted ? stage : 0][idx];
        ev[4 * q + k] = (idx < n) ? op.entry(cs[k], wv) : op.identity();
      }
    }
    // phase B: ONE row search per thread, then walk its 16 entries
    uint32_t row = 0xFFFFFFFFu;
    Val part = op.identity();
    if (le < n) {
      const uint64_t e = e0 + le;
      // row of entry e: first r with rp[r] > e, minus one
      if (fits) {
        uint32_t lo = 0, hi = nr + 1;
        while (lo < hi) {
          uint32_t mid = (lo + hi) >> 1;
          if (s_rp[mid] <= e) lo = mid + 1; else hi = mid;
        }
        row = r0 + lo - 1;
      } else {
        uint32_t lo = r0, hi = r1 + 2;
        while (lo < hi) {
          uint32_t mid = (lo + hi) >> 1;
          if (rp[mid] <= e) lo = mid + 1; else hi = mid;
        }
        row = lo - 1;
      }
      uint64_t row_end = fits ? s_rp[row - r0 + 1] : rp[row + 1];
#pragma unroll
      for (int k = 0; k < kEPT; ++k) {
        if (le + k < n) {
          const uint64_t ek = e + k;
          if (ek >= row_end) {
            // the row ended inside this thread's run: fold what we have
            op.flush(row, part, acc);
            part = op.identity();
            do {
              ++row;
              row_end = fits ? s_rp[row - r0 + 1] : rp[row + 1];
            } while (ek >= row_end);
          }
          part = op.combine(part, ev[k]);
        }
      }
    }
    // warp-level segmented inclusive scan over (row, part): lanes are in entry
    // order, so equal rows are contiguous
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t orow = __shfl_up_sync(0xffffffffu, row, o);
      const Val oval = __shfl_up_sync(0xffffffffu, part, o);
      if (lane_id() >= (uint32_t) o && orow == row) part = op.combine(oval, part);
    }
    {
      const uint32_t nrow = __shfl_down_sync(0xffffffffu, row, 1);
      const bool tail = (lane_id() == 31) || (nrow != row);
      if (tail && row != 0xFFFFFFFFu) op.flush(row, part, acc);
    }
    __syncthreads();   // everybody is done with s_col[stage] / s_rp before they are refilled
  }
  flush_acc(acc, ctrl);
  if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&ctrl->scanned, (unsigned long long) m);
}

#endif  // __CUDACC__
}  // namespace gl
