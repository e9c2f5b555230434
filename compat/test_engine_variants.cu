// test_engine_variants.cu — exercises the warp / CTA-per-item ForEachWithIndex*
// forms of the drop-in ParallelEngine (parallel_engine.h:93-271) with the
// functor signatures the reference's CDLP / LCC apps use.  Prints one line per
// variant; exit code 0 when every result matches the host computation.
#include <cstdio>
#include <vector>

#include "grape/cuda/parallel/parallel_engine.h"
#include "grape/cuda/utils/work_source.h"

using namespace grape::cuda;

struct Tester : public ParallelEngine {
  int failures = 0;
  size_t n;
  unsigned long long* d_out = nullptr;
  std::vector<unsigned long long> h_out;
  explicit Tester(size_t n_) : n(n_), h_out(n_) { cudaMalloc(&d_out, n * 8); }
  ~Tester() { cudaFree(d_out); }

  void check(const char* name, unsigned long long per_lane_count, const Stream& stream) {
    stream.Sync();
    cudaMemcpy(h_out.data(), d_out, n * 8, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
      // sum over lanes l of (work * 3 + l), work = 7 + i
      unsigned long long L = per_lane_count;
      unsigned long long want = (7 + i) * 3ull * L + L * (L - 1) / 2;
      bad += h_out[i] != want;
    }
    printf("%-32s %s (%zu items)\n", name, bad ? "MISMATCH" : "OK", n);
    failures += bad != 0;
    cudaMemset(d_out, 0, n * 8);
    cudaDeviceSynchronize();   // the memset runs on the legacy stream, the kernels on a non-blocking one
  }

  void run() {
    Stream stream;
    WorkSourceRange<size_t> ws(7, n);
    unsigned long long* out = d_out;
    cudaMemset(d_out, 0, n * 8);
    cudaDeviceSynchronize();

    ForEachWithIndexWarp(stream, ws, [=] __device__(size_t lane, size_t idx, size_t work) mutable {
      atomicAdd(out + idx, (unsigned long long) (work * 3 + lane));
    });
    check("ForEachWithIndexWarp", 32, stream);

    ForEachWithIndexWarpDynamic(stream, ws, [=] __device__(size_t lane, size_t idx, size_t work) mutable {
      atomicAdd(out + idx, (unsigned long long) (work * 3 + lane));
    });
    check("ForEachWithIndexWarpDynamic", 32, stream);

    ForEachWithIndexWarpShared(
        stream, ws,
        [=] __device__(uint32_t* shm, size_t lane, size_t cid, size_t csize, size_t cnum, size_t idx,
                       size_t work) mutable {
          // every warp of the CTA owns a 32-word slice of the 32 KB scratch
          uint32_t* mine = shm + (cid % (blockDim.x / csize)) * 32;
          mine[lane] = (uint32_t) lane;
          __syncwarp();
          unsigned long long s = work * 3 + mine[(lane + 1) % csize];   // a permutation of the lanes
          atomicAdd(out + idx, s);
          __syncwarp();
          (void) cnum;
        });
    check("ForEachWithIndexWarpShared", 32, stream);

    ForEachWithIndexBlock(stream, ws, [=] __device__(size_t lane, size_t idx, size_t work) mutable {
      atomicAdd(out + idx, (unsigned long long) (work * 3 + lane));
    });
    check("ForEachWithIndexBlock", 256, stream);

    ForEachWithIndexBlockDynamic(stream, ws, [=] __device__(size_t lane, size_t idx, size_t work) mutable {
      atomicAdd(out + idx, (unsigned long long) (work * 3 + lane));
    });
    check("ForEachWithIndexBlockDynamic", 256, stream);

    ForEachWithIndexBlockShared(
        stream, ws,
        [=] __device__(uint32_t* shm, size_t lane, size_t cid, size_t csize, size_t cnum, size_t idx,
                       size_t work) mutable {
          shm[lane] = (uint32_t) lane;
          __syncthreads();
          unsigned long long s = work * 3 + shm[(lane + 1) % csize];
          atomicAdd(out + idx, s);
          (void) cid;
          (void) cnum;
        });
    check("ForEachWithIndexBlockShared", 256, stream);

    // extra arguments are forwarded to the functor (parallel_engine.h:72-91)
    ForEachWithIndexWarp(
        stream, ws,
        [=] __device__(size_t lane, size_t idx, size_t work, unsigned long long mul) mutable {
          atomicAdd(out + idx, (unsigned long long) (work * mul + lane));
        },
        3ull);
    check("ForEachWithIndexWarp(+args)", 32, stream);
  }
};

int main() {
  int rc = 0;
  for (size_t n : {(size_t) 1, (size_t) 1000, (size_t) 300000}) {
    Tester t(n);
    t.run();
    rc |= t.failures;
  }
  return rc ? 1 : 0;
}
