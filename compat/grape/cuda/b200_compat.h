// b200_compat.h — source-level drop-in for the reference's grape/cuda/** API.
//
// The reference's GPU apps (examples/analytical_apps/cuda/{bfs,sssp,wcc,
// pagerank}/*.h) are written against grape::cuda::{GPUAppBase, GPUWorker,
// ParallelEngine, HostFragment, dev::DeviceFragment, GPUMessageManager,
// VertexArray, DenseVertexSet, Queue, WorkSource*, LaunchKernel, ...}.
// This header re-implements those names on top of the B200 engine so that the
// app sources compile UNCHANGED:
//   * storage          -> include/grape_b200.h (gl_frag_create, SoA CSR)
//   * ForEach*Edge     -> the kernel skeletons of libgrape-lite_b200/csrc/engine.cuh
//                         instantiated with the app's device lambdas
//   * containers       -> thin device bitmaps / queues
// Each class cites the reference header whose public interface it mirrors; the
// implementations are new.  Scope of this round: one fragment per process
// (the container has no MPI; the multi-fragment data plane is exercised through
// the C-ABI apps).
#ifndef GRAPE_CUDA_B200_COMPAT_H_
#define GRAPE_CUDA_B200_COMPAT_H_
#ifdef __CUDACC__

#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>
#include <thrust/pair.h>
#include <thrust/binary_search.h>
#include <thrust/execution_policy.h>

#include <algorithm>
#include <iomanip>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "grape/app/context_base.h"
#include "grape/app/void_context.h"
#include "grape/config.h"
#include "grape/fragment/immutable_edgecut_fragment.h"
#include "grape/graph/adj_list.h"
#include "grape/graph/vertex.h"
#include "grape/parallel/parallel_engine_spec.h"
#include "grape/types.h"
#include "grape/util.h"
#include "grape/utils/vertex_array.h"
#include "grape/worker/comm_spec.h"

#include "grape_b200.h"                                  // C ABI
#include "../../../libgrape-lite_b200/csrc/engine.cuh"   // kernel skeletons

#define CHECK_CUDA(err)                                                        \
  do {                                                                         \
    cudaError_t errr = (err);                                                  \
    if (errr != cudaSuccess) {                                                 \
      LOG(FATAL) << "CUDA error " << cudaGetErrorString(errr) << " at "        \
                 << __FILE__ << ":" << __LINE__;                               \
    }                                                                          \
  } while (0)
#define CHECK_GL(expr)                                                         \
  do {                                                                         \
    int st__ = (expr);                                                         \
    if (st__ != GL_OK) LOG(FATAL) << "grape_b200: " << gl_last_error();        \
  } while (0)

namespace grape {
namespace cuda {

// ------------------------------------------------------------------ stream --
// grape/cuda/utils/stream.h:22-80
enum class StreamPriority { kDefault, kHigh, kLow };
class Stream {
 public:
  explicit Stream(StreamPriority = StreamPriority::kDefault) {
    CHECK_CUDA(cudaStreamCreateWithFlags(&s_, cudaStreamNonBlocking));
  }
  Stream(const Stream&) = delete;
  Stream& operator=(const Stream&) = delete;
  Stream(Stream&& o) noexcept : s_(o.s_) { o.s_ = nullptr; }
  ~Stream() {
    if (s_) cudaStreamDestroy(s_);
  }
  void Sync() const { CHECK_CUDA(cudaStreamSynchronize(s_)); }
  cudaStream_t cuda_stream() const { return s_; }

 private:
  cudaStream_t s_ = nullptr;
};

inline void ReportMemoryUsage(const std::string& marker) {
  size_t free_b, total_b;
  CHECK_CUDA(cudaMemGetInfo(&free_b, &total_b));
  VLOG(1) << marker << ", GPU memory used: " << (total_b - free_b) / 1048576.0 << " MB";
}

// ------------------------------------------------------- small utilities --
// grape/cuda/utils/{array_view,cuda_utils(pinned_vector),device_buffer,event,
// shared_array,sorted_search,markers}.h — same public names, new bodies.
template <typename T>
struct PinnedAllocator {   // page-locked host memory (cuda_utils.h pinned_vector)
  using value_type = T;
  PinnedAllocator() = default;
  template <typename U>
  PinnedAllocator(const PinnedAllocator<U>&) {}
  T* allocate(size_t n) {
    void* p = nullptr;
    CHECK_CUDA(cudaMallocHost(&p, std::max<size_t>(n, 1) * sizeof(T)));
    return static_cast<T*>(p);
  }
  void deallocate(T* p, size_t) { cudaFreeHost(p); }
  template <typename U>
  bool operator==(const PinnedAllocator<U>&) const { return true; }
  template <typename U>
  bool operator!=(const PinnedAllocator<U>&) const { return false; }
};
template <typename T>
using pinned_vector = thrust::host_vector<T, PinnedAllocator<T>>;

// array_view.h:23-67: non-owning (pointer, size) over device / pinned storage
template <typename T>
class ArrayView {
 public:
  ArrayView() = default;
  explicit ArrayView(const thrust::device_vector<T>& v)
      : data_(const_cast<T*>(thrust::raw_pointer_cast(v.data()))), size_(v.size()) {}
  explicit ArrayView(const pinned_vector<T>& v)
      : data_(const_cast<T*>(thrust::raw_pointer_cast(v.data()))), size_(v.size()) {}
  DEV_HOST ArrayView(T* data, size_t size) : data_(data), size_(size) {}
  DEV_HOST_INLINE T* data() { return data_; }
  DEV_HOST_INLINE const T* data() const { return data_; }
  DEV_HOST_INLINE size_t size() const { return size_; }
  DEV_HOST_INLINE bool empty() const { return size_ == 0; }
  DEV_INLINE T& operator[](size_t i) { return data_[i]; }
  DEV_INLINE const T& operator[](size_t i) const { return data_[i]; }
  DEV_INLINE void Swap(ArrayView<T>& rhs) {
    T* d = data_;
    data_ = rhs.data_;
    rhs.data_ = d;
    size_t n = size_;
    size_ = rhs.size_;
    rhs.size_ = n;
  }
  DEV_INLINE T* begin() { return data_; }
  DEV_INLINE T* end() { return data_ + size_; }
  DEV_INLINE const T* begin() const { return data_; }
  DEV_INLINE const T* end() const { return data_ + size_; }

 private:
  T* data_ = nullptr;
  size_t size_ = 0;
};

// device_buffer.h:28-92: grow-only device array (resize keeps the allocation)
template <typename T>
class DeviceBuffer {
 public:
  DeviceBuffer() = default;
  explicit DeviceBuffer(size_t size) { resize(size); }
  DeviceBuffer(const DeviceBuffer& rhs) { *this = rhs; }
  DeviceBuffer(DeviceBuffer&& rhs) noexcept { *this = std::move(rhs); }
  ~DeviceBuffer() {
    if (data_) cudaFree(data_);
  }
  DeviceBuffer& operator=(const DeviceBuffer& rhs) {
    if (&rhs != this) {
      resize(rhs.size_);
      if (size_) CHECK_CUDA(cudaMemcpy(data_, rhs.data_, sizeof(T) * size_, cudaMemcpyDeviceToDevice));
    }
    return *this;
  }
  DeviceBuffer& operator=(DeviceBuffer&& rhs) noexcept {
    if (&rhs != this) {
      if (data_) cudaFree(data_);
      data_ = rhs.data_;
      size_ = rhs.size_;
      capacity_ = rhs.capacity_;
      rhs.data_ = nullptr;
      rhs.size_ = rhs.capacity_ = 0;
    }
    return *this;
  }
  void resize(size_t size) {
    if (size > capacity_) {
      T* n = nullptr;
      CHECK_CUDA(cudaMalloc(&n, sizeof(T) * size));
      if (size_) CHECK_CUDA(cudaMemcpy(n, data_, sizeof(T) * size_, cudaMemcpyDeviceToDevice));
      if (data_) cudaFree(data_);
      data_ = n;
      capacity_ = size;
    }
    size_ = size;
  }
  T* data() { return data_; }
  const T* data() const { return data_; }
  size_t size() const { return size_; }
  ArrayView<T> DeviceObject() { return ArrayView<T>(data_, size_); }

 private:
  T* data_ = nullptr;
  size_t size_ = 0, capacity_ = 0;
};

// event.h:60-125: a recordable / waitable CUDA event with shared ownership
class Event {
 public:
  Event() = default;
  static Event Create() {
    Event e;
    cudaEvent_t ev;
    CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e.ev_ = std::shared_ptr<CUevent_st>(ev, [](cudaEvent_t x) { cudaEventDestroy(x); });
    return e;
  }
  void Record(const Stream& stream) const {
    if (ev_) CHECK_CUDA(cudaEventRecord(ev_.get(), stream.cuda_stream()));
  }
  void Wait(const Stream& stream) const {
    if (ev_) CHECK_CUDA(cudaStreamWaitEvent(stream.cuda_stream(), ev_.get(), 0));
  }
  void Sync() const {
    if (ev_) CHECK_CUDA(cudaEventSynchronize(ev_.get()));
  }
  bool Query() const { return !ev_ || cudaEventQuery(ev_.get()) == cudaSuccess; }
  cudaEvent_t cuda_event() const { return ev_.get(); }

 private:
  std::shared_ptr<CUevent_st> ev_;
};

// shared_array.h:24-140: a device array with a pinned host mirror
template <typename T>
class SharedArray {
 public:
  using device_t = thrust::device_vector<T>;
  using host_t = pinned_vector<T>;
  SharedArray() = default;
  explicit SharedArray(size_t size) { resize(size); }
  void resize(size_t size) {
    d_.resize(size);
    h_.resize(size);
  }
  size_t size() const { return d_.size(); }
  void set(size_t idx, const T& t) { d_[idx] = t; }
  void set(size_t idx, const T& t, const Stream& stream) {
    h_[idx] = t;
    CHECK_CUDA(cudaMemcpyAsync(data(idx), &h_[idx], sizeof(T), cudaMemcpyHostToDevice, stream.cuda_stream()));
  }
  void fill(const T& t) { thrust::fill(d_.begin(), d_.end(), t); }
  void fill(const T& t, const Stream& stream) {
    thrust::fill(thrust::cuda::par.on(stream.cuda_stream()), d_.begin(), d_.end(), t);
  }
  typename device_t::reference get(size_t idx) { return d_[idx]; }
  typename device_t::const_reference get(size_t idx) const { return d_[idx]; }
  T get(size_t idx, const Stream& stream) const {
    T v;
    CHECK_CUDA(cudaMemcpyAsync(&v, data(idx), sizeof(T), cudaMemcpyDeviceToHost, stream.cuda_stream()));
    stream.Sync();
    return v;
  }
  const host_t& get(const Stream& stream) const {
    if (size()) {
      CHECK_CUDA(cudaMemcpyAsync(const_cast<T*>(thrust::raw_pointer_cast(h_.data())), data(), sizeof(T) * size(),
                                 cudaMemcpyDeviceToHost, stream.cuda_stream()));
      stream.Sync();
    }
    return h_;
  }
  host_t& get(const Stream& stream) { return const_cast<host_t&>(static_cast<const SharedArray*>(this)->get(stream)); }
  T* data() { return thrust::raw_pointer_cast(d_.data()); }
  const T* data() const { return thrust::raw_pointer_cast(d_.data()); }
  T* data(size_t idx) { return data() + idx; }
  const T* data(size_t idx) const { return data() + idx; }
  void Assign(const SharedArray<T>& rhs) {
    d_ = rhs.d_;
    h_.resize(d_.size());
  }
  void Assign(const SharedArray<T>& rhs, const Stream& stream) {
    resize(rhs.size());
    if (size()) CHECK_CUDA(cudaMemcpyAsync(data(), rhs.data(), sizeof(T) * size(), cudaMemcpyDeviceToDevice, stream.cuda_stream()));
  }
  void Swap(SharedArray<T>& rhs) {
    d_.swap(rhs.d_);
    h_.swap(rhs.h_);
  }

 private:
  device_t d_;
  mutable host_t h_;
};

// markers.h:25-190: NVTX ranges of the reference's PROFILING build.  The
// profiling recipe here is ncu launch lists (profiles/), so a marker only keeps
// its bookkeeping and nothing is emitted.
class RangeMarker {
 public:
  explicit RangeMarker(bool start = false, const char* = nullptr, int = 0, int = 0) : running_(start) {}
  explicit RangeMarker(const char*) {}
  void Start() { running_ = true; }
  void Stop() { running_ = false; }
  bool running() const { return running_; }
  static void MarkWorkitems(uint64_t, const char*) {}

 private:
  bool running_ = false;
};

// ------------------------------------------------------------- work source --
// grape/cuda/utils/work_source.h:22-48
template <typename T>
struct WorkSourceRange {
  DEV_HOST WorkSourceRange(T start, size_t size) : start_(start), size_(size) {}
  DEV_HOST_INLINE T GetWork(size_t i) const { return (T) (start_ + i); }
  DEV_HOST_INLINE size_t size() const { return size_; }
  T start_;
  size_t size_;
};
template <typename T>
struct WorkSourceArray {
  DEV_HOST WorkSourceArray(T* data, size_t size) : data_(data), size_(size) {}
  DEV_HOST_INLINE T GetWork(size_t i) const { return data_[i]; }
  DEV_HOST_INLINE size_t size() const { return size_; }
  T* data_;
  size_t size_;
};

// ---------------------------------------------------------------- launcher --
// grape/cuda/utils/launcher.h:23-53
template <typename F, typename... Args>
__global__ void KernelWrapper(F f, Args... args) {
  f(args...);
}
template <typename F, typename... Args>
void LaunchKernel(const Stream& stream, F f, Args&&... args) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  KernelWrapper<<<sms * 8, 256, 0, stream.cuda_stream()>>>(f, std::forward<Args>(args)...);
  CHECK_CUDA(cudaGetLastError());
}
template <typename F, typename... Args>
void LaunchKernel(const Stream& stream, size_t size, F f, Args&&... args) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  size_t blocks = std::min<size_t>((size + 255) / 256, (size_t) sms * 8);
  KernelWrapper<<<(unsigned) std::max<size_t>(blocks, 1), 256, 0, stream.cuda_stream()>>>(
      f, std::forward<Args>(args)...);
  CHECK_CUDA(cudaGetLastError());
}
template <typename F, typename... Args>
void LaunchKernelFix(const Stream& stream, size_t, F f, Args&&... args) {
  KernelWrapper<<<256, 256, 0, stream.cuda_stream()>>>(f, std::forward<Args>(args)...);
  CHECK_CUDA(cudaGetLastError());
}

// sorted_search.h:28-61: indices[i] = lower bound of needles[i] in the sorted haystack
template <typename T>
void sorted_search(const Stream& stream, T* needles, int num_needles, T* haystack, int num_haystack, T* indices) {
  if (num_needles <= 0) return;
  LaunchKernel(stream, (size_t) num_needles, [=] __device__() {
    for (int i = TID_1D; i < num_needles; i += TOTAL_THREADS_1D) {
      const T key = needles[i];
      int lo = 0, hi = num_haystack;
      while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (haystack[mid] < key) lo = mid + 1; else hi = mid;
      }
      indices[i] = (T) lo;
    }
  });
}

// --------------------------------------------------------------- dev utils --
// grape/cuda/utils/dev_utils.h:51-130 (atomics used by the apps)
namespace dev {
DEV_INLINE float atomicMinFloat(float* addr, float value) {
  // non-negative and negative halves ordered through the integer views
  return value >= 0
             ? __int_as_float(atomicMin(reinterpret_cast<int*>(addr), __float_as_int(value)))
             : __uint_as_float(atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(value)));
}
DEV_INLINE size_t atomicAdd64(size_t* address, size_t val) {
  return (size_t) atomicAdd(reinterpret_cast<unsigned long long*>(address), (unsigned long long) val);
}
DEV_INLINE int64_t atomicMin64(int64_t* address, int64_t val) {
  return (int64_t) atomicMin(reinterpret_cast<long long*>(address), (long long) val);
}
DEV_INLINE int64_t atomicCAS64(int64_t* address, int64_t compare, int64_t val) {
  return (int64_t) atomicCAS(reinterpret_cast<unsigned long long*>(address), (unsigned long long) compare,
                             (unsigned long long) val);
}
DEV_INLINE uint64_t atomicCAS64(uint64_t* address, uint64_t compare, uint64_t val) {
  return (uint64_t) atomicCAS(reinterpret_cast<unsigned long long*>(address), (unsigned long long) compare,
                              (unsigned long long) val);
}

// max / min over the lanes named by `mask`; every named lane gets the result
// (dev_utils.h:110-138).  One shuffle per member: correct for any lane subset.
template <typename T>
DEV_INLINE T reduce_max_sync(uint32_t mask, T val) {
  T best = val;
  for (uint32_t m = mask; m; m &= m - 1) {
    const T o = __shfl_sync(mask, val, __ffs(m) - 1);
    best = o > best ? o : best;
  }
  return best;
}
template <typename T>
DEV_INLINE T reduce_min_sync(uint32_t mask, T val) {
  T best = val;
  for (uint32_t m = mask; m; m &= m - 1) {
    const T o = __shfl_sync(mask, val, __ffs(m) - 1);
    best = o < best ? o : best;
  }
  return best;
}

// CTA-wide maximum delivered to every thread (dev_utils.h:281-293); the CTA
// must call it converged, blockDim.x a multiple of 32.
template <typename T>
DEV_INLINE T blockAllReduceMax(T val) {
  __shared__ T s_part[32];
  __shared__ T s_all;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const T t = __shfl_xor_sync(0xffffffffu, val, o);
    val = t > val ? t : val;
  }
  const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();   // a previous call's readers are done with s_part / s_all
  if (l == 0) s_part[w] = val;
  __syncthreads();
  if (w == 0) {
    T v = s_part[l < nw ? l : 0];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const T t = __shfl_xor_sync(0xffffffffu, v, o);
      v = t > v ? t : v;
    }
    if (l == 0) s_all = v;
  }
  __syncthreads();
  return s_all;
}

// warp / CTA sums delivered to every lane / thread (dev_utils.h:219-247)
template <typename T>
DEV_INLINE T warp_reduce(T val) {
#pragma unroll
  for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
  return val;
}
template <typename T>
DEV_INLINE T block_reduce(T val) {
  __shared__ T s_part[32];
  __shared__ T s_all;
  val = warp_reduce(val);
  const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) s_part[w] = val;
  __syncthreads();
  if (w == 0) {
    T v = l < nw ? s_part[l] : T(0);
    v = warp_reduce(v);
    if (l == 0) s_all = v;
  }
  __syncthreads();
  return s_all;
}

// Bucketed set in the CTA's 32 KB scratch (dev_utils.h:491-556), as the LCC app
// lays it out (lcc_opt.h:190-204): the first `bucket_stride` words are the fill
// counts of all buckets of the CTA, row r of the table follows at
// (1 + r) * bucket_stride; a cooperative group owns buckets [offset, offset +
// bucket_num).  Entries beyond `cached_size` per bucket go to the CTA's slice of
// a global overflow area.
template <typename T>
class ShmHashTable {
 public:
  __device__ __forceinline__ void init(T* shm_data, T* global_data, int offset, int bucket_size, int cached_size,
                                       int bucket_num, int bucket_stride) {
    fill_ = shm_data;
    rows_ = shm_data + bucket_stride;
    spill_ = global_data + (size_t) blockIdx.x * bucket_stride * (size_t) (bucket_size - cached_size);
    base_ = offset;
    cap_ = bucket_size;
    cached_ = cached_size;
    nb_ = bucket_num;
    stride_ = bucket_stride;
  }
  __device__ __forceinline__ void clear(int thread_lane, int csize) {
    for (int b = thread_lane; b < nb_; b += csize) fill_[base_ + b] = 0;
  }
  __device__ __forceinline__ bool insert(T x) {
    const int b = base_ + (int) (x & (T) (nb_ - 1));
    const int at = (int) atomicAdd(fill_ + b, (T) 1);
    if (at < cached_) rows_[(size_t) at * stride_ + b] = x;
    else if (at < cap_) spill_[(size_t) (at - cached_) * stride_ + b] = x;
    else return false;
    return true;
  }
  __device__ __forceinline__ bool lookup(T x) const {
    const int b = base_ + (int) (x & (T) (nb_ - 1));
    int n = (int) fill_[b];
    if (n > cap_) n = cap_;
    const int in_shm = n < cached_ ? n : cached_;
    for (int r = 0; r < in_shm; ++r)
      if (rows_[(size_t) r * stride_ + b] == x) return true;
    for (int r = in_shm; r < n; ++r)
      if (spill_[(size_t) (r - cached_) * stride_ + b] == x) return true;
    return false;
  }

 private:
  T* fill_;
  T* rows_;
  T* spill_;
  int base_, cap_, cached_, nb_, stride_;
};

// |A ∩ B| of two ascending arrays, counted cooperatively by a warp / a CTA: the
// shorter array is probed into the longer one by binary search; `callback(key)`
// runs once per match (dev_utils.h:558-706).  Every lane / thread gets the total.
template <typename T>
DEV_INLINE bool sorted_contains(const T* a, size_t n, T key) {
  size_t lo = 0, hi = n;
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo < n && a[lo] == key;
}
template <typename T, typename Y>
DEV_INLINE size_t intersect_num(T* a, size_t size_a, T* b, size_t size_b, Y callback) {
  const T* probe = a;
  const T* hay = b;
  size_t np = size_a, nh = size_b;
  if (size_a > size_b) {
    probe = b;
    hay = a;
    np = size_b;
    nh = size_a;
  }
  size_t mine = 0;
  if (nh)
    for (size_t i = threadIdx.x & 31; i < np; i += 32) {
      const T key = probe[i];
      if (sorted_contains(hay, nh, key)) {
        ++mine;
        callback(key);
      }
    }
  return warp_reduce(mine);
}
template <typename T, typename Y>
DEV_INLINE size_t intersect_num_blk(T* a, size_t size_a, T* b, size_t size_b, Y callback) {
  const T* probe = a;
  const T* hay = b;
  size_t np = size_a, nh = size_b;
  if (size_a > size_b) {
    probe = b;
    hay = a;
    np = size_b;
    nh = size_a;
  }
  size_t mine = 0;
  if (nh)
    for (size_t i = threadIdx.x; i < np; i += blockDim.x) {
      const T key = probe[i];
      if (sorted_contains(hay, nh, key)) {
        ++mine;
        callback(key);
      }
    }
  return block_reduce(mine);
}

// Directed-LCC variants (dev_utils.h:696-892): a match a[i] == b[j] additionally
// adds the per-entry weights of the OTHER list: *a_cnt += wb[j], *b_cnt += wa[i]
// (the weights mark reciprocal edges, lcc_directed_opt.h:186-240).  On return
// every lane / thread holds the warp / CTA totals.
template <typename T>
DEV_INLINE long long sorted_find(const T* a, size_t n, T key) {
  size_t lo = 0, hi = n;
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return (lo < n && a[lo] == key) ? (long long) lo : -1ll;
}
template <typename T, typename Y>
DEV_INLINE void intersect_weighted(T* a, size_t size_a, T* b, size_t size_b, char* wa, size_t* a_cnt, char* wb,
                                   size_t* b_cnt, Y callback, size_t first, size_t step) {
  if (size_a == 0 || size_b == 0) return;
  const bool swap = size_a > size_b;   // probe the shorter list into the longer one
  const T* probe = swap ? b : a;
  const T* hay = swap ? a : b;
  const size_t np = swap ? size_b : size_a, nh = swap ? size_a : size_b;
  for (size_t i = first; i < np; i += step) {
    const T key = probe[i];
    const long long j = sorted_find(hay, nh, key);
    if (j >= 0) {
      callback(key);
      const size_t ia = swap ? (size_t) j : i, ib = swap ? i : (size_t) j;
      *a_cnt += wb[ib];
      *b_cnt += wa[ia];
    }
  }
}
template <typename T, typename Y>
DEV_INLINE void intersect_num_d(T* a, size_t size_a, T* b, size_t size_b, char* wa, size_t* a_cnt, char* wb,
                                size_t* b_cnt, Y callback) {
  intersect_weighted(a, size_a, b, size_b, wa, a_cnt, wb, b_cnt, callback, threadIdx.x & 31, 32);
  __syncwarp();
  const size_t ra = warp_reduce(*a_cnt), rb = warp_reduce(*b_cnt);
  *a_cnt = ra;
  *b_cnt = rb;
  __syncwarp();
}
template <typename T, typename Y>
DEV_INLINE void intersect_num_blk_d(T* a, size_t size_a, T* b, size_t size_b, char* wa, size_t* a_cnt, char* wb,
                                    size_t* b_cnt, Y callback) {
  intersect_weighted(a, size_a, b, size_b, wa, a_cnt, wb, b_cnt, callback, threadIdx.x, blockDim.x);
  __syncthreads();
  const size_t ra = block_reduce(*a_cnt);
  const size_t rb = block_reduce(*b_cnt);
  __syncthreads();
  *a_cnt = ra;
  *b_cnt = rb;
}
template <typename T>
DEV_INLINE size_t intersect_num_directed(T* a, size_t size_a, T* b, size_t size_b, char* wb) {
  size_t mine = 0;
  if (size_a && size_b) {
    const bool swap = size_a > size_b;
    const T* probe = swap ? b : a;
    const T* hay = swap ? a : b;
    const size_t np = swap ? size_b : size_a, nh = swap ? size_a : size_b;
    for (size_t i = threadIdx.x & 31; i < np; i += 32) {
      const long long j = sorted_find(hay, nh, probe[i]);
      if (j >= 0) mine += wb[swap ? i : (size_t) j];
    }
  }
  const size_t total = warp_reduce(mine);
  __syncwarp();
  return total;
}

// Most-frequent-label counter of the CDLP app (dev_utils.h:384-470): an exact
// open-addressing table in the CTA's 32 KB scratch (one probe; a busy slot sends
// the label to a count-min sketch that only gives an upper bound), and an exact
// linear-probing table in global memory for the vertices where that is not
// enough.  The scratch layout is part of the interface — the app clears it
// itself (cdlp.h:378-385): [ht_size keys of T][ht_size u32 counts][cms_k x
// cms_size u32 sketch counters], per cooperative group `cid`.
template <typename T>
class MFLCounter {
 public:
  __device__ __forceinline__ void init(uint32_t* shm_data, T* global_data, uint32_t* global_cnt, int ht_size,
                                       int cms_size, int cms_k, int cid, int csize, int width, T dft) {
    uint32_t* mine = shm_data + (size_t) cid * ((size_t) (1 + width) * ht_size + (size_t) cms_size * cms_k);
    keys_ = reinterpret_cast<T*>(mine);
    cnts_ = mine + (size_t) ht_size * width;
    sketch_ = cnts_ + ht_size;
    ht_size_ = ht_size;
    cms_size_ = cms_size;
    cms_k_ = cms_k;
    gkeys_ = global_data;
    gcnts_ = global_cnt;
    empty_ = dft;
    (void) csize;
  }
  // count of `l` after this insertion, or -1 when its slot holds another label
  __device__ __forceinline__ int insert_shm_ht(T l) {
    const size_t at = slot(l, (size_t) ht_size_);
    const T seen = atomicCAS64(keys_ + at, empty_, l);
    if (seen != empty_ && seen != l) return -1;
    return (int) atomicAdd(cnts_ + at, 1u) + 1;
  }
  __device__ __forceinline__ int query_shm_ht(T l) {
    const size_t at = slot(l, (size_t) ht_size_);
    return keys_[at] == l ? (int) cnts_[at] : -1;
  }
  // upper bound of the count of `l` (minimum over cms_k sketch rows)
  __device__ __forceinline__ int insert_shm_cms(T l) {
    uint32_t bound = 0x7fffffffu;
    unsigned long long h = (unsigned long long) l;
    for (int r = 0; r < cms_k_; ++r) {
      h = (h ^ (h >> 31)) * 0x9E3779B97F4A7C15ull + (unsigned long long) r;
      const uint32_t c = atomicAdd(sketch_ + (size_t) r * cms_size_ + (size_t) (h % (unsigned long long) cms_size_), 1u) + 1;
      bound = c < bound ? c : bound;
    }
    return (int) bound;
  }
  // exact count of `l` in the vertex's private range [begin, end) of the global table
  __device__ __forceinline__ int insert_global_ht(T l, size_t begin, size_t end) {
    const size_t size = end - begin;
    size_t at = slot(l, size);
    for (;;) {
      const T seen = atomicCAS64(gkeys_ + begin + at, empty_, l);
      if (seen == empty_ || seen == l) break;
      at = at + 1 == size ? 0 : at + 1;
    }
    return (int) atomicAdd(gcnts_ + begin + at, 1u) + 1;
  }

 private:
  __device__ __forceinline__ static size_t slot(T l, size_t size) {
    return (size_t) ((unsigned long long) l % (unsigned long long) size);
  }
  T* keys_;
  uint32_t* cnts_;
  uint32_t* sketch_;
  int ht_size_, cms_size_, cms_k_;
  T* gkeys_;
  uint32_t* gcnts_;
  T empty_;
};
}  // namespace dev

namespace compat_detail {
template <typename I>
struct ToSizeT {
  __host__ __device__ size_t operator()(const I& x) const { return (size_t) x; }
};
}  // namespace compat_detail

// ---------------------------------------------------------- host utilities --
// grape/cuda/utils/cuda_utils.h:262-333: segmented key sort and prefix sums
// (temporary storage is allocated per call, as in the reference).
template <typename T>
T* SegmentSort(T* d_keys_in, T* d_keys_buffer, size_t* d_offset_lo, size_t* d_offset_hi, size_t num_items,
               size_t num_segments) {
  if (num_items == 0 || num_segments == 0) return d_keys_in;
  cub::DoubleBuffer<T> keys(d_keys_in, d_keys_buffer);
  size_t bytes = 0;
  CHECK_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, bytes, keys, (int64_t) num_items, (int64_t) num_segments,
                                                d_offset_lo, d_offset_hi));
  void* tmp = nullptr;
  CHECK_CUDA(cudaMalloc(&tmp, std::max<size_t>(bytes, 16)));
  CHECK_CUDA(cub::DeviceSegmentedSort::SortKeys(tmp, bytes, keys, (int64_t) num_items, (int64_t) num_segments,
                                                d_offset_lo, d_offset_hi));
  CHECK_CUDA(cudaDeviceSynchronize());
  CHECK_CUDA(cudaFree(tmp));
  return keys.Current();
}
template <typename I, typename O>
void InclusiveSum(I* d_in, O* d_out, size_t size, cudaStream_t stream) {
  size_t bytes = 0;
  CHECK_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, d_in, d_out, (int64_t) size, stream));
  void* tmp = nullptr;
  CHECK_CUDA(cudaMallocAsync(&tmp, std::max<size_t>(bytes, 16), stream));
  CHECK_CUDA(cub::DeviceScan::InclusiveSum(tmp, bytes, d_in, d_out, (int64_t) size, stream));
  CHECK_CUDA(cudaFreeAsync(tmp, stream));
}
template <typename I, typename O>
void ExclusiveSum64(I* d_in, O* d_out, size_t size, cudaStream_t stream) {   // accumulates in size_t
  size_t bytes = 0;
  cub::TransformInputIterator<size_t, compat_detail::ToSizeT<I>, I*> in(d_in, compat_detail::ToSizeT<I>());
  CHECK_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, d_out, (int64_t) size, stream));
  void* tmp = nullptr;
  CHECK_CUDA(cudaMallocAsync(&tmp, std::max<size_t>(bytes, 16), stream));
  CHECK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, d_out, (int64_t) size, stream));
  CHECK_CUDA(cudaFreeAsync(tmp, stream));
}

// ------------------------------------------------------------ shared value --
// grape/cuda/utils/shared_value.h:28-80: device scalar with a host mirror
template <typename T>
class SharedValue {
 public:
  SharedValue() {
    CHECK_CUDA(cudaMalloc(&d_, sizeof(T)));
    CHECK_CUDA(cudaMallocHost(&h_, sizeof(T)));
    CHECK_CUDA(cudaMemset(d_, 0, sizeof(T)));
  }
  SharedValue(const SharedValue&) = delete;
  ~SharedValue() {
    if (d_) cudaFree(d_);
    if (h_) cudaFreeHost(h_);
  }
  void set(const T& t, const Stream& s) {
    *h_ = t;
    CHECK_CUDA(cudaMemcpyAsync(d_, h_, sizeof(T), cudaMemcpyHostToDevice, s.cuda_stream()));
  }
  T get(const Stream& s) const {
    CHECK_CUDA(cudaMemcpyAsync(h_, d_, sizeof(T), cudaMemcpyDeviceToHost, s.cuda_stream()));
    s.Sync();
    return *h_;
  }
  T* data() { return d_; }
  const T* data() const { return d_; }
  void Swap(SharedValue& o) {
    std::swap(d_, o.d_);
    std::swap(h_, o.h_);
  }

 private:
  T* d_ = nullptr;
  T* h_ = nullptr;
};

// ------------------------------------------------------------ vertex array --
// grape/cuda/utils/vertex_array.h:34-169
namespace dev {
template <typename T, typename VID_T>
class VertexArray {
 public:
  VertexArray() = default;
  DEV_HOST VertexArray(T* data, VID_T begin, size_t size) : fake_(data - begin), size_(size) {}
  DEV_INLINE T& operator[](const Vertex<VID_T>& v) { return fake_[v.GetValue()]; }
  DEV_INLINE const T& operator[](const Vertex<VID_T>& v) const { return fake_[v.GetValue()]; }
  DEV_HOST_INLINE T* data() { return fake_; }
  DEV_HOST_INLINE size_t size() const { return size_; }

 private:
  T* fake_ = nullptr;   // base pointer shifted by the range start
  size_t size_ = 0;
};
}  // namespace dev

template <typename T, typename VID_T>
class VertexArray {
 public:
  VertexArray() = default;
  VertexArray(const VertexArray&) = delete;
  ~VertexArray() {
    if (d_) cudaFree(d_);
  }
  void Init(const VertexRange<VID_T>& range) {
    range_ = range;
    h_.assign(range.size(), T());
    if (d_) cudaFree(d_);
    CHECK_CUDA(cudaMalloc(&d_, sizeof(T) * std::max<size_t>(range.size(), 1)));
  }
  void Init(const VertexRange<VID_T>& range, const T& value) {
    Init(range);
    std::fill(h_.begin(), h_.end(), value);
  }
  void SetValue(const T& value) { std::fill(h_.begin(), h_.end(), value); }
  T& operator[](const Vertex<VID_T>& v) { return h_[v.GetValue() - range_.begin_value()]; }
  const T& operator[](const Vertex<VID_T>& v) const { return h_[v.GetValue() - range_.begin_value()]; }
  void H2D() {
    if (!h_.empty()) CHECK_CUDA(cudaMemcpy(d_, h_.data(), sizeof(T) * h_.size(), cudaMemcpyHostToDevice));
  }
  void H2D(const Stream& s) {
    if (!h_.empty())
      CHECK_CUDA(cudaMemcpyAsync(d_, h_.data(), sizeof(T) * h_.size(), cudaMemcpyHostToDevice, s.cuda_stream()));
  }
  void D2H() {
    if (!h_.empty()) CHECK_CUDA(cudaMemcpy(h_.data(), d_, sizeof(T) * h_.size(), cudaMemcpyDeviceToHost));
  }
  void D2H(const Stream& s) {
    if (!h_.empty())
      CHECK_CUDA(cudaMemcpyAsync(h_.data(), d_, sizeof(T) * h_.size(), cudaMemcpyDeviceToHost, s.cuda_stream()));
    s.Sync();
  }
  void Swap(VertexArray& o) {
    std::swap(range_, o.range_);
    h_.swap(o.h_);
    std::swap(d_, o.d_);
  }
  dev::VertexArray<T, VID_T> DeviceObject() {
    return dev::VertexArray<T, VID_T>(d_, range_.begin_value(), range_.size());
  }
  size_t size() const { return h_.size(); }

 private:
  VertexRange<VID_T> range_;
  std::vector<T> h_;
  T* d_ = nullptr;
};

// -------------------------------------------------------------- vertex set --
// grape/cuda/utils/bitset.h:32-103 + vertex_set.h:25-98: frontier bitmap with
// a device-side population counter
namespace dev {
template <typename VID_T>
class DenseVertexSet {
 public:
  DenseVertexSet() = default;
  DEV_HOST DenseVertexSet(VID_T beg, uint32_t* bits, unsigned long long* count)
      : beg_(beg), bits_(bits), count_(count) {}
  DEV_INLINE bool Insert(Vertex<VID_T> v) {
    const VID_T i = v.GetValue() - beg_;
    const uint32_t m = 1u << (i & 31);
    const uint32_t old = atomicOr(bits_ + (i >> 5), m);
    if (!(old & m)) {
      atomicAdd(count_, 1ull);
      return true;
    }
    return false;
  }
  DEV_INLINE bool Exist(Vertex<VID_T> v) const {
    const VID_T i = v.GetValue() - beg_;
    return (bits_[i >> 5] >> (i & 31)) & 1u;
  }
  DEV_INLINE void Clear() {}
  DEV_INLINE size_t Count() const { return (size_t) *count_; }

 private:
  VID_T beg_ = 0;
  uint32_t* bits_ = nullptr;
  unsigned long long* count_ = nullptr;
};
}  // namespace dev

template <typename VID_T>
class DenseVertexSet {
 public:
  DenseVertexSet() = default;
  DenseVertexSet(const DenseVertexSet&) = delete;
  ~DenseVertexSet() {
    if (bits_) cudaFree(bits_);
    if (count_) cudaFree(count_);
    if (h_count_) cudaFreeHost(h_count_);
  }
  void Init(const VertexRange<VID_T>& range) {
    beg_ = range.begin_value();
    words_ = (range.size() + 31) / 32 + 1;
    if (bits_) cudaFree(bits_);
    CHECK_CUDA(cudaMalloc(&bits_, sizeof(uint32_t) * words_));
    CHECK_CUDA(cudaMemset(bits_, 0, sizeof(uint32_t) * words_));
    if (!count_) {
      CHECK_CUDA(cudaMalloc(&count_, sizeof(unsigned long long)));
      CHECK_CUDA(cudaMallocHost(&h_count_, sizeof(unsigned long long)));
    }
    CHECK_CUDA(cudaMemset(count_, 0, sizeof(unsigned long long)));
  }
  dev::DenseVertexSet<VID_T> DeviceObject() { return dev::DenseVertexSet<VID_T>(beg_, bits_, count_); }
  void Clear(const Stream& s) {
    CHECK_CUDA(cudaMemsetAsync(bits_, 0, sizeof(uint32_t) * words_, s.cuda_stream()));
    CHECK_CUDA(cudaMemsetAsync(count_, 0, sizeof(unsigned long long), s.cuda_stream()));
  }
  size_t Count(const Stream& s) const {
    CHECK_CUDA(cudaMemcpyAsync(h_count_, count_, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.cuda_stream()));
    s.Sync();
    return (size_t) *h_count_;
  }
  void Swap(DenseVertexSet& o) {
    std::swap(beg_, o.beg_);
    std::swap(words_, o.words_);
    std::swap(bits_, o.bits_);
    std::swap(count_, o.count_);
    std::swap(h_count_, o.h_count_);
  }

 private:
  VID_T beg_ = 0;
  size_t words_ = 0;
  uint32_t* bits_ = nullptr;
  unsigned long long* count_ = nullptr;
  unsigned long long* h_count_ = nullptr;
};

// ------------------------------------------------------------------- queue --
// grape/cuda/utils/queue.h:47-178
namespace dev {
template <typename T, typename SIZE_T>
class Queue {
 public:
  Queue() = default;
  DEV_HOST Queue(T* data, SIZE_T* last) : data_(data), last_(last) {}
  DEV_INLINE void Append(const T& item) {
    SIZE_T at = atomicAdd(last_, (SIZE_T) 1);
    data_[at] = item;
  }
  DEV_INLINE void AppendWarp(const T& item) {
    const unsigned mask = __activemask();
    const int leader = __ffs(mask) - 1;
    const int lane = threadIdx.x & 31;
    SIZE_T base = 0;
    if (lane == leader) base = atomicAdd(last_, (SIZE_T) __popc(mask));
    base = __shfl_sync(mask, base, leader);
    data_[base + __popc(mask & ((1u << lane) - 1))] = item;
  }
  DEV_INLINE void Clear() const { *last_ = 0; }
  DEV_INLINE T& operator[](SIZE_T i) { return data_[i]; }
  DEV_INLINE SIZE_T size() const { return *last_; }

 private:
  T* data_ = nullptr;
  SIZE_T* last_ = nullptr;
};
}  // namespace dev

template <typename T, typename SIZE_T = uint32_t>
class Queue {
 public:
  using device_t = dev::Queue<T, SIZE_T>;
  Queue() = default;
  Queue(const Queue&) = delete;
  ~Queue() {
    if (data_) cudaFree(data_);
  }
  void Init(SIZE_T capacity) {
    if (data_) cudaFree(data_);
    CHECK_CUDA(cudaMalloc(&data_, sizeof(T) * std::max<size_t>(capacity, 1)));
    CHECK_CUDA(cudaMemset(counter_.data(), 0, sizeof(SIZE_T)));
  }
  void Clear(const Stream& s) {
    CHECK_CUDA(cudaMemsetAsync(counter_.data(), 0, sizeof(SIZE_T), s.cuda_stream()));
  }
  size_t size(const Stream& s) const { return counter_.get(s); }
  T* data() { return data_; }
  const T* data() const { return data_; }
  device_t DeviceObject() { return device_t(data_, counter_.data()); }
  void Swap(Queue& o) {
    std::swap(data_, o.data_);
    counter_.Swap(o.counter_);
  }

 private:
  T* data_ = nullptr;
  SharedValue<SIZE_T> counter_;
};

// ------------------------------------------------------------ coo fragment --
// grape/cuda/fragment/coo_fragment.h:29-107: the inner vertices' out-edges as a
// flat (src, dst[, data]) list on the device (what wcc_opt.h iterates).
namespace dev {
template <typename OID_T, typename VID_T, typename VDATA_T, typename EDATA_T>
class COOFragment {
 public:
  using oid_t = OID_T;
  using vid_t = VID_T;
  using vdata_t = VDATA_T;
  using edata_t = EDATA_T;
  using vertex_t = Vertex<vid_t>;
  using edge_t = Edge<vid_t, EDATA_T>;
  COOFragment() = default;
  DEV_HOST COOFragment(ArrayView<edge_t> edges) : edges_(edges) {}
  DEV_INLINE const edge_t& edge(size_t eid) const { return edges_[eid]; }
  DEV_INLINE edge_t& edge(size_t eid) { return edges_[eid]; }
  DEV_INLINE edge_t& operator[](size_t eid) const { return const_cast<ArrayView<edge_t>&>(edges_)[eid]; }
  DEV_HOST_INLINE size_t GetEdgeNum() const { return edges_.size(); }

 private:
  ArrayView<edge_t> edges_;
};
}  // namespace dev
template <typename OID_T, typename VID_T, typename VDATA_T, typename EDATA_T>
class COOFragment {
 public:
  using oid_t = OID_T;
  using vid_t = VID_T;
  using vdata_t = VDATA_T;
  using edata_t = EDATA_T;
  using vertex_t = Vertex<VID_T>;
  using edge_t = Edge<VID_T, EDATA_T>;
  using device_t = dev::COOFragment<OID_T, VID_T, VDATA_T, EDATA_T>;
  void Init(const thrust::host_vector<edge_t>& edges) { edges_ = edges; }
  device_t DeviceObject() { return device_t(ArrayView<edge_t>(edges_)); }
  size_t GetEdgeNum() const { return edges_.size(); }

 private:
  thrust::device_vector<edge_t> edges_;
};

// ------------------------------------------------------- device vertex map --
// grape/cuda/vertex_map/device_vertex_map.h:33-172 on gl_vm_* (one oid array per
// group + binary search; no hash chains).
namespace dev {
template <typename OID_T, typename VID_T>
class DeviceVertexMap {
 public:
  DeviceVertexMap() { memset(&v_, 0, sizeof(v_)); }
  explicit DeviceVertexMap(const gl_vm_view& v) : v_(v) {}
  DEV_INLINE bool GetOid(const VID_T& gid, OID_T& oid) const {
    return GetOid((fid_t) (gid >> v_.fid_offset), (VID_T) (gid & v_.id_mask), oid);
  }
  DEV_INLINE bool GetOid(fid_t fid, const VID_T& lid, OID_T& oid) const {
    if (!v_.l2o || fid >= v_.fnum || lid >= v_.off[fid + 1] - v_.off[fid]) return false;
    oid = (OID_T) v_.l2o[v_.off[fid] + lid];
    return true;
  }
  DEV_INLINE bool GetGid(fid_t fid, const OID_T& oid, VID_T& gid) const {
    if (!v_.l2o || fid >= v_.fnum) return false;
    const int64_t* keys = (v_.sorted_oid ? v_.sorted_oid : v_.l2o) + v_.off[fid];
    uint64_t lo = 0, hi = v_.off[fid + 1] - v_.off[fid];
    const uint64_t n = hi;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (keys[mid] < (int64_t) oid) lo = mid + 1; else hi = mid;
    }
    if (lo >= n || keys[lo] != (int64_t) oid) return false;
    const VID_T lid = v_.sorted_lid ? (VID_T) v_.sorted_lid[v_.off[fid] + lo] : (VID_T) lo;
    gid = ((VID_T) fid << v_.fid_offset) | lid;
    return true;
  }
  DEV_INLINE bool GetGid(const OID_T& oid, VID_T& gid) const {
    for (fid_t f = 0; f < v_.fnum; ++f)
      if (GetGid(f, oid, gid)) return true;
    return false;
  }
  DEV_HOST_INLINE bool built() const { return v_.l2o != nullptr; }

 private:
  gl_vm_view v_;
};
}  // namespace dev
template <typename HOST_VM_T>
class DeviceVertexMap {
  using OID_T = typename HOST_VM_T::oid_t;
  using VID_T = typename HOST_VM_T::vid_t;

 public:
  DeviceVertexMap() = default;
  ~DeviceVertexMap() {
    if (vm_) gl_vm_destroy(vm_);
  }
  DeviceVertexMap(const DeviceVertexMap&) = delete;
  DeviceVertexMap& operator=(const DeviceVertexMap&) = delete;
  // :103-146: the host VertexMap's (fid, lid) -> oid tables go to the device
  void Init(const Stream&, const CommSpec& comm_spec, std::unique_ptr<HOST_VM_T>& vm_ptr) {
    const fid_t fnum = comm_spec.fnum();
    std::vector<std::vector<int64_t>> oids(fnum);
    std::vector<uint64_t> ivnums(fnum);
    std::vector<const int64_t*> ptrs(fnum);
    for (fid_t f = 0; f < fnum; ++f) {
      const size_t n = vm_ptr->GetInnerVertexSize(f);
      oids[f].resize(n);
      for (size_t lid = 0; lid < n; ++lid) {
        OID_T oid;
        CHECK(vm_ptr->GetOid(f, (VID_T) lid, oid));
        oids[f][lid] = (int64_t) oid;
      }
      ivnums[f] = n;
      ptrs[f] = oids[f].data();
    }
    if (vm_) gl_vm_destroy(vm_);
    CHECK_GL(gl_vm_create(&vm_, fnum, ivnums.data(), ptrs.data()));
  }
  dev::DeviceVertexMap<OID_T, VID_T> DeviceObject() const {
    if (!vm_) return dev::DeviceVertexMap<OID_T, VID_T>();
    gl_vm_view v;
    CHECK_GL(gl_vm_view_get(vm_, &v));
    return dev::DeviceVertexMap<OID_T, VID_T>(v);
  }
  gl_vm_t* handle() const { return vm_; }

 private:
  gl_vm_t* vm_ = nullptr;
};

// --------------------------------------------------------- device fragment --
// grape/cuda/fragment/device_fragment.h:36-450 — accessors over the SoA view
namespace dev {

// grape::Nbr plus the CSR position of the entry it was loaded from.  The
// reference computes edge indices as `&nbr - row begin` in its AoS Nbr array
// (device_fragment.h:398-421); with SoA storage the neighbour handed to an app
// is a value, so it carries its position instead.
template <typename VID_T, typename EDATA_T>
struct PosNbr : public Nbr<VID_T, EDATA_T> {
  using base_t = Nbr<VID_T, EDATA_T>;
  DEV_HOST PosNbr() : base_t(), pos_(0) {}
  DEV_HOST explicit PosNbr(VID_T v, uint64_t pos = 0) : base_t(v), pos_(pos) {}
  DEV_HOST PosNbr(VID_T v, const EDATA_T& d, uint64_t pos = 0) : base_t(v, d), pos_(pos) {}
  uint64_t pos_;
};

template <typename VID_T, typename EDATA_T>
class AdjList {
 public:
  using nbr_t = PosNbr<VID_T, EDATA_T>;
  class iterator {
   public:
    DEV_HOST iterator(const uint32_t* col, const EDATA_T* w, uint64_t pos) : col_(col), w_(w), pos_(pos) { load(); }
    DEV_HOST_INLINE nbr_t& operator*() { return cur_; }
    DEV_HOST_INLINE nbr_t* operator->() { return &cur_; }
    DEV_HOST_INLINE iterator& operator++() {
      ++pos_;
      load();
      return *this;
    }
    DEV_HOST_INLINE bool operator!=(const iterator& o) const { return pos_ != o.pos_; }
    DEV_HOST_INLINE bool operator==(const iterator& o) const { return pos_ == o.pos_; }

   private:
    DEV_HOST_INLINE void load() { load_impl(std::is_same<EDATA_T, EmptyType>()); }
    DEV_HOST_INLINE void load_impl(std::true_type) {
      if (col_) cur_ = nbr_t(col_[pos_], pos_);
    }
    DEV_HOST_INLINE void load_impl(std::false_type) {
      if (col_) cur_ = nbr_t(col_[pos_], w_ ? w_[pos_] : EDATA_T(), pos_);
    }
    const uint32_t* col_;
    const EDATA_T* w_;
    uint64_t pos_;
    nbr_t cur_;
  };
  DEV_HOST AdjList() : col_(nullptr), w_(nullptr), b_(0), e_(0) {}
  DEV_HOST AdjList(const uint32_t* col, const EDATA_T* w, uint64_t b, uint64_t e) : col_(col), w_(w), b_(b), e_(e) {}
  // `end()` never dereferences (load() is skipped by passing a null column)
  DEV_HOST_INLINE iterator begin() const { return b_ < e_ ? iterator(col_, w_, b_) : iterator(nullptr, nullptr, e_); }
  DEV_HOST_INLINE iterator end() const { return iterator(nullptr, nullptr, e_); }
  DEV_HOST_INLINE size_t Size() const { return (size_t) (e_ - b_); }
  DEV_HOST_INLINE bool Empty() const { return b_ == e_; }
  DEV_HOST_INLINE bool NotEmpty() const { return b_ != e_; }

 private:
  const uint32_t* col_;
  const EDATA_T* w_;
  uint64_t b_, e_;
};

template <typename OID_T, typename VID_T, typename VDATA_T, typename EDATA_T,
          grape::LoadStrategy _load_strategy = grape::LoadStrategy::kOnlyOut>
class DeviceFragment {
 public:
  using vertex_t = Vertex<VID_T>;
  using nbr_t = PosNbr<VID_T, EDATA_T>;
  using vertex_range_t = VertexRange<VID_T>;
  using adj_list_t = AdjList<VID_T, EDATA_T>;
  using const_adj_list_t = AdjList<VID_T, EDATA_T>;
  using vid_t = VID_T;
  using oid_t = OID_T;
  using vdata_t = VDATA_T;
  using edata_t = EDATA_T;
  static constexpr grape::LoadStrategy load_strategy = _load_strategy;

  DeviceFragment() = default;
  explicit DeviceFragment(const gl_frag_view& v) : v_(v) {}
  DeviceFragment(const gl_frag_view& v, const DeviceVertexMap<OID_T, VID_T>& vm) : v_(v), vm_(vm) {}

  DEV_HOST_INLINE const gl_frag_view& view() const { return v_; }
  DEV_HOST_INLINE vertex_range_t Vertices() const { return vertex_range_t(0, v_.ivnum + v_.ovnum); }
  DEV_HOST_INLINE vertex_range_t InnerVertices() const { return vertex_range_t(0, v_.ivnum); }
  DEV_HOST_INLINE vertex_range_t OuterVertices() const { return vertex_range_t(v_.ivnum, v_.ivnum + v_.ovnum); }
  DEV_INLINE vertex_range_t OuterVertices(fid_t fid) const {
    return vertex_range_t(v_.outer_range[fid], v_.outer_range[fid + 1]);
  }
  DEV_HOST_INLINE fid_t fid() const { return v_.fid; }
  DEV_HOST_INLINE fid_t fnum() const { return v_.fnum; }
  DEV_HOST_INLINE VID_T GetInnerVerticesNum() const { return v_.ivnum; }
  DEV_HOST_INLINE VID_T GetOuterVerticesNum() const { return v_.ovnum; }
  DEV_HOST_INLINE VID_T GetVerticesNum() const { return v_.ivnum + v_.ovnum; }
  DEV_HOST_INLINE size_t GetTotalVerticesNum() const { return v_.total_vnum; }
  DEV_HOST_INLINE bool IsInnerVertex(const vertex_t& v) const { return v.GetValue() < v_.ivnum; }
  DEV_HOST_INLINE bool IsOuterVertex(const vertex_t& v) const {
    return v.GetValue() >= v_.ivnum && v.GetValue() < v_.ivnum + v_.ovnum;
  }
  DEV_INLINE VID_T GetInnerVertexGid(const vertex_t& v) const { return (v_.fid << v_.fid_offset) | v.GetValue(); }
  DEV_INLINE VID_T GetOuterVertexGid(const vertex_t& v) const { return v_.ovgid[v.GetValue() - v_.ivnum]; }
  DEV_INLINE VID_T Vertex2Gid(const vertex_t& v) const {
    return IsInnerVertex(v) ? GetInnerVertexGid(v) : GetOuterVertexGid(v);
  }
  DEV_INLINE fid_t GetFragId(const vertex_t& v) const {
    return IsInnerVertex(v) ? v_.fid : (fid_t) (GetOuterVertexGid(v) >> v_.fid_offset);
  }
  DEV_INLINE bool InnerVertexGid2Vertex(VID_T gid, vertex_t& v) const {
    v.SetValue(gid & v_.id_mask);
    return (gid >> v_.fid_offset) == v_.fid;
  }
  DEV_INLINE bool OuterVertexGid2Vertex(VID_T gid, vertex_t& v) const {
    // outer gids are stored ascending: binary search (replaces the chained
    // hash map of thirdparty/cuda_hashmap, device_fragment.h:179-187)
    uint32_t lo = 0, hi = v_.ovnum;
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (v_.ovgid[mid] < gid) lo = mid + 1; else hi = mid;
    }
    if (lo < v_.ovnum && v_.ovgid[lo] == gid) {
      v.SetValue(v_.ivnum + lo);
      return true;
    }
    return false;
  }
  DEV_INLINE bool Gid2Vertex(VID_T gid, vertex_t& v) const {
    return (gid >> v_.fid_offset) == v_.fid ? InnerVertexGid2Vertex(gid, v) : OuterVertexGid2Vertex(gid, v);
  }
  DEV_INLINE int GetLocalOutDegree(const vertex_t& v) const {
    return v.GetValue() < v_.ivnum ? (int) (v_.oe_rp[v.GetValue() + 1] - v_.oe_rp[v.GetValue()]) : 0;
  }
  DEV_INLINE int GetLocalInDegree(const vertex_t& v) const {
    if (v.GetValue() < v_.ivnum) return (int) (v_.ie_rp[v.GetValue() + 1] - v_.ie_rp[v.GetValue()]);
    const VID_T o = v.GetValue() - v_.ivnum;
    return v_.ovie_rp ? (int) (v_.ovie_rp[o + 1] - v_.ovie_rp[o]) : 0;
  }
  // position of an entry inside the whole CSR / inside u's row
  // (device_fragment.h:398-421)
  DEV_INLINE size_t GetOutgoingEdgeIndex(const nbr_t& nbr) const { return (size_t) nbr.pos_; }
  DEV_INLINE size_t GetOutgoingEdgeIndex(const vertex_t& u, const nbr_t& nbr) const {
    return (size_t) (nbr.pos_ - v_.oe_rp[u.GetValue()]);
  }
  DEV_INLINE size_t GetIncomingEdgeIndex(const nbr_t& nbr) const { return (size_t) nbr.pos_; }
  DEV_INLINE size_t GetIncomingEdgeIndex(const vertex_t& u, const nbr_t& nbr) const {
    return (size_t) (nbr.pos_ - v_.ie_rp[u.GetValue()]);
  }
  DEV_INLINE adj_list_t GetOutgoingAdjList(const vertex_t& v) const {
    const VID_T u = v.GetValue();
    if (u >= v_.ivnum) return adj_list_t();
    return adj_list_t(v_.oe_col, (const EDATA_T*) v_.oe_w, v_.oe_rp[u], v_.oe_rp[u + 1]);
  }
  DEV_INLINE adj_list_t GetIncomingAdjList(const vertex_t& v) const {
    const VID_T u = v.GetValue();
    if (u < v_.ivnum) return adj_list_t(v_.ie_col, (const EDATA_T*) v_.ie_w, v_.ie_rp[u], v_.ie_rp[u + 1]);
    // outer vertex: its reverse adjacency (inner neighbours)
    const VID_T o = u - v_.ivnum;
    if (!v_.ovie_rp) return adj_list_t();
    return adj_list_t(v_.ovie_col, nullptr, v_.ovie_rp[o], v_.ovie_rp[o + 1]);
  }
  DEV_INLINE adj_list_t GetOutgoingInnerVertexAdjList(const vertex_t& v) const {
    const VID_T u = v.GetValue();
    return adj_list_t(v_.oe_col, (const EDATA_T*) v_.oe_w, v_.oe_rp[u], v_.oe_split[u]);
  }
  DEV_INLINE adj_list_t GetOutgoingOuterVertexAdjList(const vertex_t& v) const {
    const VID_T u = v.GetValue();
    return adj_list_t(v_.oe_col, (const EDATA_T*) v_.oe_w, v_.oe_split[u], v_.oe_rp[u + 1]);
  }
  DEV_INLINE adj_list_t GetIncomingInnerVertexAdjList(const vertex_t& v) const {
    const VID_T u = v.GetValue();
    return adj_list_t(v_.ie_col, (const EDATA_T*) v_.ie_w, v_.ie_rp[u], v_.ie_split[u]);
  }
  DEV_INLINE adj_list_t GetIncomingOuterVertexAdjList(const vertex_t& v) const {
    const VID_T u = v.GetValue();
    return adj_list_t(v_.ie_col, (const EDATA_T*) v_.ie_w, v_.ie_split[u], v_.ie_rp[u + 1]);
  }
  // device_fragment.h:64-172: oid <-> vertex; outer vertices and foreign gids need the
  // device vertex map (apps that set need_build_device_vm)
  DEV_INLINE OID_T GetId(const vertex_t& v) const {
    const VID_T u = v.GetValue();
    if (u < v_.ivnum) return v_.inner_oids ? (OID_T) v_.inner_oids[u] : (OID_T) (v_.oid_base + u);
    return Gid2Oid(GetOuterVertexGid(v));
  }
  DEV_INLINE OID_T Gid2Oid(const VID_T& gid) const {
    OID_T oid = OID_T();
    bool ok = vm_.GetOid(gid, oid);
    assert(ok);
    (void) ok;
    return oid;
  }
  DEV_INLINE bool Oid2Gid(const OID_T& oid, VID_T& gid) const { return vm_.GetGid(oid, gid); }
  DEV_INLINE bool GetVertex(const OID_T& oid, vertex_t& v) const {
    VID_T gid;
    return vm_.GetGid(oid, gid) && Gid2Vertex(gid, v);
  }
  DEV_INLINE bool GetInnerVertex(const OID_T& oid, vertex_t& v) const {
    VID_T gid;
    return vm_.GetGid((fid_t) v_.fid, oid, gid) && InnerVertexGid2Vertex(gid, v);
  }
  DEV_INLINE bool GetOuterVertex(const OID_T& oid, vertex_t& v) const {
    VID_T gid;
    return vm_.GetGid(oid, gid) && OuterVertexGid2Vertex(gid, v);
  }
  DEV_INLINE OID_T GetInnerVertexId(const vertex_t& v) const { return GetId(v); }
  DEV_INLINE OID_T GetOuterVertexId(const vertex_t& v) const { return Gid2Oid(GetOuterVertexGid(v)); }

 private:
  gl_frag_view v_;
  DeviceVertexMap<OID_T, VID_T> vm_;
};
}  // namespace dev

inline int b200_pick_device(int local_id) {
  // run_cuda_app.h:207-214 binds rank i to device i; with more ranks than
  // devices (several fragments time-slicing one GPU) ranks wrap around
  int n = 0;
  CHECK_CUDA(cudaGetDeviceCount(&n));
  return n > 0 ? local_id % n : 0;
}

// ------------------------------------------------------------ host fragment --
// grape/cuda/fragment/host_fragment.h:66-660: the CPU fragment of the
// reference + a device-resident SoA copy owned by the C-ABI library.
template <typename OID_T, typename VID_T, typename VDATA_T, typename EDATA_T,
          grape::LoadStrategy _load_strategy = grape::LoadStrategy::kOnlyOut>
class HostFragment
    : public ImmutableEdgecutFragment<OID_T, VID_T, VDATA_T, EDATA_T, _load_strategy> {
 public:
  using base_t = ImmutableEdgecutFragment<OID_T, VID_T, VDATA_T, EDATA_T, _load_strategy>;
  using internal_vertex_t = typename base_t::internal_vertex_t;
  using edge_t = typename base_t::edge_t;
  using nbr_t = typename base_t::nbr_t;
  using vertex_t = typename base_t::vertex_t;
  using const_adj_list_t = typename base_t::const_adj_list_t;
  using adj_list_t = typename base_t::adj_list_t;
  using traits_t = typename base_t::traits_t;
  using vid_t = VID_T;
  using oid_t = OID_T;
  using vdata_t = VDATA_T;
  using edata_t = EDATA_T;
  using vertex_range_t = typename base_t::vertex_range_t;
  using inner_vertices_t = typename base_t::inner_vertices_t;
  using outer_vertices_t = typename base_t::outer_vertices_t;
  using device_t = dev::DeviceFragment<OID_T, VID_T, VDATA_T, EDATA_T, _load_strategy>;
  using coo_t = COOFragment<OID_T, VID_T, VDATA_T, EDATA_T>;
  using dev_vertex_map_t = cuda::DeviceVertexMap<VertexMap<OID_T, VID_T>>;
  using IsEdgeCut = std::true_type;
  using IsVertexCut = std::false_type;
  static constexpr grape::LoadStrategy load_strategy = _load_strategy;

  HostFragment() : FragmentBase<OID_T, VDATA_T, EDATA_T>() {}
  ~HostFragment() {
    if (handle_) gl_frag_destroy(handle_);
  }

  void Init(const CommSpec& comm_spec, bool directed, std::unique_ptr<VertexMap<OID_T, VID_T>>&& vm_ptr,
            std::vector<internal_vertex_t>& vertices, std::vector<edge_t>& edges) {
    base_t::Init(comm_spec, directed, std::move(vm_ptr), vertices, edges);
    Upload(b200_pick_device(comm_spec.local_id()));
  }

  template <typename IOADAPTOR_T>
  void Deserialize(const CommSpec& comm_spec, std::unique_ptr<VertexMap<OID_T, VID_T>>&& vm_ptr,
                   const std::string& prefix) {
    base_t::template Deserialize<IOADAPTOR_T>(comm_spec, std::move(vm_ptr), prefix);
    Upload(b200_pick_device(comm_spec.local_id()));
  }

  void PrepareToRunApp(const CommSpec& comm_spec, PrepareConf conf, const ParallelEngineSpec& pe_spec) {
    base_t::PrepareToRunApp(comm_spec, conf, pe_spec);
    // split positions / outer ranges are part of the device layout already;
    // host_fragment.h:205-215: the device vertex map is built on request
    if (conf.need_build_device_vm && !d_vm_) {
      d_vm_ = std::make_shared<dev_vertex_map_t>();
      Stream stream;
      d_vm_->Init(stream, comm_spec, this->vm_ptr_);
    }
  }

  device_t DeviceObject() const {
    gl_frag_view v;
    CHECK_GL(gl_frag_view_get(handle_, &v));
    return d_vm_ ? device_t(v, d_vm_->DeviceObject()) : device_t(v);
  }
  // host_fragment.h:496-522: the inner vertices' out-edges as a device edge list.
  // release_csr is accepted and ignored: the SoA CSR stays resident (the apps
  // that convert keep using the fragment's id accessors).
  std::shared_ptr<coo_t> ConvertToCOO(bool release_csr = false) {
    (void) release_csr;
    if (!coo_frag_) {
      thrust::host_vector<typename coo_t::edge_t> edges;
      edges.reserve(this->GetEdgeNum());
      for (auto u : this->InnerVertices())
        for (auto& e : this->GetOutgoingAdjList(u))
          edges.push_back(typename coo_t::edge_t(u.GetValue(), e.get_neighbor().GetValue(), e.get_data()));
      coo_frag_ = std::make_shared<coo_t>();
      coo_frag_->Init(edges);
    }
    return coo_frag_;
  }
  gl_frag_t* handle() const { return handle_; }
  void OffloadTopology() const { CHECK_GL(gl_frag_offload(handle_)); }
  void ReloadTopology() const { CHECK_GL(gl_frag_reload(handle_)); }

 private:
  template <typename E>
  static typename std::enable_if<std::is_same<E, EmptyType>::value>::type put_w(std::vector<E>&, const nbr_t&) {}
  template <typename E>
  static typename std::enable_if<!std::is_same<E, EmptyType>::value>::type put_w(std::vector<E>& w, const nbr_t& n) {
    w.push_back(n.get_data());
  }

  // AoS Nbr rows of the CPU fragment -> SoA (row_ptr, col, w) -> gl_frag_create
  void Upload(int local_id) {
    CHECK_CUDA(cudaSetDevice(local_id));   // run_cuda_app.h:207-214
    const VID_T ivnum = this->GetInnerVerticesNum();
    const VID_T ovnum = this->GetOuterVerticesNum();
    auto iv = this->InnerVertices();
    std::vector<uint64_t> orp(ivnum + 1, 0), irp(ivnum + 1, 0);
    std::vector<uint32_t> ocol, icol;
    std::vector<EDATA_T> ow, iw;
    size_t k = 0;
    for (auto v : iv) {
      for (auto& e : this->GetOutgoingAdjList(v)) {
        ocol.push_back(e.get_neighbor().GetValue());
        put_w<EDATA_T>(ow, e);
      }
      orp[++k] = ocol.size();
    }
    const bool both = _load_strategy == grape::LoadStrategy::kBothOutIn && this->directed();
    if (both) {
      k = 0;
      for (auto v : iv) {
        for (auto& e : this->GetIncomingAdjList(v)) {
          icol.push_back(e.get_neighbor().GetValue());
          put_w<EDATA_T>(iw, e);
        }
        irp[++k] = icol.size();
      }
    }
    std::vector<uint32_t> ovgid(ovnum);
    for (auto v : this->OuterVertices()) ovgid[v.GetValue() - ivnum] = this->GetOuterVertexGid(v);
    std::vector<int64_t> oids(ivnum);
    for (auto v : iv) oids[v.GetValue()] = (int64_t) this->GetId(v);
    gl_frag_desc d;
    memset(&d, 0, sizeof(d));
    d.fid = this->fid();
    d.fnum = this->fnum();
    d.directed = this->directed() ? 1 : 0;
    d.load_strategy = both ? GL_LOAD_BOTH_OUT_IN : GL_LOAD_ONLY_OUT;
    d.ivnum = ivnum;
    d.ovnum = ovnum;
    d.total_vnum = this->GetTotalVerticesNum();
    d.edata_bytes = std::is_same<EDATA_T, EmptyType>::value ? 0 : (int) sizeof(EDATA_T);
    d.oe.row_ptr = orp.data();
    d.oe.col = ocol.data();
    d.oe.edata = d.edata_bytes ? (const void*) ow.data() : nullptr;
    d.oe.rows = ivnum;
    if (both) {
      d.ie.row_ptr = irp.data();
      d.ie.col = icol.data();
      d.ie.edata = d.edata_bytes ? (const void*) iw.data() : nullptr;
      d.ie.rows = ivnum;
    }
    d.ovgid = ovgid.data();
    d.inner_oids = oids.data();
    if (handle_) gl_frag_destroy(handle_);
    CHECK_GL(gl_frag_create(&handle_, &d));
  }

  gl_frag_t* handle_ = nullptr;
  std::shared_ptr<coo_t> coo_frag_;
  std::shared_ptr<dev_vertex_map_t> d_vm_;
};

// ---------------------------------------------------------- message manager --
// grape/cuda/parallel/gpu_message_manager.h:45-458 on the C ABI's gl_mm_*
// (include/grape_b200.h): the InArchive a producer appends to IS the
// destination fragment's landing slot, mapped over NVLink / CUDA IPC -- the
// store is the transfer.  No staging archive, no ncclSend/ncclRecv, no host
// MPI_Allgather of sizes; FinishARound = one device-side barrier + vote.

namespace dev {
// dev::OutArchive (serialization/out_archive.h:32-89): what one source fragment
// sent me in the previous round
class OutArchive {
 public:
  OutArchive() = default;
  DEV_HOST_INLINE OutArchive(const char* data, uint32_t size) : data_(data), size_(size) {}
  DEV_HOST_INLINE uint32_t size() const { return size_; }
  DEV_HOST_INLINE const char* data() const { return data_; }
  DEV_HOST_INLINE bool Empty() const { return size_ == 0; }

 private:
  const char* data_ = nullptr;
  uint32_t size_ = 0;
};

// dev::InArchive (serialization/in_archive.h:36-103): append-only byte buffer
// of ONE destination fragment
class InArchive {
 public:
  InArchive() = default;
  DEV_HOST_INLINE InArchive(char* slot, uint32_t* bytes, uint32_t cap) : slot_(slot), bytes_(bytes), cap_(cap) {}
  template <typename T>
  DEV_INLINE void AddBytes(const T& elem) {
    const uint32_t off = atomicAdd(bytes_, (uint32_t) sizeof(T));
    if (off + sizeof(T) <= cap_) *reinterpret_cast<T*>(slot_ + off) = elem;
  }
  // one reservation per warp (all active lanes append to THIS archive)
  template <typename T>
  DEV_INLINE void AddBytesWarp(const T& elem) {
    const uint32_t active = __activemask();
    const uint32_t leader = __ffs(active) - 1;
    uint32_t base = 0;
    if ((threadIdx.x & 31) == leader) base = atomicAdd(bytes_, (uint32_t) (__popc(active) * sizeof(T)));
    base = __shfl_sync(active, base, leader);
    const uint32_t off = base + __popc(active & ((1u << (threadIdx.x & 31)) - 1)) * (uint32_t) sizeof(T);
    if (off + sizeof(T) <= cap_) *reinterpret_cast<T*>(slot_ + off) = elem;
  }

 private:
  char* slot_ = nullptr;
  uint32_t* bytes_ = nullptr;
  uint32_t cap_ = 0;
};

class MessageManager {
 public:
  // The object holds a pointer to a DEVICE-resident gl_mm_view that the host side refreshes
  // whenever the slots change (every StartARound, every InitBuffer): like the reference's
  // ArrayView over d_to_send_ (gpu_message_manager.h:342-344), a DeviceObject() taken before a
  // re-InitBuffer in the same round (lcc.h:392 vs :416) keeps working.
  MessageManager() = default;
  explicit MessageManager(const gl_mm_view* d_view) : mvp_(d_view) {}

  // (gid of the outer vertex, msg) to the vertex's owner (:53-82)
  template <typename GRAPH_T, typename MESSAGE_T>
  DEV_INLINE void SyncStateOnOuterVertex(const GRAPH_T& frag, const typename GRAPH_T::vertex_t& v,
                                         const MESSAGE_T& msg) {
    archive(frag.GetFragId(v)).AddBytes(thrust::make_pair(frag.GetOuterVertexGid(v), msg));
  }
  template <typename GRAPH_T>
  DEV_INLINE void SyncStateOnOuterVertex(const GRAPH_T& frag, const typename GRAPH_T::vertex_t& v) {
    archive(frag.GetFragId(v)).AddBytes(frag.GetOuterVertexGid(v));
  }
  // lanes of a warp that target the same owner share one reservation
  // (in_archive.h:52-67 AddBytesWarpOpt)
  template <typename GRAPH_T, typename MESSAGE_T>
  DEV_INLINE void SyncStateOnOuterVertexWarpOpt(const GRAPH_T& frag, const typename GRAPH_T::vertex_t& v,
                                                const MESSAGE_T& msg) {
    warp_opt(frag.GetFragId(v), thrust::make_pair(frag.GetOuterVertexGid(v), msg));
  }
  template <typename GRAPH_T>
  DEV_INLINE void SyncStateOnOuterVertexWarpOpt(const GRAPH_T& frag, const typename GRAPH_T::vertex_t& v) {
    warp_opt(frag.GetFragId(v), frag.GetOuterVertexGid(v));
  }
  // (gid of the inner vertex, msg) to every fragment that holds a copy reachable
  // through in / out / both edge sets (:84-139; DestList = IEDests/OEDests/IOEDests)
  template <typename GRAPH_T, typename MESSAGE_T>
  DEV_INLINE void SendMsgThroughIEdges(const GRAPH_T& frag, const typename GRAPH_T::vertex_t& v,
                                       const MESSAGE_T& msg) {
    through(frag, v, 2, thrust::make_pair(frag.GetInnerVertexGid(v), msg));
  }
  template <typename GRAPH_T, typename MESSAGE_T>
  DEV_INLINE void SendMsgThroughOEdges(const GRAPH_T& frag, const typename GRAPH_T::vertex_t& v,
                                       const MESSAGE_T& msg) {
    through(frag, v, 1, thrust::make_pair(frag.GetInnerVertexGid(v), msg));
  }
  template <typename GRAPH_T>
  DEV_INLINE void SendMsgThroughOEdges(const GRAPH_T& frag, const typename GRAPH_T::vertex_t& v) {
    through(frag, v, 1, frag.GetInnerVertexGid(v));
  }
  template <typename GRAPH_T, typename MESSAGE_T>
  DEV_INLINE void SendMsgThroughEdges(const GRAPH_T& frag, const typename GRAPH_T::vertex_t& v,
                                      const MESSAGE_T& msg) {
    through(frag, v, 3, thrust::make_pair(frag.GetInnerVertexGid(v), msg));
  }
  template <typename MESSAGE_T>
  DEV_INLINE void SendToFragment(fid_t dst_fid, const MESSAGE_T& msg) {
    archive(dst_fid).AddBytes(msg);
  }
  template <typename MESSAGE_T>
  DEV_INLINE void SendToFragmentWarpOpt(fid_t dst_fid, const MESSAGE_T& msg) {
    warp_opt(dst_fid, msg);
  }

 private:
  DEV_INLINE InArchive archive(fid_t f) const {
    const gl_mm_view& mv_ = *mvp_;
    return InArchive(mv_.send_slot[f], mv_.send_bytes + f, mv_.capacity_bytes);
  }
  template <typename T>
  DEV_INLINE void warp_opt(fid_t f, const T& item) const {
    const gl_mm_view& mv_ = *mvp_;
    const uint32_t active = __activemask();
    const uint32_t peers = __match_any_sync(active, f);
    const uint32_t leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((threadIdx.x & 31) == leader) base = atomicAdd(mv_.send_bytes + f, (uint32_t) (__popc(peers) * sizeof(T)));
    base = __shfl_sync(peers, base, leader);
    const uint32_t off = base + __popc(peers & ((1u << (threadIdx.x & 31)) - 1)) * (uint32_t) sizeof(T);
    if (off + sizeof(T) <= mv_.capacity_bytes) *reinterpret_cast<T*>(mv_.send_slot[f] + off) = item;
  }
  // Destination fragments of v = owners of the outer neighbours in the chosen
  // edge sets.  The outer part of a row ([split, end)) is sorted by gid, i.e.
  // grouped by owner, so the owner set is read off the row itself -- the
  // reference precomputes the same set as fid lists (idst_/odst_/iodst_,
  // immutable_edgecut_fragment.h:606-722).  fnum <= 64: a 64-bit set.
  template <typename GRAPH_T, typename T>
  DEV_INLINE void through(const GRAPH_T& frag, const typename GRAPH_T::vertex_t& v, int which, const T& item) const {
    const gl_frag_view& fv = frag.view();
    const uint32_t u = v.GetValue();
    unsigned long long set = 0;
    if (which & 1)
      for (uint64_t p = fv.oe_split[u]; p < fv.oe_rp[u + 1]; ++p)
        set |= 1ull << (fv.ovgid[fv.oe_col[p] - fv.ivnum] >> fv.fid_offset);
    if ((which & 2) && !((which & 1) && fv.ie_col == fv.oe_col))
      for (uint64_t p = fv.ie_split[u]; p < fv.ie_rp[u + 1]; ++p)
        set |= 1ull << (fv.ovgid[fv.ie_col[p] - fv.ivnum] >> fv.fid_offset);
    while (set) {
      const fid_t f = (fid_t) (__ffsll((long long) set) - 1);
      set &= set - 1;
      archive(f).AddBytes(item);
    }
  }
  const gl_mm_view* mvp_ = nullptr;
};

// message_kernels.h:28-127 ProcessMsg: apply func to every received unit
template <typename GRAPH_T, typename MESSAGE_T, typename FUNC_T>
__global__ void ProcessMsg(gl_mm_view mv, const GRAPH_T frag, FUNC_T func) {
  using vid_t = typename GRAPH_T::vid_t;
  for (fid_t src = 0; src < mv.fnum; ++src) {
    if (src == mv.fid) continue;
    uint32_t bytes = mv.recv_bytes[src];
    if (bytes > mv.capacity_bytes) bytes = mv.capacity_bytes;
    if constexpr (std::is_same<MESSAGE_T, grape::EmptyType>::value) {
      const vid_t* units = reinterpret_cast<const vid_t*>(mv.recv_slot[src]);
      const size_t n = bytes / sizeof(vid_t);
      for (size_t i = TID_1D; i < n; i += TOTAL_THREADS_1D) {
        typename GRAPH_T::vertex_t v;
        bool ok = frag.Gid2Vertex(units[i], v);
        assert(ok);
        (void) ok;
        func(v);
      }
    } else {
      using unit_t = thrust::pair<vid_t, MESSAGE_T>;
      const unit_t* units = reinterpret_cast<const unit_t*>(mv.recv_slot[src]);
      const size_t n = bytes / sizeof(unit_t);
      for (size_t i = TID_1D; i < n; i += TOTAL_THREADS_1D) {
        const unit_t unit = units[i];
        typename GRAPH_T::vertex_t v;
        bool ok = frag.Gid2Vertex(unit.first, v);
        assert(ok);
        (void) ok;
        func(v, unit.second);
      }
    }
  }
}
template <typename MESSAGE_T, typename FUNC_T>
__global__ void ProcessRawMsg(gl_mm_view mv, FUNC_T func) {
  for (fid_t src = 0; src < mv.fnum; ++src) {
    if (src == mv.fid) continue;
    uint32_t bytes = mv.recv_bytes[src];
    if (bytes > mv.capacity_bytes) bytes = mv.capacity_bytes;
    const MESSAGE_T* units = reinterpret_cast<const MESSAGE_T*>(mv.recv_slot[src]);
    const size_t n = bytes / sizeof(MESSAGE_T);
    for (size_t i = TID_1D; i < n; i += TOTAL_THREADS_1D) func(units[i]);
  }
}
}  // namespace dev

class GPUMessageManager {
 public:
  GPUMessageManager() = default;
  ~GPUMessageManager() {
    Release(false);
    if (d_view_) cudaFree(d_view_);
  }
  GPUMessageManager(const GPUMessageManager&) = delete;
  GPUMessageManager& operator=(const GPUMessageManager&) = delete;

  // :160-194.  The NCCL communicator of the reference is replaced by a
  // fragment-group communicator whose landing areas are mapped at InitBuffer
  // time (the slot capacity is only known then).
  void Init(const grape::CommSpec& comm_spec) {
    comm_spec_ = comm_spec;
    fnum_ = comm_spec.fnum();
    fid_ = comm_spec.fid();
    CHECK_CUDA(cudaSetDevice(b200_pick_device(comm_spec.local_id())));
    if (fnum_ == 1) CHECK_GL(gl_mm_create(&mm_, nullptr));
  }

  // :196-204.  Collective: every fragment's app context calls it with sizes
  // derived from its own vertex counts; the group agrees on the largest.
  void InitBuffer(size_t send_buffer_capacity, size_t recv_buffer_capacity) {
    if (fnum_ == 1) return;
    unsigned long long need = std::max(send_buffer_capacity, recv_buffer_capacity) + 256, all = 0;
    MPI_Allreduce(&need, &all, 1, MPI_UNSIGNED_LONG_LONG, MPI_MAX, comm_spec_.comm());
    if (mm_ && all <= capacity_) return;
    Release(true);
    gl_comm_desc d;
    memset(&d, 0, sizeof(d));
    d.fid = fid_;
    d.fnum = fnum_;
    d.allreduce = &GPUMessageManager::HostAllReduce;
    d.user = this;
    d.landing_bytes = (size_t) all;
    d.mirror_bytes = 0;
    CHECK_GL(gl_comm_create(&comm_, &d));
    std::vector<char> mine(GL_IPC_HANDLE_BYTES), handles((size_t) GL_IPC_HANDLE_BYTES * fnum_);
    CHECK_GL(gl_comm_export(comm_, mine.data(), mine.size()));
    MPI_Allgather(mine.data(), GL_IPC_HANDLE_BYTES, MPI_CHAR, handles.data(), GL_IPC_HANDLE_BYTES, MPI_CHAR,
                  comm_spec_.comm());
    CHECK_GL(gl_comm_open(comm_, handles.data(), handles.size()));
    CHECK_GL(gl_mm_create(&mm_, comm_));
    CHECK_GL(gl_mm_init_buffer(mm_, send_buffer_capacity, recv_buffer_capacity));
    capacity_ = all;
    MPI_Barrier(comm_spec_.comm());   // every landing area is mapped before the first round
    // a re-InitBuffer in the middle of a round (lcc.h:416): the new manager joins the round in progress
    if (round_started_) CHECK_GL(gl_mm_start_round(mm_, stream_.cuda_stream()));
    if (force_continue_) CHECK_GL(gl_mm_force_continue(mm_));
    RefreshDeviceView();
  }
  void DropBuffer() { Release(true); }

  void Start() {
    if (mm_) CHECK_GL(gl_mm_start(mm_));
    round_started_ = false;
    force_continue_ = false;
  }
  void StartARound() {
    EnsureManager();
    CHECK_GL(gl_mm_start_round(mm_, stream_.cuda_stream()));
    round_started_ = true;
    force_continue_ = false;
    RefreshDeviceView();
  }
  void FinishARound() {
    EnsureManager();
    CHECK_GL(gl_mm_finish_round(mm_, stream_.cuda_stream()));
    int t = 0;
    CHECK_GL(gl_mm_to_terminate(mm_, &t));
    terminate_ = t != 0;
    round_started_ = false;
  }
  void Finalize() const {}
  bool ToTerminate() const { return terminate_; }
  void ForceContinue() {
    EnsureManager();
    CHECK_GL(gl_mm_force_continue(mm_));
    force_continue_ = true;
  }
  size_t GetMsgSize() const { return mm_ ? (size_t) gl_mm_bytes_sent(mm_) : 0; }
  double GetAccumulatedCommTime() const { return 0.0; }
  Stream& stream() { return stream_; }
  void* nccl_comm() { return nullptr; }
  gl_mm_t* handle() const { return mm_; }
  dev::MessageManager DeviceObject() {
    EnsureManager();
    if (!d_view_) RefreshDeviceView();
    return dev::MessageManager(d_view_);
  }
  // :362-393
  template <typename GRAPH_T, typename MESSAGE_T = grape::EmptyType, typename FUNC_T>
  void ParallelProcess(const GRAPH_T& frag, FUNC_T func) {
    if (fnum_ == 1) return;   // nothing is ever received
    gl_mm_view mv;
    CHECK_GL(gl_mm_view_get(mm_, &mv));
    dev::ProcessMsg<GRAPH_T, MESSAGE_T, FUNC_T><<<256, 256, 0, stream_.cuda_stream()>>>(mv, frag, func);
    CHECK_CUDA(cudaGetLastError());
    stream_.Sync();
  }
  template <typename MESSAGE_T = grape::EmptyType, typename FUNC_T>
  void ParallelProcess(FUNC_T func) {
    if (fnum_ == 1) return;
    gl_mm_view mv;
    CHECK_GL(gl_mm_view_get(mm_, &mv));
    dev::ProcessRawMsg<MESSAGE_T, FUNC_T><<<256, 256, 0, stream_.cuda_stream()>>>(mv, func);
    CHECK_CUDA(cudaGetLastError());
    stream_.Sync();
  }

 private:
  static int HostAllReduce(void* user, void* inout, int n, int is_double, int op) {
    auto* self = static_cast<GPUMessageManager*>(user);
    const MPI_Op mop = op == 0 ? MPI_SUM : (op == 1 ? MPI_MIN : MPI_MAX);
    return MPI_Allreduce(MPI_IN_PLACE, inout, n, is_double ? MPI_DOUBLE : MPI_INT64_T, mop, self->comm_spec_.comm());
  }
  void EnsureManager() {
    // an app that never called InitBuffer still gets a working round protocol
    if (!mm_) InitBuffer(4096, 4096);
  }
  void RefreshDeviceView() {
    if (!mm_) return;
    if (!d_view_) CHECK_CUDA(cudaMalloc(&d_view_, sizeof(gl_mm_view)));
    gl_mm_view mv;
    CHECK_GL(gl_mm_view_get(mm_, &mv));
    CHECK_CUDA(cudaMemcpyAsync(d_view_, &mv, sizeof(mv), cudaMemcpyHostToDevice, stream_.cuda_stream()));
    stream_.Sync();
  }
  void Release(bool collective) {
    if (mm_) {
      stream_.Sync();
      gl_mm_destroy(mm_);
      mm_ = nullptr;
    }
    if (comm_) {
      // unmap the peers' landing areas, agree that everybody did, then free mine
      gl_comm_close_peers(comm_);
      if (collective && fnum_ > 1) MPI_Barrier(comm_spec_.comm());
      gl_comm_destroy(comm_);
      comm_ = nullptr;
    }
    capacity_ = 0;
    if (fnum_ == 1 && collective) CHECK_GL(gl_mm_create(&mm_, nullptr));
  }

  grape::CommSpec comm_spec_;
  fid_t fid_ = 0, fnum_ = 1;
  gl_comm_t* comm_ = nullptr;
  gl_mm_t* mm_ = nullptr;
  unsigned long long capacity_ = 0;
  Stream stream_;
  bool terminate_ = false;
  bool round_started_ = false;
  bool force_continue_ = false;
  gl_mm_view* d_view_ = nullptr;
};

// ------------------------------------------------- batch shuffle (dense sync) --
// grape/cuda/parallel/batch_shuffle_message_manager.h:42-273,
// grape/cuda/app/batch_shuffle_app_base.h:39-92,
// grape/cuda/worker/gpu_batch_shuffle_worker.h:33-120: the message strategy in
// which an owner pushes the state of ALL its mirrored inner vertices to the
// fragments that hold copies.  On the C ABI's dense mirror sync
// (gl_mm_mirror_plan / gl_mm_sync_values_to_ghosts): owners write straight into
// the holders' mirror slots over NVLink, no ncclSend/ncclRecv staging buffers.
class BatchShuffleMessageManager {
 public:
  BatchShuffleMessageManager() = default;
  ~BatchShuffleMessageManager() { Release(); }
  BatchShuffleMessageManager(const BatchShuffleMessageManager&) = delete;
  BatchShuffleMessageManager& operator=(const BatchShuffleMessageManager&) = delete;

  void Init(const grape::CommSpec& comm_spec) {
    comm_spec_ = comm_spec;
    fnum_ = comm_spec.fnum();
    fid_ = comm_spec.fid();
    CHECK_CUDA(cudaSetDevice(b200_pick_device(comm_spec.local_id())));
  }
  void Start() {}
  void StartARound() {
    to_terminate_ = true;
    sent_size_ = 0;
  }
  void FinishARound() {}
  void Finalize() const {}
  // :142-220.  DATA_T must be 4 or 8 bytes wide (the C ABI's value sync).
  template <typename GRAPH_T, typename DATA_T>
  void SyncInnerVertices(const GRAPH_T& h_frag, VertexArray<DATA_T, typename GRAPH_T::vid_t>& h_data) {
    static_assert(sizeof(DATA_T) == 4 || sizeof(DATA_T) == 8, "SyncInnerVertices: 4- or 8-byte vertex data");
    to_terminate_ = false;
    if (fnum_ == 1) return;
    Ensure(h_frag);
    auto d_data = h_data.DeviceObject();
    CHECK_GL(gl_mm_sync_values_to_ghosts(mm_, stream_.cuda_stream(), (void*) d_data.data(), (int) sizeof(DATA_T)));
    stream_.Sync();
    sent_size_ += (size_t) h_frag.GetInnerVerticesNum() * sizeof(DATA_T);
  }
  bool ToTerminate() const { return to_terminate_; }
  size_t GetMsgSize() const { return sent_size_; }
  void ForceContinue() { to_terminate_ = false; }
  Stream& stream() { return stream_; }
  void* nccl_comm() { return nullptr; }
  double GetAccumulatedCommTime() const { return 0.0; }

 private:
  static int HostAllReduce(void* user, void* inout, int n, int is_double, int op) {
    auto* self = static_cast<BatchShuffleMessageManager*>(user);
    const MPI_Op mop = op == 0 ? MPI_SUM : (op == 1 ? MPI_MIN : MPI_MAX);
    return MPI_Allreduce(MPI_IN_PLACE, inout, n, is_double ? MPI_DOUBLE : MPI_INT64_T, mop, self->comm_spec_.comm());
  }
  template <typename GRAPH_T>
  void Ensure(const GRAPH_T& h_frag) {
    if (mm_) return;
    unsigned long long iv = h_frag.GetInnerVerticesNum(), mx = 0;
    MPI_Allreduce(&iv, &mx, 1, MPI_UNSIGNED_LONG_LONG, MPI_MAX, comm_spec_.comm());
    gl_comm_desc d;
    memset(&d, 0, sizeof(d));
    d.fid = fid_;
    d.fnum = fnum_;
    d.allreduce = &BatchShuffleMessageManager::HostAllReduce;
    d.user = this;
    d.landing_bytes = 4096;
    d.mirror_bytes = (size_t) 8 * (mx + 1024);
    CHECK_GL(gl_comm_create(&comm_, &d));
    std::vector<char> mine(GL_IPC_HANDLE_BYTES), handles((size_t) GL_IPC_HANDLE_BYTES * fnum_);
    CHECK_GL(gl_comm_export(comm_, mine.data(), mine.size()));
    MPI_Allgather(mine.data(), GL_IPC_HANDLE_BYTES, MPI_CHAR, handles.data(), GL_IPC_HANDLE_BYTES, MPI_CHAR,
                  comm_spec_.comm());
    CHECK_GL(gl_comm_open(comm_, handles.data(), handles.size()));
    CHECK_GL(gl_mm_create(&mm_, comm_));
    MPI_Barrier(comm_spec_.comm());
    CHECK_GL(gl_mm_mirror_plan(mm_, stream_.cuda_stream(), h_frag.handle()));
  }
  void Release() {
    if (mm_) {
      gl_mm_destroy(mm_);
      mm_ = nullptr;
    }
    if (comm_) {
      gl_comm_close_peers(comm_);
      if (fnum_ > 1) MPI_Barrier(comm_spec_.comm());
      gl_comm_destroy(comm_);
      comm_ = nullptr;
    }
  }
  grape::CommSpec comm_spec_;
  fid_t fid_ = 0, fnum_ = 1;
  gl_comm_t* comm_ = nullptr;
  gl_mm_t* mm_ = nullptr;
  Stream stream_;
  size_t sent_size_ = 0;
  bool to_terminate_ = false;
};

template <typename APP_T>
class GPUBatchShuffleWorker;

template <typename FRAG_T, typename CONTEXT_T>
class BatchShuffleAppBase {
 public:
  static constexpr bool need_split_edges = false;
  static constexpr bool need_build_device_vm = false;
  static constexpr grape::MessageStrategy message_strategy = grape::MessageStrategy::kSyncOnOuterVertex;
  static constexpr grape::LoadStrategy load_strategy = grape::LoadStrategy::kOnlyOut;
  using message_manager_t = BatchShuffleMessageManager;
  BatchShuffleAppBase() = default;
  virtual ~BatchShuffleAppBase() = default;
  virtual void PEval(const FRAG_T& graph, CONTEXT_T& context, message_manager_t& messages) = 0;
  virtual void IncEval(const FRAG_T& graph, CONTEXT_T& context, message_manager_t& messages) = 0;
};

#define INSTALL_GPU_BATCH_SHUFFLE_WORKER(APP_T, CONTEXT_T, FRAG_T)   \
 public:                                                             \
  using fragment_t = FRAG_T;                                         \
  using context_t = CONTEXT_T;                                       \
  using worker_t = grape::cuda::GPUBatchShuffleWorker<APP_T>;        \
  using message_manager_t = grape::cuda::BatchShuffleMessageManager; \
  using dev_message_manager_t = grape::cuda::dev::MessageManager;    \
  virtual ~APP_T() {}                                                \
  static std::shared_ptr<worker_t> CreateWorker(                     \
      std::shared_ptr<APP_T> app, std::shared_ptr<FRAG_T> frag) {    \
    return std::shared_ptr<worker_t>(new worker_t(app, frag));       \
  }

template <typename APP_T>
class GPUBatchShuffleWorker {
 public:
  using fragment_t = typename APP_T::fragment_t;
  using context_t = typename APP_T::context_t;
  using message_manager_t = BatchShuffleMessageManager;
  GPUBatchShuffleWorker(std::shared_ptr<APP_T> app, std::shared_ptr<fragment_t> graph)
      : app_(std::move(app)), context_(std::make_shared<context_t>(*graph)), messages_() {}
  template <class... Args>
  void Init(const grape::CommSpec& comm_spec, Args&&... args) {
    auto& graph = const_cast<fragment_t&>(context_->fragment());
    PrepareConf conf;
    conf.message_strategy = APP_T::message_strategy;
    conf.need_split_edges = APP_T::need_split_edges;
    conf.need_split_edges_by_fragment = false;
    conf.need_mirror_info = true;
    conf.need_build_device_vm = APP_T::need_build_device_vm;
    graph.PrepareToRunApp(comm_spec, conf, grape::DefaultParallelEngineSpec());
    comm_spec_ = comm_spec;
    messages_.Init(comm_spec);
    InitCommunicator(app_, comm_spec.comm(), messages_.nccl_comm());
    context_->Init(messages_, std::forward<Args>(args)...);
  }
  void Finalize() {}
  void Query() {
    auto& graph = context_->fragment();
    messages_.Start();
    messages_.StartARound();
    app_->PEval(graph, *context_, messages_);
    messages_.FinishARound();
    MPI_Barrier(comm_spec_.comm());
    int step = 1;
    while (!messages_.ToTerminate()) {
      messages_.StartARound();
      app_->IncEval(graph, *context_, messages_);
      messages_.FinishARound();
      MPI_Barrier(comm_spec_.comm());
      ++step;
    }
    supersteps_ = step;
    messages_.Finalize();
  }
  int supersteps() const { return supersteps_; }
  std::shared_ptr<context_t> GetContext() { return context_; }
  void Output(std::ostream& os) { context_->Output(os); }

 private:
  std::shared_ptr<APP_T> app_;
  std::shared_ptr<context_t> context_;
  message_manager_t messages_;
  grape::CommSpec comm_spec_;
  int supersteps_ = 0;
};

// -------------------------------------------------------------- communicator --
// grape/cuda/communication/communicator.h:41-95: one scalar per rank.  The
// reference reduces host scalars with MPI and device scalars with NCCL; here
// both go through MPI on the host (one word per superstep at most).
class Communicator {
 public:
  Communicator() = default;
  virtual ~Communicator() = default;
  void InitCommunicator(MPI_Comm comm, void*) { comm_ = comm; }
  template <typename T>
  void Sum(T msg_in, T& msg_out) { reduce(msg_in, msg_out, MPI_SUM); }
  template <typename T>
  void Min(T msg_in, T& msg_out) { reduce(msg_in, msg_out, MPI_MIN); }
  template <typename T>
  void Max(T msg_in, T& msg_out) { reduce(msg_in, msg_out, MPI_MAX); }
  template <typename T>
  void Sum(T msg_in, T& msg_out, const Stream& stream) { stream.Sync(); reduce(msg_in, msg_out, MPI_SUM); }
  template <typename T>
  void Min(T msg_in, T& msg_out, const Stream& stream) { stream.Sync(); reduce(msg_in, msg_out, MPI_MIN); }
  template <typename T>
  void Max(T msg_in, T& msg_out, const Stream& stream) { stream.Sync(); reduce(msg_in, msg_out, MPI_MAX); }
  template <typename T>
  std::vector<T> AllGather(T msg_in) {
    int n = 1;
    MPI_Comm_size(comm_, &n);
    std::vector<T> out((size_t) n);
    MPI_Allgather(&msg_in, (int) sizeof(T), MPI_CHAR, out.data(), (int) sizeof(T), MPI_CHAR, comm_);
    return out;
  }

 private:
  template <typename T>
  static MPI_Datatype mpi_type() {
    static_assert(std::is_arithmetic<T>::value && !std::is_same<T, bool>::value, "unsupported type");
    if (std::is_floating_point<T>::value) return sizeof(T) == 8 ? MPI_DOUBLE : MPI_FLOAT;
    if (std::is_signed<T>::value) return sizeof(T) == 8 ? MPI_INT64_T : (sizeof(T) == 4 ? MPI_INT : MPI_INT8_T);
    return sizeof(T) == 8 ? MPI_UINT64_T : (sizeof(T) == 4 ? MPI_UINT32_T : MPI_UINT8_T);
  }
  template <typename T>
  void reduce(T in, T& out, MPI_Op op) {
    out = in;
    MPI_Allreduce(MPI_IN_PLACE, &out, 1, mpi_type<T>(), op, comm_);
  }
  MPI_Comm comm_ = MPI_COMM_WORLD;
};
template <typename APP_T>
typename std::enable_if<std::is_base_of<Communicator, APP_T>::value>::type InitCommunicator(
    std::shared_ptr<APP_T> app, MPI_Comm comm, void* nccl) {
  app->InitCommunicator(comm, nccl);
}
template <typename APP_T>
typename std::enable_if<!std::is_base_of<Communicator, APP_T>::value>::type InitCommunicator(
    std::shared_ptr<APP_T>, MPI_Comm, void*) {}

// ------------------------------------------------------------ parallel engine --
// grape/cuda/parallel/parallel_engine.h:51-70, 72-293, 987-1182
enum class LoadBalancing { kCMOld, kCM, kWarp, kCTA, kStrict, kNone };
inline LoadBalancing ParseLoadBalancing(const std::string& s) {
  if (s == "cmold" || s == "CMOLD") return LoadBalancing::kCMOld;
  if (s == "cm" || s == "CM") return LoadBalancing::kCM;
  if (s == "cta" || s == "CTA") return LoadBalancing::kCTA;
  if (s == "wm" || s == "WM") return LoadBalancing::kWarp;
  if (s == "strict" || s == "STRICT") return LoadBalancing::kStrict;
  if (s == "none" || s == "NONE") return LoadBalancing::kNone;
  LOG(FATAL) << "Invalid lb: " + s;
  return LoadBalancing::kNone;
}

template <typename VID_T, typename METADATA_T>
struct VertexMetadata {
  Vertex<VID_T> vertex;
  METADATA_T metadata;
  DEV_HOST_INLINE void set_metadata(const METADATA_T& m) { metadata = m; }
};

namespace compat_detail {
template <typename WS>
struct SrcOf;
template <typename VID_T>
struct SrcOf<WorkSourceRange<Vertex<VID_T>>> {
  using type = ::gl::RangeSrc;
  static type make(const WorkSourceRange<Vertex<VID_T>>& ws) { return type{(uint32_t) ws.start_.GetValue()}; }
};
template <typename VID_T>
struct SrcOf<WorkSourceArray<Vertex<VID_T>>> {
  using type = ::gl::ArraySrc;
  static type make(const WorkSourceArray<Vertex<VID_T>>& ws) {
    static_assert(sizeof(Vertex<VID_T>) == sizeof(uint32_t), "32-bit vertex ids");
    return type{reinterpret_cast<const uint32_t*>(ws.data_)};
  }
};

template <typename EDATA_T>
struct WeightOf {
  using type = EDATA_T;
  static constexpr bool weighted = true;
};
template <>
struct WeightOf<EmptyType> {
  using type = float;
  static constexpr bool weighted = false;
};

template <typename VID_T, typename EDATA_T, typename W>
DEV_INLINE dev::PosNbr<VID_T, EDATA_T> make_nbr(uint32_t v, W w, uint64_t pos, std::false_type) {
  return dev::PosNbr<VID_T, EDATA_T>((VID_T) v, (EDATA_T) w, pos);
}
template <typename VID_T, typename EDATA_T, typename W>
DEV_INLINE dev::PosNbr<VID_T, EDATA_T> make_nbr(uint32_t v, W, uint64_t pos, std::true_type) {
  return dev::PosNbr<VID_T, EDATA_T>((VID_T) v, pos);
}

// adapter: (assign_op, edge_op(VertexMetadata, nbr)) -> engine Op.
// The metadata type is derived from the closure INSIDE the class (like the
// reference does inside its __device__ LB functions, parallel_engine.h:631):
// it must not appear among the kernel's template arguments, because the
// return type of an extended __device__ lambda is not visible to host code and
// a host/device mismatch there changes the kernel's mangled name.
template <typename FRAG_T, typename ASSIGN_OP, typename EDGE_OP>
struct MetaOp {
  using vid_t = typename FRAG_T::vid_t;
  using edata_t = typename FRAG_T::edata_t;
  using Meta = typename std::result_of<ASSIGN_OP&(Vertex<vid_t>&)>::type;
  using W = typename WeightOf<edata_t>::type;
  static constexpr bool kWeighted = WeightOf<edata_t>::weighted;
  ASSIGN_OP assign_op;
  EDGE_OP edge_op;
  __device__ Meta assign(uint32_t u) const {
    ASSIGN_OP a = assign_op;
    Vertex<vid_t> vu(u);
    return a(vu);
  }
  static constexpr bool kWantsPos = true;   // the nbr handed to the app knows its CSR position
  template <typename M>
  __device__ void edge_at(uint32_t u, const M& m, uint32_t v, W w, uint64_t pos, ::gl::ScanAcc&) const {
    EDGE_OP e = edge_op;
    VertexMetadata<vid_t, M> vm;
    vm.vertex = Vertex<vid_t>(u);
    vm.metadata = m;
    e(vm, make_nbr<vid_t, edata_t, W>(v, w, pos, std::is_same<edata_t, EmptyType>()));
  }
};
// adapter: edge_op(vertex, nbr) -> engine Op
template <typename FRAG_T, typename EDGE_OP>
struct PlainOp {
  using vid_t = typename FRAG_T::vid_t;
  using edata_t = typename FRAG_T::edata_t;
  using Meta = uint32_t;
  using W = typename WeightOf<edata_t>::type;
  static constexpr bool kWeighted = WeightOf<edata_t>::weighted;
  EDGE_OP edge_op;
  __device__ Meta assign(uint32_t) const { return 0; }
  static constexpr bool kWantsPos = true;   // the nbr handed to the app knows its CSR position
  __device__ void edge_at(uint32_t u, Meta, uint32_t v, W w, uint64_t pos, ::gl::ScanAcc&) const {
    EDGE_OP e = edge_op;
    e(Vertex<vid_t>(u), make_nbr<vid_t, edata_t, W>(v, w, pos, std::is_same<edata_t, EmptyType>()));
  }
};

template <typename WS, typename F>
__global__ void k_for_each(WS ws, F f) {
  for (size_t i = threadIdx.x + (size_t) blockIdx.x * blockDim.x; i < ws.size();
       i += (size_t) gridDim.x * blockDim.x)
    f(ws.GetWork(i));
}
template <typename WS, typename F>
__global__ void k_for_each_index(WS ws, F f) {
  for (size_t i = threadIdx.x + (size_t) blockIdx.x * blockDim.x; i < ws.size();
       i += (size_t) gridDim.x * blockDim.x)
    f(i, ws.GetWork(i));
}

// warp / CTA per work item (parallel_engine.h:93-271).  Static variants keep the
// reference's strided item -> warp / CTA assignment (the functors receive the
// warp or CTA id and the number of warps / CTAs and may partition `shm` with
// them); the *Dynamic variants draw the next item from a device ticket.
template <typename WS, typename F, typename... Args>
__global__ void __launch_bounds__(256) k_fe_warp(WS ws, F f, Args... args) {
  const size_t tid = threadIdx.x + (size_t) blockIdx.x * blockDim.x;
  const size_t n_warp = ((size_t) gridDim.x * blockDim.x) >> 5, warp_id = tid >> 5, lane = tid & 31;
  for (size_t i = warp_id; i < ws.size(); i += n_warp) f(lane, i, ws.GetWork(i), args...);
}
template <typename WS, typename F, typename... Args>
__global__ void __launch_bounds__(256) k_fe_warp_shared(WS ws, F f, Args... args) {
  __shared__ uint32_t shm[8192];
  const size_t tid = threadIdx.x + (size_t) blockIdx.x * blockDim.x;
  const size_t n_warp = ((size_t) gridDim.x * blockDim.x) >> 5, warp_id = tid >> 5, lane = tid & 31;
  for (size_t i = warp_id; i < ws.size(); i += n_warp)
    f(shm, lane, warp_id, (size_t) 32, n_warp, i, ws.GetWork(i), args...);
}
template <typename WS, typename F, typename... Args>
__global__ void __launch_bounds__(256) k_fe_warp_dynamic(WS ws, unsigned long long* ticket, F f, Args... args) {
  const size_t tid = threadIdx.x + (size_t) blockIdx.x * blockDim.x;
  const size_t lane = tid & 31;
  for (size_t i = tid >> 5; i < ws.size();) {
    f(lane, i, ws.GetWork(i), args...);
    __syncwarp();
    unsigned long long nx = 0;
    if (lane == 0) nx = atomicAdd(ticket, 1ull);
    i = (size_t) __shfl_sync(0xffffffffu, nx, 0);
  }
}
template <typename WS, typename F, typename... Args>
__global__ void __launch_bounds__(256) k_fe_block(WS ws, F f, Args... args) {
  for (size_t i = blockIdx.x; i < ws.size(); i += gridDim.x) {
    __syncthreads();
    f((size_t) threadIdx.x, i, ws.GetWork(i), args...);
    __syncthreads();
  }
}
template <typename WS, typename F, typename... Args>
__global__ void __launch_bounds__(256) k_fe_block_shared(WS ws, F f, Args... args) {
  __shared__ uint32_t shm[8192];
  for (size_t i = blockIdx.x; i < ws.size(); i += gridDim.x) {
    __syncthreads();
    f(shm, (size_t) threadIdx.x, (size_t) blockIdx.x, (size_t) blockDim.x, (size_t) gridDim.x, i, ws.GetWork(i), args...);
    __syncthreads();
  }
}
template <typename WS, typename F, typename... Args>
__global__ void __launch_bounds__(256) k_fe_block_dynamic(WS ws, unsigned long long* ticket, F f, Args... args) {
  __shared__ size_t shared_idx;
  for (size_t i = blockIdx.x; i < ws.size();) {
    __syncthreads();
    f((size_t) threadIdx.x, i, ws.GetWork(i), args...);
    __syncthreads();
    if (threadIdx.x == 0) shared_idx = (size_t) atomicAdd(ticket, 1ull);
    __syncthreads();
    i = shared_idx;
  }
}
static __global__ void k_set_ticket(unsigned long long* t, unsigned long long v) { *t = v; }
}  // namespace compat_detail

class ParallelEngine {
 public:
  ParallelEngine() {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms_, cudaDevAttrMultiProcessorCount, dev);
  }
  virtual ~ParallelEngine() {
    if (ctrl_) cudaFree(ctrl_);
    if (hubs_) cudaFree(hubs_);
    if (deg_) cudaFree(deg_);
    if (pfx_) cudaFree(pfx_);
    if (scan_tmp_) cudaFree(scan_tmp_);
    if (ticket_) cudaFree(ticket_);
  }

  // ForEach (parallel_engine.h:72-91)
  template <typename WORK_SOURCE_T, typename FUNC_T>
  void ForEach(const Stream& stream, const WORK_SOURCE_T& ws, FUNC_T f) {
    if (ws.size() == 0) return;
    unsigned grid = (unsigned) std::min<size_t>((ws.size() + 255) / 256, (size_t) sms_ * 8);
    compat_detail::k_for_each<<<grid, 256, 0, stream.cuda_stream()>>>(ws, f);
    CHECK_CUDA(cudaGetLastError());
  }
  // ForEachWithIndex (parallel_engine.h:273-293)
  template <typename WORK_SOURCE_T, typename FUNC_T>
  void ForEachWithIndex(const Stream& stream, const WORK_SOURCE_T& ws, FUNC_T f) {
    if (ws.size() == 0) return;
    unsigned grid = (unsigned) std::min<size_t>((ws.size() + 255) / 256, (size_t) sms_ * 8);
    compat_detail::k_for_each_index<<<grid, 256, 0, stream.cuda_stream()>>>(ws, f);
    CHECK_CUDA(cudaGetLastError());
  }


  // ForEachWithIndexWarp / WarpShared / WarpDynamic / Block / BlockShared /
  // BlockDynamic (parallel_engine.h:93-271): a warp (or a CTA) per work item.
  // Grids are sized from the SM count instead of the reference's fixed
  // <<<256,256>>> (launcher.h:47-53), which leaves 40 % of a B200 idle.
  template <typename WORK_SOURCE_T, typename FUNC_T, typename... Args>
  void ForEachWithIndexWarp(const Stream& stream, const WORK_SOURCE_T& ws, FUNC_T f, Args... args) {
    if (ws.size() == 0) return;
    compat_detail::k_fe_warp<<<WarpGrid(ws.size()), 256, 0, stream.cuda_stream()>>>(ws, f, args...);
    CHECK_CUDA(cudaGetLastError());
  }
  template <typename WORK_SOURCE_T, typename FUNC_T, typename... Args>
  void ForEachWithIndexWarpShared(const Stream& stream, const WORK_SOURCE_T& ws, FUNC_T f, Args... args) {
    if (ws.size() == 0) return;
    compat_detail::k_fe_warp_shared<<<SharedGrid((ws.size() + 7) / 8), 256, 0, stream.cuda_stream()>>>(ws, f, args...);
    CHECK_CUDA(cudaGetLastError());
  }
  template <typename WORK_SOURCE_T, typename FUNC_T, typename... Args>
  void ForEachWithIndexWarpDynamic(const Stream& stream, const WORK_SOURCE_T& ws, FUNC_T f, Args... args) {
    if (ws.size() == 0) return;
    const unsigned grid = WarpGrid(ws.size());
    compat_detail::k_set_ticket<<<1, 1, 0, stream.cuda_stream()>>>(Ticket(), (unsigned long long) grid * 8);
    compat_detail::k_fe_warp_dynamic<<<grid, 256, 0, stream.cuda_stream()>>>(ws, Ticket(), f, args...);
    CHECK_CUDA(cudaGetLastError());
  }
  template <typename WORK_SOURCE_T, typename FUNC_T, typename... Args>
  void ForEachWithIndexBlock(const Stream& stream, const WORK_SOURCE_T& ws, FUNC_T f, Args... args) {
    if (ws.size() == 0) return;
    compat_detail::k_fe_block<<<BlockGrid(ws.size()), 256, 0, stream.cuda_stream()>>>(ws, f, args...);
    CHECK_CUDA(cudaGetLastError());
  }
  template <typename WORK_SOURCE_T, typename FUNC_T, typename... Args>
  void ForEachWithIndexBlockShared(const Stream& stream, const WORK_SOURCE_T& ws, FUNC_T f, Args... args) {
    if (ws.size() == 0) return;
    compat_detail::k_fe_block_shared<<<SharedGrid(ws.size()), 256, 0, stream.cuda_stream()>>>(ws, f, args...);
    CHECK_CUDA(cudaGetLastError());
  }
  template <typename WORK_SOURCE_T, typename FUNC_T, typename... Args>
  void ForEachWithIndexBlockDynamic(const Stream& stream, const WORK_SOURCE_T& ws, FUNC_T f, Args... args) {
    if (ws.size() == 0) return;
    const unsigned grid = BlockGrid(ws.size());
    compat_detail::k_set_ticket<<<1, 1, 0, stream.cuda_stream()>>>(Ticket(), (unsigned long long) grid);
    compat_detail::k_fe_block_dynamic<<<grid, 256, 0, stream.cuda_stream()>>>(ws, Ticket(), f, args...);
    CHECK_CUDA(cudaGetLastError());
  }

 private:
  unsigned WarpGrid(size_t items) const {   // 8 warps per CTA, 8 CTAs per SM
    return (unsigned) std::max<size_t>(1, std::min<size_t>((items + 7) / 8, (size_t) sms_ * 8));
  }
  unsigned BlockGrid(size_t items) const {
    return (unsigned) std::max<size_t>(1, std::min<size_t>(items, (size_t) sms_ * 8));
  }
  unsigned SharedGrid(size_t ctas_wanted) const {
    // the *Shared forms keep the reference's bound of 256 CTAs (launcher.h:47-53): apps size
    // per-CTA global scratch with that constant (lcc_opt.h:189-195)
    return (unsigned) std::max<size_t>(1, std::min<size_t>(ctas_wanted, 256));
  }
  unsigned long long* Ticket() {
    if (!ticket_) CHECK_CUDA(cudaMalloc(&ticket_, sizeof(unsigned long long)));
    return ticket_;
  }
  unsigned long long* ticket_ = nullptr;

 public:
  // ForEachOutgoingEdge / ForEachIncomingEdge (parallel_engine.h:987-1182)
  template <typename FRAG_T, typename WORK_SOURCE_T, typename ASSIGN_OP, typename EDGE_OP>
  void ForEachOutgoingEdge(const Stream& stream, const FRAG_T& dev_frag, const WORK_SOURCE_T& ws,
                           ASSIGN_OP assign_op, EDGE_OP op, LoadBalancing lb) {
    compat_detail::MetaOp<FRAG_T, ASSIGN_OP, EDGE_OP> o{assign_op, op};
    const gl_frag_view& v = dev_frag.view();
    Scan(stream, ws, ::gl::EdgeRange{v.oe_rp, v.oe_col, v.oe_w}, v.oe_num, o, lb);
  }
  template <typename FRAG_T, typename WORK_SOURCE_T, typename EDGE_OP>
  void ForEachOutgoingEdge(const Stream& stream, const FRAG_T& dev_frag, const WORK_SOURCE_T& ws,
                           EDGE_OP op, LoadBalancing lb) {
    compat_detail::PlainOp<FRAG_T, EDGE_OP> o{op};
    const gl_frag_view& v = dev_frag.view();
    Scan(stream, ws, ::gl::EdgeRange{v.oe_rp, v.oe_col, v.oe_w}, v.oe_num, o, lb);
  }
  template <typename FRAG_T, typename WORK_SOURCE_T, typename ASSIGN_OP, typename EDGE_OP>
  void ForEachIncomingEdge(const Stream& stream, const FRAG_T& dev_frag, const WORK_SOURCE_T& ws,
                           ASSIGN_OP assign_op, EDGE_OP op, LoadBalancing lb) {
    compat_detail::MetaOp<FRAG_T, ASSIGN_OP, EDGE_OP> o{assign_op, op};
    const gl_frag_view& v = dev_frag.view();
    Scan(stream, ws, ::gl::EdgeRange{v.ie_rp, v.ie_col, v.ie_w}, v.ie_num, o, lb);
  }
  template <typename FRAG_T, typename WORK_SOURCE_T, typename EDGE_OP>
  void ForEachIncomingEdge(const Stream& stream, const FRAG_T& dev_frag, const WORK_SOURCE_T& ws,
                           EDGE_OP op, LoadBalancing lb) {
    compat_detail::PlainOp<FRAG_T, EDGE_OP> o{op};
    const gl_frag_view& v = dev_frag.view();
    Scan(stream, ws, ::gl::EdgeRange{v.ie_rp, v.ie_col, v.ie_w}, v.ie_num, o, lb);
  }

  // dispatch onto the engine's kernel skeletons (public: kernels instantiated
  // with extended-lambda types may not be launched from private members)
  template <typename WORK_SOURCE_T, typename OP>
  void Scan(const Stream& stream, const WORK_SOURCE_T& ws, ::gl::EdgeRange er, uint64_t entries, const OP& op,
            LoadBalancing lb) {
    using namespace ::gl;  // NOLINT
    using src_t = typename compat_detail::SrcOf<WORK_SOURCE_T>::type;
    const uint32_t n = (uint32_t) ws.size();
    if (n == 0) return;
    cudaStream_t s = stream.cuda_stream();
    src_t src = compat_detail::SrcOf<WORK_SOURCE_T>::make(ws);
    {
      cudaError_t pre = cudaGetLastError();
      if (pre != cudaSuccess) LOG(FATAL) << "error pending before the edge scan: " << cudaGetErrorString(pre);
    }
    Ensure(n, entries);
    CHECK_CUDA(cudaMemsetAsync(ctrl_, 0, sizeof(ScanCtrl), s));
    const int grid_v = std::max(1, std::min<int>(sms_ * 8, (int) ((n + kTB - 1) / kTB)));
    const int grid_t = std::max(1, std::min<int>(sms_ * 8, (int) ((n + kTileV - 1) / kTileV)));
    switch (lb) {
      case LoadBalancing::kNone:
        k_queue_scan_none<OP, src_t><<<grid_v, kTB, 0, s>>>(src, n, er, op, ctrl_);
        break;
      case LoadBalancing::kWarp:
        k_queue_scan_warp<OP, src_t><<<grid_v, kTB, 0, s>>>(src, n, er, op, ctrl_);
        break;
      case LoadBalancing::kCM:
      case LoadBalancing::kCMOld:
        // CTA-cooperative mapping; rows longer than 8 tiles' worth of threads would serialise one CTA
        // (measured on R-MAT: 2x slower than the reference's own cm), so they are cut into work items
        k_queue_scan_cta<OP, src_t><<<grid_t, kTB, 0, s>>>(src, n, er, op, ctrl_, hubs_, hub_cap_, 8 * kHubDeg);
        HubScan(s, er, op);
        break;
      case LoadBalancing::kCTA:
        k_queue_scan_cta<OP, src_t><<<grid_t, kTB, 0, s>>>(src, n, er, op, ctrl_, hubs_, hub_cap_, kHubDeg);
        HubScan(s, er, op);
        break;
      case LoadBalancing::kStrict:
        k_queue_degrees<src_t><<<(n + 1 + 255) / 256, 256, 0, s>>>(src, n, er.rp, deg_);
        CHECK_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp_, scan_bytes_, deg_, pfx_, (int) (n + 1), s));
        k_queue_scan_strict<OP, src_t><<<sms_ * 8, kTB, 0, s>>>(src, n, pfx_, er, op, ctrl_);
        break;
    }
    CHECK_CUDA(cudaGetLastError());
  }

  // long rows: 1024-entry work items over all SMs, staged by the TMA engine where the op allows it
  template <typename OP>
  void HubScan(cudaStream_t s, ::gl::EdgeRange er, const OP& op) {
    using namespace ::gl;  // NOLINT
    if constexpr (op_tma_ok<OP>::value) k_hub_scan_tma<OP><<<sms_ * 8, kTB, 0, s>>>(er, op, ctrl_, hubs_, hub_cap_);
    else k_hub_scan<OP><<<sms_ * 8, kTB, 0, s>>>(er, op, ctrl_, hubs_, hub_cap_);
  }

  void Ensure(uint32_t n, uint64_t entries) {
    using namespace ::gl;  // NOLINT
    if (!ctrl_) CHECK_CUDA(cudaMalloc(&ctrl_, sizeof(ScanCtrl)));
    // every long row has > kHubDeg entries => #items <= M/kHubChunk + M/kHubDeg
    uint32_t need = (uint32_t) std::min<uint64_t>(entries / kHubChunk + entries / kHubDeg + 1024, 0x7FFFFFFFull);
    if (need > hub_cap_) {
      if (hubs_) cudaFree(hubs_);
      CHECK_CUDA(cudaMalloc(&hubs_, sizeof(HubItem) * (size_t) need));
      hub_cap_ = need;
    }
    if (n + 1 > pfx_cap_) {
      if (deg_) cudaFree(deg_);
      if (pfx_) cudaFree(pfx_);
      if (scan_tmp_) cudaFree(scan_tmp_);
      CHECK_CUDA(cudaMalloc(&deg_, sizeof(uint64_t) * ((size_t) n + 1)));
      CHECK_CUDA(cudaMalloc(&pfx_, sizeof(uint64_t) * ((size_t) n + 1)));
      scan_bytes_ = 0;
      CHECK_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes_, deg_, pfx_, (int) (n + 1)));
      CHECK_CUDA(cudaMalloc(&scan_tmp_, std::max<size_t>(scan_bytes_, 16)));
      pfx_cap_ = n + 1;
    }
  }


 private:
  int sms_ = 148;
  ::gl::ScanCtrl* ctrl_ = nullptr;
  ::gl::HubItem* hubs_ = nullptr;
  uint32_t hub_cap_ = 0;
  uint64_t* deg_ = nullptr;
  uint64_t* pfx_ = nullptr;
  uint32_t pfx_cap_ = 0;
  void* scan_tmp_ = nullptr;
  size_t scan_bytes_ = 0;
};

// ----------------------------------------------------------- app base / worker --
// grape/cuda/app/gpu_app_base.h:39-92, grape/cuda/worker/gpu_worker.h:44-107
template <typename APP_T>
class GPUWorker;

template <typename FRAG_T, typename CONTEXT_T>
class GPUAppBase {
 public:
  static constexpr bool need_split_edges = false;
  static constexpr bool need_build_device_vm = false;
  static constexpr grape::MessageStrategy message_strategy = grape::MessageStrategy::kSyncOnOuterVertex;
  static constexpr grape::LoadStrategy load_strategy = grape::LoadStrategy::kOnlyOut;
  using message_manager_t = GPUMessageManager;
  GPUAppBase() = default;
  virtual ~GPUAppBase() = default;
  virtual void PEval(const FRAG_T& graph, CONTEXT_T& context, message_manager_t& messages) = 0;
  virtual void IncEval(const FRAG_T& graph, CONTEXT_T& context, message_manager_t& messages) = 0;
};

#define INSTALL_GPU_WORKER(APP_T, CONTEXT_T, FRAG_T)              \
 public:                                                          \
  using fragment_t = FRAG_T;                                      \
  using context_t = CONTEXT_T;                                    \
  using worker_t = grape::cuda::GPUWorker<APP_T>;                 \
  using message_manager_t = grape::cuda::GPUMessageManager;       \
  using dev_message_manager_t = grape::cuda::dev::MessageManager; \
  virtual ~APP_T() {}                                             \
  static std::shared_ptr<worker_t> CreateWorker(                  \
      std::shared_ptr<APP_T> app, std::shared_ptr<FRAG_T> frag) { \
    return std::shared_ptr<worker_t>(new worker_t(app, frag));    \
  }

template <typename APP_T>
class GPUWorker {
 public:
  using fragment_t = typename APP_T::fragment_t;
  using context_t = typename APP_T::context_t;
  using message_manager_t = GPUMessageManager;

  GPUWorker(std::shared_ptr<APP_T> app, std::shared_ptr<fragment_t> graph)
      : app_(std::move(app)), context_(std::make_shared<context_t>(*graph)), messages_() {}

  template <class... Args>
  void Init(const grape::CommSpec& comm_spec, Args&&... args) {
    auto& graph = const_cast<fragment_t&>(context_->fragment());
    PrepareConf conf;
    conf.message_strategy = APP_T::message_strategy;
    conf.need_split_edges = APP_T::need_split_edges;
    conf.need_split_edges_by_fragment = false;
    conf.need_mirror_info = false;
    conf.need_build_device_vm = APP_T::need_build_device_vm;
    graph.PrepareToRunApp(comm_spec, conf, grape::DefaultParallelEngineSpec());
    comm_spec_ = comm_spec;
    messages_.Init(comm_spec);
    InitCommunicator(app_, comm_spec.comm(), messages_.nccl_comm());
    context_->Init(messages_, std::forward<Args>(args)...);
  }
  void Finalize() {}
  // PEval once, IncEval until every fragment is idle
  void Query() {
    auto& graph = context_->fragment();
    messages_.Start();
    messages_.StartARound();
    app_->PEval(graph, *context_, messages_);
    messages_.FinishARound();
    int step = 1;
    while (!messages_.ToTerminate()) {
      double t0 = grape::GetCurrentTime();
      messages_.StartARound();
      app_->IncEval(graph, *context_, messages_);
      messages_.FinishARound();
      VLOG(1) << "[Coordinator]: Finished IncEval - " << step << " Time: "
              << (grape::GetCurrentTime() - t0) * 1000 << " ms";
      ++step;
    }
    supersteps_ = step;
    messages_.Finalize();
  }
  int supersteps() const { return supersteps_; }
  std::shared_ptr<context_t> GetContext() { return context_; }
  void Output(std::ostream& os) { context_->Output(os); }

 private:
  std::shared_ptr<APP_T> app_;
  std::shared_ptr<context_t> context_;
  message_manager_t messages_;
  grape::CommSpec comm_spec_;
  int supersteps_ = 0;
};

}  // namespace cuda
}  // namespace grape

#endif  // __CUDACC__
#endif  // GRAPE_CUDA_B200_COMPAT_H_
