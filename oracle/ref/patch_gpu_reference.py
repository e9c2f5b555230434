#!/usr/bin/env python
"""Build recipe of the "existing kernel" baseline: the reference's OWN
GPU code (grape/cuda/**) compiled for sm_100a with the CUDA 12.9 toolchain of this image.

TEST / MEASUREMENT INFRASTRUCTURE.  The reference tree is read-only and does not compile as it is
with this toolchain; this script writes a PATCHED COPY of grape/cuda/** into oracle/_ref/gpu_patched/
(git-ignored build output, never committed) and applies four small, documented fixes that do not
touch any algorithm:

 1. fragment/host_fragment.h: PrepareToRunApp forwards an undeclared `pe_spec`; the parameter is
    called `engine_spec` (a plain typo in the reference).
 2. utils/cuda_utils.h: with thrust > 1.17 `pinned_vector<T>::data()` is a fancy pointer, but the
    sources hand data() to kernels and to ArrayView(T*, n); a std-style cudaMallocHost allocator keeps
    it a raw pointer.
 3. utils/cuda_utils.h: PrefixSumKernel64 instantiates cub::DispatchScan with a template list that
    no longer exists in CUDA 12.9's CUB; cub::DeviceScan::ExclusiveScan with a size_t initial value
    is the same scan with the same size_t accumulator.
 4. vertex_map/device_vertex_map.h: a thrust device_reference (d_o2l_[fid]) is passed through a
    kernel's variadic arguments; it is converted to the raw pointer on the host first.

usage: patch_gpu_reference.py <reference root> <output dir>"""
import os
import shutil
import sys


def sub(path, old, new, count=1):
    s = open(path).read()
    if old not in s:
        raise SystemExit("patch_gpu_reference: pattern not found in %s:\n%s" % (path, old[:120]))
    open(path, "w").write(s.replace(old, new, count))


def main():
    ref, out = sys.argv[1], sys.argv[2]
    dst = os.path.join(out, "grape", "cuda")
    shutil.rmtree(out, ignore_errors=True)
    shutil.copytree(os.path.join(ref, "grape", "cuda"), dst)
    for root, _, files in os.walk(dst):
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    sub(os.path.join(dst, "fragment", "host_fragment.h"),
        "base_t::PrepareToRunApp(comm_spec, conf, pe_spec);",
        "base_t::PrepareToRunApp(comm_spec, conf, engine_spec);  // [B200 build patch 1]")
    cu = os.path.join(dst, "utils", "cuda_utils.h")
    s = open(cu).read()
    a = s.index("#if THRUST_VERSION > 101700\nusing memory_resource")
    b = s.index("#endif", a) + len("#endif")
    s = s[:a] + """// [B200 build patch 2] raw-pointer data() for page-locked host vectors
template <typename T>
struct ref_pinned_allocator {
  using value_type = T;
  ref_pinned_allocator() = default;
  template <typename U>
  ref_pinned_allocator(const ref_pinned_allocator<U>&) {}
  T* allocate(size_t n) {
    void* p = nullptr;
    CHECK_CUDA(cudaMallocHost(&p, (n ? n : 1) * sizeof(T)));
    return static_cast<T*>(p);
  }
  void deallocate(T* p, size_t) { cudaFreeHost(p); }
  template <typename U>
  bool operator==(const ref_pinned_allocator<U>&) const { return true; }
  template <typename U>
  bool operator!=(const ref_pinned_allocator<U>&) const { return false; }
};
template <typename T>
using pinned_vector = thrust::host_vector<T, ref_pinned_allocator<T>>;""" + s[b:]
    a = s.index("#if CUB_VERSION > 200000\n  struct InitValue {")
    b = s.index("#endif", a) + len("#endif")
    s = s[:a] + """  // [B200 build patch 3] public CUB API, size_t accumulator
  (void) debug_synchronous;
  return cub::DeviceScan::ExclusiveScan(d_temp_storage, temp_storage_bytes, d_in, d_out, cub::Sum(), (size_t) 0,
                                        num_items, stream);""" + s[b:]
    open(cu, "w").write(s)
    sub(os.path.join(dst, "vertex_map", "device_vertex_map.h"),
        "oids.data(), ivnum, d_o2l_[fid]);",
        "oids.data(), ivnum,\n          static_cast<CUDASTL::HashMap<OID_T, VID_T>*>(d_o2l_[fid]));  // [B200 build patch 4]")
    print("patched copy of grape/cuda written to", dst)


if __name__ == "__main__":
    main()
