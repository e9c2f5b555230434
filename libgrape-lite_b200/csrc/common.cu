// common.cu — error reporting, device probing and the small device-memory
// helpers of the C ABI.
#include <cstdlib>

#include "common.cuh"

namespace gl {

static thread_local char g_err[1024] = "";
thread_local uint64_t g_kernel_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int device_info(DeviceInfo** out) {
  static thread_local DeviceInfo info;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no usable CUDA device: %s (this library has no CPU fallback)",
              cudaGetErrorString(e));
    return GL_ERR_CUDA;
  }
  if (info.device != dev) {
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) {
      set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
      return GL_ERR_CUDA;
    }
    info.device = dev;
    info.sm_count = p.multiProcessorCount;
    info.cc = p.major * 10 + p.minor;
    info.l2_bytes = (size_t) p.l2CacheSize;
    info.hbm_bytes = p.totalGlobalMem;
  }
  *out = &info;
  return GL_OK;
}

namespace {
struct L2Cfg {
  int device = -1;
  size_t persist_max = 0, window_max = 0;
  bool enabled = true;
  bool carved = false;   // the persisting set-aside is currently taken out of the L2
};
L2Cfg* l2cfg() {
  static thread_local L2Cfg c;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  if (c.device != dev) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return nullptr;
    c.device = dev;
    c.persist_max = (size_t) p.persistingL2CacheMaxSize;
    c.window_max = (size_t) p.accessPolicyMaxWindowSize;
    const char* e = getenv("GL_L2_PERSIST");
    c.enabled = !(e && atoi(e) == 0) && c.persist_max > 0 && c.window_max > 0;
    c.carved = false;
  }
  return &c;
}
}  // namespace

int l2_persist_window(cudaStream_t s, const void* ptr, size_t bytes) {
  L2Cfg* c = l2cfg();
  if (!c || !c->enabled || !ptr || !bytes) return GL_OK;
  // The set-aside is carved out of the L2 for EVERY kernel of the device, also for apps that install
  // no window (measured: PageRank's atomic push drops from 7.0 to 9.6 ms per round while it is taken):
  // it is taken only while a window is installed and given back by l2_persist_clear.
  if (!c->carved) {
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, c->persist_max) != cudaSuccess) {
      cudaGetLastError();
      c->enabled = false;
      return GL_OK;
    }
    c->carved = true;
  }
  cudaStreamAttrValue a;
  memset(&a, 0, sizeof(a));
  a.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
  a.accessPolicyWindow.num_bytes = bytes < c->window_max ? bytes : c->window_max;
  // the fraction of the window that may persist: what the set-aside can hold
  double ratio = (double) c->persist_max * 0.95 / (double) a.accessPolicyWindow.num_bytes;
  a.accessPolicyWindow.hitRatio = (float) (ratio > 1.0 ? 1.0 : ratio);
  a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &a) != cudaSuccess) cudaGetLastError();
  return GL_OK;
}

int l2_persist_clear(cudaStream_t s) {
  L2Cfg* c = l2cfg();
  if (!c || !c->enabled) return GL_OK;
  cudaStreamAttrValue a;
  memset(&a, 0, sizeof(a));
  a.accessPolicyWindow.num_bytes = 0;
  if (cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &a) != cudaSuccess) cudaGetLastError();
  cudaCtxResetPersistingL2Cache();
  if (c->carved) {
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0) != cudaSuccess) cudaGetLastError();
    c->carved = false;
  }
  return GL_OK;
}

}  // namespace gl

using namespace gl;

extern "C" {

const char* gl_last_error(void) { return last_error(); }
int gl_abi_version(void) { return GL_ABI_VERSION; }

int gl_device_info(int* sm_count, int* cc, size_t* l2_bytes, size_t* hbm_bytes) {
  DeviceInfo* di;
  GL_TRY(device_info(&di));
  if (sm_count) *sm_count = di->sm_count;
  if (cc) *cc = di->cc;
  if (l2_bytes) *l2_bytes = di->l2_bytes;
  if (hbm_bytes) *hbm_bytes = di->hbm_bytes;
  return GL_OK;
}

int gl_dev_alloc(void** out, size_t bytes) {
  GL_ARG(out, "null argument");
  GL_CUDA(cudaMalloc(out, bytes ? bytes : 16));
  return GL_OK;
}
int gl_dev_free(void* p) {
  if (p) GL_CUDA(cudaFree(p));
  return GL_OK;
}
int gl_dev_memset(void* p, int byte, size_t bytes) {
  GL_CUDA(cudaMemset(p, byte, bytes));
  return GL_OK;
}
int gl_dev_h2d(void* dst, const void* src, size_t bytes) {
  GL_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return GL_OK;
}
int gl_dev_d2h(void* dst, const void* src, size_t bytes) {
  GL_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return GL_OK;
}
int gl_dev_sync(void) {
  GL_CUDA(cudaDeviceSynchronize());
  return GL_OK;
}

uint64_t gl_kernel_launch_count(void) { return g_kernel_launches; }

int gl_host_alloc_pinned(void** out, size_t bytes) {
  GL_ARG(out, "null argument");
  GL_CUDA(cudaMallocHost(out, bytes ? bytes : 16));
  return GL_OK;
}
int gl_host_free_pinned(void* p) {
  if (p) GL_CUDA(cudaFreeHost(p));
  return GL_OK;
}

}  // extern "C"
