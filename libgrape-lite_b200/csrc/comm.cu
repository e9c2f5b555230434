// comm.cu — landing-area allocation / CUDA-IPC mapping and the round protocol.
#include "comm.h"

#include <algorithm>
#include <cstdlib>

namespace gl {

namespace {
__global__ void k_publish_counts(uint32_t fnum, uint32_t fid,
                                 uint32_t* const* peer_count,
                                 uint32_t* send_count, uint32_t* h_out) {
  // one thread per destination: store my item count into the owner's header
  uint32_t d = threadIdx.x;
  uint32_t total = 0;
  if (d < fnum) {
    uint32_t c = send_count[d];
    if (d != fid) {
      *peer_count[d] = c;  // NVLink peer store
      total = c;
    }
    h_out[d] = c;
    send_count[d] = 0;  // ready for the next round
  }
  __threadfence_system();
}
// One block, one thread per peer: publish my contribution to every peer, wait
// for every peer's contribution of the same sequence number, reduce.
// With peer_count != nullptr it first publishes this round's item counts
// (FinishARound) and adds their total to i1, so that a round closes with ONE
// kernel and ONE stream synchronisation.
__global__ void k_peer_allreduce(uint32_t fnum, uint32_t fid, PeerSlot* const* my_slot_at_peer,
                                 const PeerSlot* local_slots, unsigned long long tag,
                                 long long i0, long long i1, double d0, int op,
                                 uint32_t* const* peer_count, uint32_t* send_count,
                                 uint32_t* h_counts, PeerSlot* h_out) {
  __shared__ long long s_i0[GL_MAX_FNUM], s_i1[GL_MAX_FNUM];
  __shared__ double s_d0[GL_MAX_FNUM];
  __shared__ unsigned int s_total;
  const uint32_t p = threadIdx.x;
  if (p == 0) s_total = 0;
  __syncthreads();
  if (peer_count && p < fnum) {
    uint32_t c = send_count[p];
    if (p != fid) {
      *peer_count[p] = c;  // NVLink peer store
      atomicAdd(&s_total, c);
    }
    h_counts[p] = c;
    send_count[p] = 0;
  }
  __syncthreads();
  i1 += (long long) s_total;
  if (p < fnum) {
    PeerSlot* dst = my_slot_at_peer[p];        // peer p's header, slot [me] (p == me: local)
    dst->i0 = i0;
    dst->i1 = i1;
    dst->d0 = d0;
    __threadfence_system();
    *(volatile unsigned long long*) &dst->tag = tag;
    // wait for peer p's contribution in MY header, slot [p]
    const volatile PeerSlot* src = local_slots + p;
    while (src->tag != tag) __nanosleep(64);
    __threadfence_system();
    s_i0[p] = src->i0;
    s_i1[p] = src->i1;
    s_d0[p] = src->d0;
  }
  __syncthreads();
  if (p == 0) {
    long long a = s_i0[0], b = s_i1[0];
    double c = s_d0[0];
    for (uint32_t q = 1; q < fnum; ++q) {
      if (op == 0) { a += s_i0[q]; b += s_i1[q]; c += s_d0[q]; }
      else if (op == 1) { a = a < s_i0[q] ? a : s_i0[q]; b = b < s_i1[q] ? b : s_i1[q]; c = c < s_d0[q] ? c : s_d0[q]; }
      else { a = a > s_i0[q] ? a : s_i0[q]; b = b > s_i1[q] ? b : s_i1[q]; c = c > s_d0[q] ? c : s_d0[q]; }
    }
    h_out->i0 = a;
    h_out->i1 = b;
    h_out->d0 = c;
    h_out->tag = tag;
  }
  (void) fid;
}
}  // namespace

int MessageManager::PeerAllReduce(cudaStream_t s, long long* i0, long long* i1, double* d0, int op) {
  const unsigned long long tag = ++comm->seq_base;
  const int par = (int) (tag & 1);
  const PeerSlot* local = (const PeerSlot*) (comm->local_base + GL_COMM_SLOT_OFF) + (size_t) par * GL_MAX_FNUM;
  k_peer_allreduce<<<1, GL_MAX_FNUM, 0, s>>>(fnum, fid, d_peer_slot[par], local, tag, *i0, *i1, *d0, op,
                                             nullptr, nullptr, nullptr, h_result);
  GL_COUNT_LAUNCH();
  GL_CUDA(cudaGetLastError());
  GL_CUDA(cudaStreamSynchronize(s));
  *i0 = h_result->i0;
  *i1 = h_result->i1;
  *d0 = h_result->d0;
  return GL_OK;
}

int MessageManager::Init(gl_comm* c, const gl_frag_view& fv, uint32_t item_bytes_) {
  comm = c;
  fid = fv.fid;
  fnum = fv.fnum;
  fid_offset = fv.fid_offset;
  id_mask = fv.id_mask;
  item_bytes = item_bytes_;
  if (fnum == 1) return GL_OK;
  if (!c || !c->opened || c->fnum != fnum || c->fid != fid) {
    set_error("fragment has fnum=%u but no opened communicator was supplied", fnum);
    return GL_ERR_COMM;
  }
  GL_CUDA(cudaMalloc(&d_send_count, sizeof(uint32_t) * fnum));
  GL_CUDA(cudaMemset(d_send_count, 0, sizeof(uint32_t) * fnum));
  GL_CUDA(cudaMallocHost(&h_send_count, sizeof(uint32_t) * (fnum + 1)));
  GL_CUDA(cudaMallocHost(&h_result, sizeof(PeerSlot)));
  for (int par = 0; par < 2; ++par) {
    std::vector<char*> send(fnum);
    std::vector<const char*> recv(fnum);
    std::vector<uint32_t*> pc(fnum);
    for (uint32_t p = 0; p < fnum; ++p) {
      // my slot at peer p: (parity, src = me)
      send[p] = c->peer_base[p] + GL_COMM_HEADER +
                ((size_t) par * fnum + fid) * c->landing_bytes;
      pc[p] = (uint32_t*) (c->peer_base[p]) + (size_t) par * GL_MAX_FNUM + fid;
      // peer p's slot in my area
      recv[p] = c->local_base + GL_COMM_HEADER +
                ((size_t) par * fnum + p) * c->landing_bytes;
    }
    std::vector<PeerSlot*> ps(fnum);
    for (uint32_t p = 0; p < fnum; ++p)
      ps[p] = (PeerSlot*) (c->peer_base[p] + GL_COMM_SLOT_OFF) + (size_t) par * GL_MAX_FNUM + fid;
    GL_CUDA(cudaMalloc(&d_peer_slot[par], sizeof(PeerSlot*) * fnum));
    GL_CUDA(cudaMemcpy(d_peer_slot[par], ps.data(), sizeof(PeerSlot*) * fnum, cudaMemcpyHostToDevice));
    GL_CUDA(cudaMalloc(&d_send_slot[par], sizeof(char*) * fnum));
    GL_CUDA(cudaMalloc(&d_recv_slot[par], sizeof(char*) * fnum));
    GL_CUDA(cudaMalloc(&d_peer_count[par], sizeof(uint32_t*) * fnum));
    GL_CUDA(cudaMemcpy(d_send_slot[par], send.data(), sizeof(char*) * fnum, cudaMemcpyHostToDevice));
    GL_CUDA(cudaMemcpy(d_recv_slot[par], recv.data(), sizeof(char*) * fnum, cudaMemcpyHostToDevice));
    GL_CUDA(cudaMemcpy(d_peer_count[par], pc.data(), sizeof(uint32_t*) * fnum, cudaMemcpyHostToDevice));
  }
  return GL_OK;
}

void MessageManager::Destroy() {
  for (int par = 0; par < 2; ++par) {
    if (d_send_slot[par]) cudaFree(d_send_slot[par]);
    if (d_recv_slot[par]) cudaFree((void*) d_recv_slot[par]);
    if (d_peer_count[par]) cudaFree(d_peer_count[par]);
    if (d_peer_slot[par]) cudaFree(d_peer_slot[par]);
    d_peer_slot[par] = nullptr;
    d_send_slot[par] = nullptr;
    d_recv_slot[par] = nullptr;
    d_peer_count[par] = nullptr;
  }
  if (d_send_count) cudaFree(d_send_count);
  if (h_send_count) cudaFreeHost(h_send_count);
  if (h_result) cudaFreeHost(h_result);
  h_result = nullptr;
  d_send_count = nullptr;
  h_send_count = nullptr;
}

MsgView MessageManager::view() const {
  MsgView v;
  memset(&v, 0, sizeof(v));
  v.fid = fid;
  v.fnum = fnum;
  v.fid_offset = fid_offset;
  v.id_mask = id_mask;
  v.item_bytes = item_bytes;
  if (fnum > 1) {
    int par = round & 1, prev = (round + 1) & 1;
    v.capacity = (uint32_t) std::min<size_t>(comm->landing_bytes / item_bytes, 0xFFFFFFFFu);
    v.send_slot = d_send_slot[par];
    v.send_count = d_send_count;
    v.recv_slot = d_recv_slot[prev];
    v.recv_count = (const uint32_t*) comm->local_base + (size_t) prev * GL_MAX_FNUM;
  }
  return v;
}

int MessageManager::StartARound(cudaStream_t s) {
  stream_for_collectives = s;
  force_continue = false;
  return GL_OK;
}

// publish counts -> sync -> all-reduce (barrier + termination vote)
int MessageManager::FinishARound(cudaStream_t s) {
  int64_t vote[2] = {force_continue ? 1 : 0, 0};
  if (fnum > 1) {
    int par = round & 1;
    if (use_peer_barrier) {
      // publish counts + barrier + termination vote: one kernel, one sync
      const unsigned long long tag = ++comm->seq_base;
      const int bp = (int) (tag & 1);
      const PeerSlot* local = (const PeerSlot*) (comm->local_base + GL_COMM_SLOT_OFF) + (size_t) bp * GL_MAX_FNUM;
      k_peer_allreduce<<<1, GL_MAX_FNUM, 0, s>>>(fnum, fid, d_peer_slot[bp], local, tag, (long long) vote[0], 0ll,
                                                 0.0, 0, d_peer_count[par], d_send_count, h_send_count, h_result);
      GL_COUNT_LAUNCH();
      GL_CUDA(cudaGetLastError());
      GL_CUDA(cudaStreamSynchronize(s));
      vote[0] = h_result->i0;
      vote[1] = h_result->i1;
    } else {
      k_publish_counts<<<1, GL_MAX_FNUM, 0, s>>>(fnum, fid, d_peer_count[par], d_send_count, h_send_count);
      GL_COUNT_LAUNCH();
      GL_CUDA(cudaGetLastError());
      GL_CUDA(cudaStreamSynchronize(s));
    }
    uint64_t sent = 0;
    for (uint32_t p = 0; p < fnum; ++p)
      if (p != fid) {
        if (h_send_count[p] > comm->landing_bytes / item_bytes) {
          set_error("landing slot overflow: %u items to fragment %u", h_send_count[p], p);
          return GL_ERR_COMM;
        }
        sent += h_send_count[p];
      }
    bytes_sent += sent * item_bytes;
    if (!use_peer_barrier) {
      vote[1] = (int64_t) sent;
      GL_TRY(AllReduceI64(vote, 2, 0));
    }
  } else {
    GL_CUDA(cudaStreamSynchronize(s));
  }
  terminate = (vote[0] == 0 && vote[1] == 0);
  ++round;
  return GL_OK;
}

int MessageManager::AllReduceI64(int64_t* v, int n, int op) {
  if (fnum == 1) return GL_OK;
  if (!comm || !comm->allreduce) {
    set_error("communicator has no allreduce callback");
    return GL_ERR_COMM;
  }
  int st = comm->allreduce(comm->user, v, n, 0, op);
  if (st != 0) {
    set_error("allreduce callback failed (%d)", st);
    return GL_ERR_COMM;
  }
  return GL_OK;
}
int MessageManager::AllReduceF64(double* v, int n, int op) {
  if (fnum == 1) return GL_OK;
  if (use_peer_barrier && n == 1 && comm && comm->opened) {
    long long a = 0, b = 0;
    cudaStream_t s0 = stream_for_collectives;
    return PeerAllReduce(s0, &a, &b, v, op);
  }
  if (!comm || !comm->allreduce) {
    set_error("communicator has no allreduce callback");
    return GL_ERR_COMM;
  }
  int st = comm->allreduce(comm->user, v, n, 1, op);
  if (st != 0) {
    set_error("allreduce callback failed (%d)", st);
    return GL_ERR_COMM;
  }
  return GL_OK;
}

}  // namespace gl

using namespace gl;

extern "C" {

int gl_comm_create(gl_comm_t** out, const gl_comm_desc* d) {
  GL_ARG(out && d, "null argument");
  GL_ARG(d->fnum >= 1 && d->fnum <= GL_MAX_FNUM && d->fid < d->fnum, "bad fid/fnum");
  GL_ARG(d->fnum == 1 || d->allreduce, "allreduce callback required when fnum > 1");
  DeviceInfo* di;
  GL_TRY(device_info(&di));
  gl_comm* c = new gl_comm;
  c->fid = d->fid;
  c->fnum = d->fnum;
  c->allreduce = d->allreduce;
  c->user = d->user;
  c->landing_bytes = (d->landing_bytes + 255) & ~(size_t) 255;
  cudaError_t e = cudaMalloc(&c->local_base, c->total_bytes());
  if (e != cudaSuccess) {
    set_error("landing area (%zu bytes): %s", c->total_bytes(), cudaGetErrorString(e));
    delete c;
    return GL_ERR_NOMEM;
  }
  cudaMemset(c->local_base, 0, GL_COMM_HEADER);
  c->peer_base.assign(c->fnum, nullptr);
  c->peer_base[c->fid] = c->local_base;
  if (c->fnum == 1) c->opened = true;
  *out = c;
  return GL_OK;
}

int gl_comm_export(gl_comm_t* c, void* handles_out, size_t bytes) {
  GL_ARG(c && handles_out && bytes >= GL_IPC_HANDLE_BYTES, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) <= GL_IPC_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h;
  GL_CUDA(cudaIpcGetMemHandle(&h, c->local_base));
  memset(handles_out, 0, GL_IPC_HANDLE_BYTES);
  memcpy(handles_out, &h, sizeof(h));
  return GL_OK;
}

int gl_comm_open(gl_comm_t* c, const void* all, size_t bytes) {
  GL_ARG(c && all && bytes >= (size_t) c->fnum * GL_IPC_HANDLE_BYTES, "bad argument");
  for (uint32_t p = 0; p < c->fnum; ++p) {
    if (p == c->fid) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*) all + (size_t) p * GL_IPC_HANDLE_BYTES, sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("cudaIpcOpenMemHandle(peer %u): %s", p, cudaGetErrorString(e));
      return GL_ERR_COMM;
    }
    c->peer_base[p] = (char*) ptr;
  }
  c->opened = true;
  return GL_OK;
}

void gl_comm_destroy(gl_comm_t* c) {
  if (!c) return;
  for (uint32_t p = 0; p < c->fnum; ++p)
    if (p != c->fid && c->peer_base[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
  if (c->local_base) cudaFree(c->local_base);
  delete c;
}

}  // extern "C"
