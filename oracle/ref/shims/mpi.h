// mpi.h — functional SINGLE-RANK MPI shim (test infrastructure).
//
// The container has no MPI.  This header lets the UNMODIFIED reference CPU
// sources (grape/**, examples/analytical_apps/**) compile and run as one rank:
// rank 0 / size 1, collectives are memcpy (honouring MPI_IN_PLACE), and
// point-to-point traffic goes through a thread-safe per-communicator
// self-mailbox, because ParallelMessageManager runs a receiver thread blocked
// in MPI_Probe that is stopped by a zero-length self send
// (grape/parallel/parallel_message_manager.h:436-459,521-525) and the loaders
// shuffle edges to self (grape/fragment/basic_fragment_loader.h:122-195).
#ifndef ORACLE_REF_SHIM_MPI_H_
#define ORACLE_REF_SHIM_MPI_H_

#include <unistd.h>  // real mpi.h pulls this in; grape/util.h:65 and local_io_adaptor.cc rely on it

#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;  // = element size in bytes
typedef int MPI_Op;
struct MPI_Status {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
  long long _bytes;
};
struct shim_request;
typedef shim_request* MPI_Request;

#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_SUCCESS 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_THREAD_MULTIPLE 3
#define MPI_STATUS_IGNORE ((MPI_Status*) 0)
#define MPI_STATUSES_IGNORE ((MPI_Status*) 0)
#define MPI_IN_PLACE ((void*) 1)
#define MPI_REQUEST_NULL ((MPI_Request) 0)
#define MPI_UNDEFINED (-32766)

#define MPI_CHAR 1
#define MPI_BYTE 1
#define MPI_INT8_T 1
#define MPI_UINT8_T 1
#define MPI_INT 4
#define MPI_INT32_T 4
#define MPI_UINT32_T 4
#define MPI_FLOAT 4
#define MPI_UNSIGNED 4
#define MPI_DOUBLE 8
#define MPI_INT64_T 8
#define MPI_UINT64_T 8
#define MPI_LONG_LONG_INT 8
#define MPI_LONG_LONG 8
#define MPI_UNSIGNED_LONG 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_LONG 8

#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3

namespace shim_mpi {
struct Msg {
  int tag;
  std::vector<char> data;
};
struct Box {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Msg> q;
};
inline std::mutex& boxes_mu() {
  static std::mutex m;
  return m;
}
inline Box& box(MPI_Comm c) {
  static std::map<int, std::unique_ptr<Box>> boxes;
  std::lock_guard<std::mutex> g(boxes_mu());
  auto& p = boxes[c];
  if (!p) p.reset(new Box);
  return *p;
}
inline int next_comm() {
  static int n = 100;
  std::lock_guard<std::mutex> g(boxes_mu());
  return ++n;
}
inline bool match(const Msg& m, int tag) { return tag == MPI_ANY_TAG || m.tag == tag; }
}  // namespace shim_mpi

struct shim_request {
  bool is_recv = false, done = false, cancelled = false;
  MPI_Comm comm = 0;
  int tag = 0;
  void* buf = nullptr;
  long long cap = 0;
  MPI_Status st{};
};

inline int MPI_Init_thread(int*, char***, int, int* provided) {
  if (provided) *provided = MPI_THREAD_MULTIPLE;
  return 0;
}
inline int MPI_Init(int*, char***) { return 0; }
inline int MPI_Finalize() { return 0; }
inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return 0; }
inline int MPI_Comm_size(MPI_Comm, int* s) { *s = 1; return 0; }
inline int MPI_Comm_dup(MPI_Comm, MPI_Comm* n) { *n = shim_mpi::next_comm(); return 0; }
inline int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm* n) { *n = shim_mpi::next_comm(); return 0; }
inline int MPI_Comm_free(MPI_Comm* c) { if (c) *c = MPI_COMM_NULL; return 0; }
inline int MPI_Barrier(MPI_Comm) { return 0; }
inline int MPI_Get_processor_name(char* name, int* len) {
  std::strcpy(name, "localhost");
  *len = 9;
  return 0;
}
inline int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return 0; }
inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
  if (s != MPI_IN_PLACE && s != r) std::memcpy(r, s, (size_t) n * t);
  return 0;
}
inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) {
  if (s != MPI_IN_PLACE && s != r && r) std::memcpy(r, s, (size_t) n * t);
  return 0;
}
inline int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int, MPI_Datatype, MPI_Comm) {
  if (s != MPI_IN_PLACE && s != r) std::memcpy(r, s, (size_t) sn * st);
  return 0;
}
inline int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int*, const int* displs,
                          MPI_Datatype rt, MPI_Comm) {
  if (s != MPI_IN_PLACE) std::memcpy((char*) r + (size_t) (displs ? displs[0] : 0) * rt, s, (size_t) sn * st);
  return 0;
}
inline int MPI_Alltoall(const void* s, int sn, MPI_Datatype st, void* r, int, MPI_Datatype, MPI_Comm) {
  if (s != MPI_IN_PLACE && s != r) std::memcpy(r, s, (size_t) sn * st);
  return 0;
}
inline int MPI_Gather(const void* s, int sn, MPI_Datatype st, void* r, int, MPI_Datatype, int, MPI_Comm) {
  if (s != MPI_IN_PLACE && s != r) std::memcpy(r, s, (size_t) sn * st);
  return 0;
}

inline int MPI_Send(const void* buf, int n, MPI_Datatype t, int, int tag, MPI_Comm c) {
  shim_mpi::Box& b = shim_mpi::box(c);
  shim_mpi::Msg m;
  m.tag = tag;
  m.data.assign((const char*) buf, (const char*) buf + (size_t) n * t);
  {
    std::lock_guard<std::mutex> g(b.mu);
    b.q.push_back(std::move(m));
  }
  b.cv.notify_all();
  return 0;
}
inline int MPI_Isend(const void* buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* req) {
  MPI_Send(buf, n, t, dst, tag, c);
  shim_request* r = new shim_request;
  r->done = true;
  *req = r;
  return 0;
}
// blocking dequeue of the first message matching tag
inline void shim_take(MPI_Comm c, int tag, void* buf, long long cap, MPI_Status* st) {
  shim_mpi::Box& b = shim_mpi::box(c);
  std::unique_lock<std::mutex> g(b.mu);
  for (;;) {
    for (auto it = b.q.begin(); it != b.q.end(); ++it) {
      if (shim_mpi::match(*it, tag)) {
        long long n = (long long) it->data.size();
        if (n && buf) std::memcpy(buf, it->data.data(), (size_t) (n < cap ? n : cap));
        if (st) {
          st->MPI_SOURCE = 0;
          st->MPI_TAG = it->tag;
          st->MPI_ERROR = 0;
          st->_bytes = n;
        }
        b.q.erase(it);
        return;
      }
    }
    b.cv.wait(g);
  }
}
inline int MPI_Recv(void* buf, int n, MPI_Datatype t, int, int tag, MPI_Comm c, MPI_Status* st) {
  shim_take(c, tag, buf, (long long) n * t, st);
  return 0;
}
inline int MPI_Irecv(void* buf, int n, MPI_Datatype t, int, int tag, MPI_Comm c, MPI_Request* req) {
  shim_request* r = new shim_request;
  r->is_recv = true;
  r->comm = c;
  r->tag = tag;
  r->buf = buf;
  r->cap = (long long) n * t;
  *req = r;
  return 0;
}
inline int MPI_Probe(int, int tag, MPI_Comm c, MPI_Status* st) {
  shim_mpi::Box& b = shim_mpi::box(c);
  std::unique_lock<std::mutex> g(b.mu);
  for (;;) {
    for (auto& m : b.q) {
      if (shim_mpi::match(m, tag)) {
        if (st) {
          st->MPI_SOURCE = 0;
          st->MPI_TAG = m.tag;
          st->MPI_ERROR = 0;
          st->_bytes = (long long) m.data.size();
        }
        return 0;
      }
    }
    b.cv.wait(g);
  }
}
inline int MPI_Iprobe(int, int tag, MPI_Comm c, int* flag, MPI_Status* st) {
  shim_mpi::Box& b = shim_mpi::box(c);
  std::lock_guard<std::mutex> g(b.mu);
  *flag = 0;
  for (auto& m : b.q) {
    if (shim_mpi::match(m, tag)) {
      *flag = 1;
      if (st) {
        st->MPI_SOURCE = 0;
        st->MPI_TAG = m.tag;
        st->MPI_ERROR = 0;
        st->_bytes = (long long) m.data.size();
      }
      break;
    }
  }
  return 0;
}
inline int MPI_Get_count(const MPI_Status* st, MPI_Datatype t, int* count) {
  *count = (int) (st->_bytes / (t ? t : 1));
  return 0;
}
inline int MPI_Wait(MPI_Request* req, MPI_Status* st) {
  shim_request* r = *req;
  if (!r) return 0;
  if (r->is_recv && !r->done && !r->cancelled) shim_take(r->comm, r->tag, r->buf, r->cap, &r->st);
  if (st) *st = r->st;
  delete r;
  *req = MPI_REQUEST_NULL;
  return 0;
}
inline int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status* sts) {
  for (int i = 0; i < n; ++i) MPI_Wait(&reqs[i], sts ? &sts[i] : nullptr);
  return 0;
}
inline int MPI_Waitany(int n, MPI_Request* reqs, int* index, MPI_Status* st) {
  for (int i = 0; i < n; ++i) {
    if (reqs[i]) {
      MPI_Wait(&reqs[i], st);
      *index = i;
      return 0;
    }
  }
  *index = MPI_UNDEFINED;
  return 0;
}
inline int MPI_Cancel(MPI_Request* req) {
  if (req && *req) (*req)->cancelled = true;
  return 0;
}
inline int MPI_Request_free(MPI_Request* req) {
  if (req && *req) {
    delete *req;
    *req = MPI_REQUEST_NULL;
  }
  return 0;
}
inline double MPI_Wtime() { return 0.0; }
inline int MPI_Abort(MPI_Comm, int code) { std::abort(); return code; }

#endif  // ORACLE_REF_SHIM_MPI_H_
