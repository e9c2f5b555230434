// mpi.h — functional MPI shim (test infrastructure; the container has no MPI).
//
// Lets the UNMODIFIED reference sources (grape/**, examples/analytical_apps/**)
// compile and run, as ONE rank by default or as N ranks on one host:
//
//   GL_MPI_NP=N ./binary ...      behaves like   mpirun -n N ./binary ...
//
// MPI_Init forks N-1 copies of the calling process (it is the first thing the
// reference's InitMPIComm does: no threads, no CUDA context yet) after mapping
// one MAP_SHARED|MAP_ANONYMOUS region, so every rank sees the region at the
// same address.  The region holds one byte ring per destination rank; a send
// appends [header|payload] frames to the destination's ring, every receive /
// probe / wait first drains the caller's own ring into a process-local queue
// of complete messages and then matches (source, tag, communicator context)
// in arrival order — MPI's non-overtaking rule.  A blocked sender drains its
// own ring too, so two ranks exchanging large messages cannot deadlock.
// Collectives are built from point-to-point messages on a reserved tag of the
// same communicator context (gather to the comm's rank 0, then fan out), which
// makes them safe against concurrent traffic from other threads: the loaders
// and ParallelMessageManager run receiver threads blocked in MPI_Probe
// (grape/parallel/parallel_message_manager.h:436-459,521-525;
// grape/fragment/basic_fragment_loader.h:122-195).
// Rank 0 is the original process: it reaps the others in MPI_Finalize and
// exits non-zero when one of them failed.  Every blocking wait gives up (abort)
// when a peer died or after GL_MPI_TIMEOUT_S (default 600) seconds.
#ifndef ORACLE_REF_SHIM_MPI_H_
#define ORACLE_REF_SHIM_MPI_H_

#include <signal.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>  // real mpi.h pulls this in; grape/util.h:65 and local_io_adaptor.cc rely on it

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;  // = element size in bytes (+ kind bits above 8)
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

// datatype = size | kind << 8   (kind: 0 bytes/unsigned, 1 signed int, 2 float)
#define SHIM_DT(size, kind) ((size) | ((kind) << 8))
#define MPI_CHAR SHIM_DT(1, 1)
#define MPI_BYTE SHIM_DT(1, 0)
#define MPI_INT8_T SHIM_DT(1, 1)
#define MPI_UINT8_T SHIM_DT(1, 0)
#define MPI_INT SHIM_DT(4, 1)
#define MPI_INT32_T SHIM_DT(4, 1)
#define MPI_UINT32_T SHIM_DT(4, 0)
#define MPI_FLOAT SHIM_DT(4, 2)
#define MPI_UNSIGNED SHIM_DT(4, 0)
#define MPI_DOUBLE SHIM_DT(8, 2)
#define MPI_INT64_T SHIM_DT(8, 1)
#define MPI_UINT64_T SHIM_DT(8, 0)
#define MPI_LONG_LONG_INT SHIM_DT(8, 1)
#define MPI_LONG_LONG SHIM_DT(8, 1)
#define MPI_UNSIGNED_LONG SHIM_DT(8, 0)
#define MPI_UNSIGNED_LONG_LONG SHIM_DT(8, 0)
#define MPI_LONG SHIM_DT(8, 1)

#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3

namespace shim_mpi {

inline int dt_size(MPI_Datatype t) { return t & 0xFF; }
inline int dt_kind(MPI_Datatype t) { return (t >> 8) & 0xFF; }

constexpr int kMaxRanks = 16;
constexpr size_t kRingBytes = (size_t) 32 << 20;   // per destination rank
constexpr size_t kMaxFrame = (size_t) 4 << 20;     // payload bytes per frame
constexpr int kCollTag = 0x7FFF0000;               // reserved tag of the collectives

struct FrameHdr {
  int src;            // world rank of the sender
  int ctx;            // communicator context
  int tag;
  int pad;
  unsigned long long msg_id;   // per-sender message number (reassembly key)
  long long total;    // bytes of the whole message
  long long off;      // offset of this frame's payload
  long long len;      // payload bytes in this frame
};

struct Ring {
  std::atomic<int> lock;
  int pad;
  unsigned long long head, tail;   // byte counters (tail - head = bytes queued)
  char data[kRingBytes];
};

struct Shared {
  int np;
  std::atomic<int> dead[kMaxRanks];      // 1 = the rank process ended abnormally
  std::atomic<int> finalized[kMaxRanks];
  Ring ring[kMaxRanks];
};

struct Msg {
  int src, ctx, tag;
  std::vector<char> data;
};
struct Partial {
  Msg m;
  long long got = 0;
};

struct CommInfo {
  int ctx = 0;
  int rank = 0;                 // my rank in this communicator
  std::vector<int> group;       // comm rank -> world rank
  int children = 0;             // communicators derived so far (collective order)
};

struct State {
  Shared* sh = nullptr;
  int np = 1, rank = 0;
  std::vector<pid_t> kids;
  std::mutex mu;                              // guards everything below
  std::deque<Msg> q;                          // complete, unmatched messages (arrival order)
  std::map<std::pair<int, unsigned long long>, Partial> partial;
  std::map<int, CommInfo> comms;              // handle -> info
  int next_handle = 100;
  unsigned long long next_msg_id = 1;
  double timeout_s = 600.0;
};
inline State& S() {
  static State s;
  return s;
}

inline void ring_lock(Ring& r) {
  int expected = 0;
  while (!r.lock.compare_exchange_weak(expected, 1, std::memory_order_acquire)) {
    expected = 0;
    sched_yield();
  }
}
inline void ring_unlock(Ring& r) { r.lock.store(0, std::memory_order_release); }
inline void ring_put(Ring& r, unsigned long long pos, const void* src, size_t n) {
  const size_t o = (size_t) (pos % kRingBytes), first = std::min(n, kRingBytes - o);
  std::memcpy(r.data + o, src, first);
  if (n > first) std::memcpy(r.data, (const char*) src + first, n - first);
}
inline void ring_get(const Ring& r, unsigned long long pos, void* dst, size_t n) {
  const size_t o = (size_t) (pos % kRingBytes), first = std::min(n, kRingBytes - o);
  std::memcpy(dst, r.data + o, first);
  if (n > first) std::memcpy((char*) dst + first, r.data, n - first);
}

[[noreturn]] inline void die(const char* why) {
  State& s = S();
  fprintf(stderr, "[mpi-shim rank %d] %s\n", s.rank, why);
  if (s.sh) s.sh->dead[s.rank].store(1);
  _exit(86);
}
inline void check_peers() {
  State& s = S();
  if (!s.sh) return;
  for (int r = 0; r < s.np; ++r)
    if (r != s.rank && s.sh->dead[r].load()) die("a peer rank died; giving up");
}

// ---- incoming side: drain my ring into the local queue (caller holds s.mu) --
inline void drain_locked() {
  State& s = S();
  if (!s.sh) return;
  Ring& r = s.sh->ring[s.rank];
  ring_lock(r);
  while (r.tail - r.head >= sizeof(FrameHdr)) {
    FrameHdr h;
    ring_get(r, r.head, &h, sizeof(h));
    unsigned long long p = r.head + sizeof(h);
    Partial* part;
    auto key = std::make_pair(h.src, h.msg_id);
    auto it = s.partial.find(key);
    if (it == s.partial.end()) {
      Partial np_;
      np_.m.src = h.src;
      np_.m.ctx = h.ctx;
      np_.m.tag = h.tag;
      np_.m.data.resize((size_t) h.total);
      it = s.partial.emplace(key, std::move(np_)).first;
    }
    part = &it->second;
    if (h.len) ring_get(r, p, part->m.data.data() + (size_t) h.off, (size_t) h.len);
    part->got += h.len;
    r.head = p + (unsigned long long) h.len;
    if (part->got >= h.total) {
      s.q.push_back(std::move(part->m));
      s.partial.erase(it);
    }
  }
  ring_unlock(r);
}

struct Waiter {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  int spins = 0;
  void pause() {
    if (++spins < 200) {
      sched_yield();
    } else {
      usleep(spins < 2000 ? 20 : 200);
      if ((spins & 1023) == 0) {
        check_peers();
        double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (el > S().timeout_s) die("blocking MPI call timed out");
      }
    }
  }
};

// ---- outgoing side -----------------------------------------------------------
inline void send_bytes(int dst_world, int ctx, int tag, const void* buf, long long bytes) {
  State& s = S();
  if (dst_world == s.rank || !s.sh) {
    Msg m;
    m.src = s.rank;
    m.ctx = ctx;
    m.tag = tag;
    m.data.assign((const char*) buf, (const char*) buf + bytes);
    std::lock_guard<std::mutex> g(s.mu);
    s.q.push_back(std::move(m));
    return;
  }
  unsigned long long id;
  {
    std::lock_guard<std::mutex> g(s.mu);
    id = s.next_msg_id++;
  }
  Ring& r = s.sh->ring[dst_world];
  long long off = 0;
  bool first = true;
  Waiter w;
  while (first || off < bytes) {
    long long len = std::min<long long>(bytes - off, (long long) kMaxFrame);
    ring_lock(r);
    size_t free_b = kRingBytes - (size_t) (r.tail - r.head);
    if (free_b < sizeof(FrameHdr) + (size_t) len) {
      ring_unlock(r);
      {  // progress: my own ring may be what the peer is blocked on
        std::lock_guard<std::mutex> g(s.mu);
        drain_locked();
      }
      w.pause();
      continue;
    }
    FrameHdr h;
    h.src = s.rank;
    h.ctx = ctx;
    h.tag = tag;
    h.pad = 0;
    h.msg_id = id;
    h.total = bytes;
    h.off = off;
    h.len = len;
    unsigned long long p = r.tail;
    ring_put(r, p, &h, sizeof(h));
    p += sizeof(h);
    if (len) ring_put(r, p, (const char*) buf + off, (size_t) len);
    r.tail = p + (unsigned long long) len;
    ring_unlock(r);
    off += len;
    first = false;
  }
}

inline bool match(const Msg& m, int src_world, int ctx, int tag) {
  if (m.ctx != ctx) return false;
  if (src_world != MPI_ANY_SOURCE && m.src != src_world) return false;
  if (tag == MPI_ANY_TAG) return m.tag != kCollTag;   // wildcards never see collective traffic
  return m.tag == tag;
}

// non-blocking: find (and optionally remove) the first matching message
inline bool try_take(int src_world, int ctx, int tag, bool remove, Msg* out, MPI_Status* st_world) {
  State& s = S();
  std::lock_guard<std::mutex> g(s.mu);
  drain_locked();
  for (auto it = s.q.begin(); it != s.q.end(); ++it) {
    if (match(*it, src_world, ctx, tag)) {
      if (st_world) {
        st_world->MPI_SOURCE = it->src;
        st_world->MPI_TAG = it->tag;
        st_world->MPI_ERROR = 0;
        st_world->_bytes = (long long) it->data.size();
      }
      if (remove) {
        if (out) *out = std::move(*it);
        s.q.erase(it);
      }
      return true;
    }
  }
  return false;
}

inline CommInfo comm_info(MPI_Comm c) {
  State& s = S();
  std::lock_guard<std::mutex> g(s.mu);
  auto it = s.comms.find(c);
  if (it == s.comms.end()) {
    if (c == MPI_COMM_WORLD) {
      CommInfo w;
      w.ctx = 0;
      w.rank = s.rank;
      for (int r = 0; r < s.np; ++r) w.group.push_back(r);
      s.comms[c] = w;
      return w;
    }
    die("use of an unknown / freed communicator");
  }
  return it->second;
}
inline int to_world(const CommInfo& ci, int r) { return r == MPI_ANY_SOURCE ? MPI_ANY_SOURCE : ci.group[(size_t) r]; }
inline int to_comm(const CommInfo& ci, int world) {
  for (size_t i = 0; i < ci.group.size(); ++i)
    if (ci.group[i] == world) return (int) i;
  return world;
}

inline void recv_blocking(const CommInfo& ci, int src, int tag, void* buf, long long cap, MPI_Status* st) {
  Msg m;
  MPI_Status ws;
  Waiter w;
  while (!try_take(to_world(ci, src), ci.ctx, tag, true, &m, &ws)) w.pause();
  long long n = (long long) m.data.size();
  if (n && buf) std::memcpy(buf, m.data.data(), (size_t) std::min(n, cap));
  if (st) {
    *st = ws;
    st->MPI_SOURCE = to_comm(ci, ws.MPI_SOURCE);
  }
}

template <typename T>
inline void reduce_t(void* acc, const void* in, int n, MPI_Op op) {
  T* a = (T*) acc;
  const T* b = (const T*) in;
  for (int i = 0; i < n; ++i) {
    if (op == MPI_SUM) a[i] = a[i] + b[i];
    else if (op == MPI_MIN) a[i] = b[i] < a[i] ? b[i] : a[i];
    else a[i] = b[i] > a[i] ? b[i] : a[i];
  }
}
inline void reduce_any(void* acc, const void* in, int n, MPI_Datatype t, MPI_Op op) {
  const int sz = dt_size(t), k = dt_kind(t);
  if (k == 2 && sz == 8) reduce_t<double>(acc, in, n, op);
  else if (k == 2 && sz == 4) reduce_t<float>(acc, in, n, op);
  else if (k == 1 && sz == 8) reduce_t<long long>(acc, in, n, op);
  else if (k == 1 && sz == 4) reduce_t<int>(acc, in, n, op);
  else if (k == 1 && sz == 1) reduce_t<signed char>(acc, in, n, op);
  else if (sz == 8) reduce_t<unsigned long long>(acc, in, n, op);
  else if (sz == 4) reduce_t<unsigned int>(acc, in, n, op);
  else reduce_t<unsigned char>(acc, in, n, op);
}

// gather `bytes` from every rank of the comm at its rank 0, in rank order
inline void coll_gather0(const CommInfo& ci, const void* mine, long long bytes, std::vector<char>* all) {
  const int n = (int) ci.group.size();
  if (ci.rank == 0) {
    all->resize((size_t) bytes * n);
    std::memcpy(all->data(), mine, (size_t) bytes);
    for (int r = 1; r < n; ++r) recv_blocking(ci, r, kCollTag, all->data() + (size_t) r * bytes, bytes, nullptr);
  } else {
    send_bytes(ci.group[0], ci.ctx, kCollTag, mine, bytes);
  }
}
inline void coll_bcast0(const CommInfo& ci, void* buf, long long bytes) {
  const int n = (int) ci.group.size();
  if (ci.rank == 0) {
    for (int r = 1; r < n; ++r) send_bytes(ci.group[(size_t) r], ci.ctx, kCollTag, buf, bytes);
  } else {
    recv_blocking(ci, 0, kCollTag, buf, bytes, nullptr);
  }
}

inline void start_world(int* provided) {
  if (provided) *provided = MPI_THREAD_MULTIPLE;
  State& s = S();
  if (s.sh || s.np > 1) return;
  const char* e = getenv("GL_MPI_NP");
  int np = e ? atoi(e) : 1;
  if (const char* t = getenv("GL_MPI_TIMEOUT_S")) s.timeout_s = atof(t);
  if (np <= 1) return;
  if (np > kMaxRanks) np = kMaxRanks;
  void* mem = mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (mem == MAP_FAILED) {
    perror("mpi-shim mmap");
    _exit(87);
  }
  s.sh = (Shared*) mem;   // anonymous mappings are zero-filled: locks free, counters 0
  s.sh->np = np;
  s.np = np;
  s.rank = 0;
  fflush(stdout);
  fflush(stderr);
  for (int r = 1; r < np; ++r) {
    pid_t pid = fork();
    if (pid < 0) {
      perror("mpi-shim fork");
      _exit(88);
    }
    if (pid == 0) {
      s.rank = r;
      s.kids.clear();
      break;
    }
    s.kids.push_back(pid);
  }
  // a rank that ends without MPI_Finalize (abort, LOG(FATAL), crash) is reported dead
  atexit([] {
    State& st = S();
    if (st.sh && !st.sh->finalized[st.rank].load()) st.sh->dead[st.rank].store(1);
  });
  for (int sig : {SIGABRT, SIGSEGV, SIGBUS, SIGFPE, SIGILL, SIGTERM})
    signal(sig, [](int sg) {
      State& st = S();
      if (st.sh) st.sh->dead[st.rank].store(1);
      signal(sg, SIG_DFL);
      raise(sg);
    });
}

}  // namespace shim_mpi

struct shim_request {
  bool is_recv = false, done = false, cancelled = false;
  shim_mpi::CommInfo ci;
  int src = 0, tag = 0;
  void* buf = nullptr;
  long long cap = 0;
  MPI_Status st{};
};

inline int MPI_Init_thread(int*, char***, int, int* provided) {
  shim_mpi::start_world(provided);
  return 0;
}
inline int MPI_Init(int*, char***) {
  shim_mpi::start_world(nullptr);
  return 0;
}
inline int MPI_Comm_rank(MPI_Comm c, int* r) { *r = shim_mpi::comm_info(c).rank; return 0; }
inline int MPI_Comm_size(MPI_Comm c, int* s) { *s = (int) shim_mpi::comm_info(c).group.size(); return 0; }
inline int MPI_Get_processor_name(char* name, int* len) {
  std::strcpy(name, "localhost");
  *len = 9;
  return 0;
}
inline double MPI_Wtime() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline int MPI_Abort(MPI_Comm, int code) {
  shim_mpi::die("MPI_Abort");
  return code;
}

// ---- point to point -----------------------------------------------------------
inline int MPI_Send(const void* buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) {
  shim_mpi::CommInfo ci = shim_mpi::comm_info(c);
  shim_mpi::send_bytes(ci.group[(size_t) dst], ci.ctx, tag, buf, (long long) n * shim_mpi::dt_size(t));
  return 0;
}
inline int MPI_Isend(const void* buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* req) {
  MPI_Send(buf, n, t, dst, tag, c);   // buffered: complete on return
  shim_request* r = new shim_request;
  r->done = true;
  *req = r;
  return 0;
}
inline int MPI_Recv(void* buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* st) {
  shim_mpi::CommInfo ci = shim_mpi::comm_info(c);
  shim_mpi::recv_blocking(ci, src, tag, buf, (long long) n * shim_mpi::dt_size(t), st);
  return 0;
}
inline int MPI_Irecv(void* buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* req) {
  shim_request* r = new shim_request;
  r->is_recv = true;
  r->ci = shim_mpi::comm_info(c);
  r->src = src;
  r->tag = tag;
  r->buf = buf;
  r->cap = (long long) n * shim_mpi::dt_size(t);
  *req = r;
  return 0;
}
inline int MPI_Iprobe(int src, int tag, MPI_Comm c, int* flag, MPI_Status* st) {
  shim_mpi::CommInfo ci = shim_mpi::comm_info(c);
  MPI_Status ws;
  *flag = shim_mpi::try_take(shim_mpi::to_world(ci, src), ci.ctx, tag, false, nullptr, &ws) ? 1 : 0;
  if (*flag && st) {
    *st = ws;
    st->MPI_SOURCE = shim_mpi::to_comm(ci, ws.MPI_SOURCE);
  }
  return 0;
}
inline int MPI_Probe(int src, int tag, MPI_Comm c, MPI_Status* st) {
  int flag = 0;
  shim_mpi::Waiter w;
  for (;;) {
    MPI_Iprobe(src, tag, c, &flag, st);
    if (flag) return 0;
    w.pause();
  }
}
inline int MPI_Get_count(const MPI_Status* st, MPI_Datatype t, int* count) {
  int sz = shim_mpi::dt_size(t);
  *count = (int) (st->_bytes / (sz ? sz : 1));
  return 0;
}
// completes a request if it can be completed now (recv: a matching message is queued)
inline bool shim_test(shim_request* r) {
  if (!r || r->done || r->cancelled || !r->is_recv) return true;
  shim_mpi::Msg m;
  MPI_Status ws;
  if (!shim_mpi::try_take(shim_mpi::to_world(r->ci, r->src), r->ci.ctx, r->tag, true, &m, &ws)) return false;
  long long n = (long long) m.data.size();
  if (n && r->buf) std::memcpy(r->buf, m.data.data(), (size_t) std::min(n, r->cap));
  r->st = ws;
  r->st.MPI_SOURCE = shim_mpi::to_comm(r->ci, ws.MPI_SOURCE);
  r->done = true;
  return true;
}
inline int MPI_Wait(MPI_Request* req, MPI_Status* st) {
  shim_request* r = *req;
  if (!r) return 0;
  shim_mpi::Waiter w;
  while (!shim_test(r)) w.pause();
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
  bool any = false;
  for (int i = 0; i < n; ++i) any |= reqs[i] != MPI_REQUEST_NULL;
  if (!any) {
    *index = MPI_UNDEFINED;
    return 0;
  }
  shim_mpi::Waiter w;
  for (;;) {
    for (int i = 0; i < n; ++i) {
      if (reqs[i] && shim_test(reqs[i])) {
        if (st) *st = reqs[i]->st;
        delete reqs[i];
        reqs[i] = MPI_REQUEST_NULL;
        *index = i;
        return 0;
      }
    }
    w.pause();
  }
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

// ---- collectives ------------------------------------------------------------
inline int MPI_Barrier(MPI_Comm c) {
  shim_mpi::CommInfo ci = shim_mpi::comm_info(c);
  if (ci.group.size() <= 1) return 0;
  char x = 0;
  std::vector<char> all;
  shim_mpi::coll_gather0(ci, &x, 1, &all);
  shim_mpi::coll_bcast0(ci, &x, 1);
  return 0;
}
inline int MPI_Bcast(void* buf, int n, MPI_Datatype t, int root, MPI_Comm c) {
  shim_mpi::CommInfo ci = shim_mpi::comm_info(c);
  const int np = (int) ci.group.size();
  if (np <= 1) return 0;
  const long long bytes = (long long) n * shim_mpi::dt_size(t);
  if (ci.rank == root) {
    for (int r = 0; r < np; ++r)
      if (r != root) shim_mpi::send_bytes(ci.group[(size_t) r], ci.ctx, shim_mpi::kCollTag, buf, bytes);
  } else {
    shim_mpi::recv_blocking(ci, root, shim_mpi::kCollTag, buf, bytes, nullptr);
  }
  return 0;
}
inline int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) {
  shim_mpi::CommInfo ci = shim_mpi::comm_info(c);
  const int np = (int) ci.group.size();
  const long long bytes = s == MPI_IN_PLACE ? (long long) rn * shim_mpi::dt_size(rt) : (long long) sn * shim_mpi::dt_size(st);
  const char* mine = s == MPI_IN_PLACE ? (const char*) r + (size_t) ci.rank * bytes : (const char*) s;
  std::vector<char> own(mine, mine + bytes), all;
  shim_mpi::coll_gather0(ci, own.data(), bytes, &all);
  if (ci.rank == 0) std::memcpy(r, all.data(), (size_t) bytes * np);
  shim_mpi::coll_bcast0(ci, r, bytes * np);
  return 0;
}
inline int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* counts, const int* displs,
                          MPI_Datatype rt, MPI_Comm c) {
  shim_mpi::CommInfo ci = shim_mpi::comm_info(c);
  const int np = (int) ci.group.size();
  const int rsz = shim_mpi::dt_size(rt);
  // every rank sends its piece to every rank (pieces differ in size)
  const char* mine = s == MPI_IN_PLACE ? (const char*) r + (size_t) displs[ci.rank] * rsz : (const char*) s;
  const long long mybytes = s == MPI_IN_PLACE ? (long long) counts[ci.rank] * rsz : (long long) sn * shim_mpi::dt_size(st);
  std::vector<char> own(mine, mine + mybytes);
  for (int q = 0; q < np; ++q)
    if (q != ci.rank) shim_mpi::send_bytes(ci.group[(size_t) q], ci.ctx, shim_mpi::kCollTag, own.data(), mybytes);
  std::memcpy((char*) r + (size_t) displs[ci.rank] * rsz, own.data(), (size_t) mybytes);
  for (int q = 0; q < np; ++q)
    if (q != ci.rank)
      shim_mpi::recv_blocking(ci, q, shim_mpi::kCollTag, (char*) r + (size_t) displs[q] * rsz, (long long) counts[q] * rsz, nullptr);
  MPI_Barrier(c);   // keeps two consecutive all-to-all style collectives apart
  return 0;
}
inline int MPI_Alltoall(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) {
  shim_mpi::CommInfo ci = shim_mpi::comm_info(c);
  const int np = (int) ci.group.size();
  const long long sb = (long long) sn * shim_mpi::dt_size(st), rb = (long long) rn * shim_mpi::dt_size(rt);
  std::vector<char> src((const char*) s, (const char*) s + (size_t) sb * np);
  for (int q = 0; q < np; ++q)
    if (q != ci.rank) shim_mpi::send_bytes(ci.group[(size_t) q], ci.ctx, shim_mpi::kCollTag, src.data() + (size_t) q * sb, sb);
  std::memcpy((char*) r + (size_t) ci.rank * rb, src.data() + (size_t) ci.rank * sb, (size_t) std::min(sb, rb));
  for (int q = 0; q < np; ++q)
    if (q != ci.rank) shim_mpi::recv_blocking(ci, q, shim_mpi::kCollTag, (char*) r + (size_t) q * rb, rb, nullptr);
  MPI_Barrier(c);
  return 0;
}
inline int MPI_Gather(const void* s, int sn, MPI_Datatype st, void* r, int, MPI_Datatype, int root, MPI_Comm c) {
  shim_mpi::CommInfo ci = shim_mpi::comm_info(c);
  const int np = (int) ci.group.size();
  const long long bytes = (long long) sn * shim_mpi::dt_size(st);
  if (ci.rank == root) {
    if (s != MPI_IN_PLACE) std::memcpy((char*) r + (size_t) root * bytes, s, (size_t) bytes);
    for (int q = 0; q < np; ++q)
      if (q != root) shim_mpi::recv_blocking(ci, q, shim_mpi::kCollTag, (char*) r + (size_t) q * bytes, bytes, nullptr);
  } else {
    shim_mpi::send_bytes(ci.group[(size_t) root], ci.ctx, shim_mpi::kCollTag, s, bytes);
  }
  return 0;
}
inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
  shim_mpi::CommInfo ci = shim_mpi::comm_info(c);
  const int np = (int) ci.group.size();
  const long long bytes = (long long) n * shim_mpi::dt_size(t);
  if (ci.rank == root) {
    std::vector<char> acc((const char*) (s == MPI_IN_PLACE ? r : s), (const char*) (s == MPI_IN_PLACE ? r : s) + bytes);
    std::vector<char> in((size_t) bytes);
    // fixed rank order => the same floating-point result on every run
    for (int q = 0; q < np; ++q) {
      if (q == root) continue;
      shim_mpi::recv_blocking(ci, q, shim_mpi::kCollTag, in.data(), bytes, nullptr);
      shim_mpi::reduce_any(acc.data(), in.data(), n, t, op);
    }
    if (r) std::memcpy(r, acc.data(), (size_t) bytes);
  } else {
    shim_mpi::send_bytes(ci.group[(size_t) root], ci.ctx, shim_mpi::kCollTag, s == MPI_IN_PLACE ? r : s, bytes);
  }
  return 0;
}
inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  shim_mpi::CommInfo ci = shim_mpi::comm_info(c);
  const long long bytes = (long long) n * shim_mpi::dt_size(t);
  if (s != MPI_IN_PLACE && s != r) std::memcpy(r, s, (size_t) bytes);
  if (ci.group.size() <= 1) return 0;
  MPI_Reduce(MPI_IN_PLACE, r, n, t, op, 0, c);
  if (ci.rank != 0) {
    // non-root ranks sent r above (MPI_IN_PLACE on a non-root sends r)
  }
  shim_mpi::coll_bcast0(ci, r, bytes);
  return 0;
}

// ---- communicators ----------------------------------------------------------
inline MPI_Comm shim_register(const shim_mpi::CommInfo& ci) {
  shim_mpi::State& s = shim_mpi::S();
  std::lock_guard<std::mutex> g(s.mu);
  int h = ++s.next_handle;
  s.comms[h] = ci;
  return h;
}
inline int shim_child_ctx(MPI_Comm parent) {
  // collective calls on one communicator happen in the same order on all of
  // its ranks, so (parent context, n-th child) names the same communicator
  shim_mpi::State& s = shim_mpi::S();
  shim_mpi::comm_info(parent);
  std::lock_guard<std::mutex> g(s.mu);
  shim_mpi::CommInfo& p = s.comms[parent];
  const int k = ++p.children;
  unsigned long long x = (unsigned long long) (unsigned) p.ctx * 0x9E3779B97F4A7C15ull + (unsigned long long) k * 0xD1B54A32D192ED03ull;
  x ^= x >> 29;
  int ctx = (int) (x & 0x3FFFFFFF);
  return ctx ? ctx : 1;
}
inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm* n) {
  shim_mpi::CommInfo ci = shim_mpi::comm_info(c);
  ci.ctx = shim_child_ctx(c);
  ci.children = 0;
  *n = shim_register(ci);
  return 0;
}
inline int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* n) {
  shim_mpi::CommInfo ci = shim_mpi::comm_info(c);
  const int np = (int) ci.group.size();
  int mine[2] = {color, key};
  std::vector<int> all((size_t) 2 * np);
  MPI_Allgather(mine, 2, MPI_INT, all.data(), 2, MPI_INT, c);
  const int base = shim_child_ctx(c);
  if (color == MPI_UNDEFINED) {
    *n = MPI_COMM_NULL;
    return 0;
  }
  std::vector<std::pair<std::pair<int, int>, int>> members;   // ((key, parent rank), world)
  for (int r = 0; r < np; ++r)
    if (all[(size_t) 2 * r] == color) members.push_back({{all[(size_t) 2 * r + 1], r}, ci.group[(size_t) r]});
  std::sort(members.begin(), members.end());
  shim_mpi::CommInfo ni;
  ni.ctx = (int) (((unsigned) base * 31u + (unsigned) color * 7919u + 17u) & 0x3FFFFFFF);
  if (!ni.ctx) ni.ctx = 2;
  for (size_t i = 0; i < members.size(); ++i) {
    ni.group.push_back(members[i].second);
    if (members[i].second == ci.group[(size_t) ci.rank]) ni.rank = (int) i;
  }
  *n = shim_register(ni);
  return 0;
}
inline int MPI_Comm_free(MPI_Comm* c) {
  if (c && *c > 0) {
    shim_mpi::State& s = shim_mpi::S();
    std::lock_guard<std::mutex> g(s.mu);
    s.comms.erase(*c);
  }
  if (c) *c = MPI_COMM_NULL;
  return 0;
}

inline int MPI_Finalize() {
  shim_mpi::State& s = shim_mpi::S();
  if (!s.sh) return 0;
  MPI_Barrier(MPI_COMM_WORLD);
  s.sh->finalized[s.rank].store(1);
  if (s.rank == 0) {
    int bad = 0;
    for (pid_t k : s.kids) {
      int status = 0;
      if (waitpid(k, &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) ++bad;
    }
    if (bad) {
      fprintf(stderr, "[mpi-shim] %d rank process(es) failed\n", bad);
      fflush(stdout);
      _exit(1);
    }
  }
  return 0;
}

#endif  // ORACLE_REF_SHIM_MPI_H_
