// CPU check of the multi-rank MPI shim (tests/test_mpi_shim.py)
#include <mpi.h>
#include <cstdlib>
#include <cstdio>
#include <vector>
#include <thread>
int main(int argc, char** argv) {
  int prov; MPI_Init_thread(&argc, &argv, MPI_THREAD_MULTIPLE, &prov);
  int r, n; MPI_Comm_rank(MPI_COMM_WORLD, &r); MPI_Comm_size(MPI_COMM_WORLD, &n);
  if (getenv("SHIM_CHECK_DIE") && r == 1) abort();
  MPI_Comm c2; MPI_Comm_dup(MPI_COMM_WORLD, &c2);
  long long v = r + 1, s = 0; MPI_Allreduce(&v, &s, 1, MPI_LONG_LONG_INT, MPI_SUM, c2);
  double d = r * 0.5, dm = 0; MPI_Allreduce(&d, &dm, 1, MPI_DOUBLE, MPI_MAX, c2);
  std::vector<int> all(n); int me = r * 10; MPI_Allgather(&me, 1, MPI_INT, all.data(), 1, MPI_INT, c2);
  // big exchange: everyone sends 80 MB to next, concurrently
  size_t big = 10u << 20; std::vector<long long> out(big, r), in(big, -1);
  MPI_Request rq[2];
  MPI_Irecv(in.data(), (int) big, MPI_LONG_LONG_INT, (r + n - 1) % n, 5, c2, &rq[0]);
  MPI_Isend(out.data(), (int) big, MPI_LONG_LONG_INT, (r + 1) % n, 5, c2, &rq[1]);
  MPI_Waitall(2, rq, MPI_STATUSES_IGNORE);
  bool ok = in[0] == (r + n - 1) % n && in[big - 1] == (r + n - 1) % n;
  // probe thread stopped by zero-length self send
  std::thread t([&] { MPI_Status st; MPI_Probe(MPI_ANY_SOURCE, MPI_ANY_TAG, c2, &st); int cnt; MPI_Get_count(&st, MPI_CHAR, &cnt); char b; MPI_Recv(&b, 0, MPI_CHAR, st.MPI_SOURCE, st.MPI_TAG, c2, MPI_STATUS_IGNORE); ok = ok && cnt == 0 && st.MPI_SOURCE == r; });
  MPI_Barrier(c2);
  MPI_Send(nullptr, 0, MPI_CHAR, r, 9, c2);
  t.join();
  MPI_Comm sp; MPI_Comm_split(MPI_COMM_WORLD, r % 2, r, &sp); int sr, sn; MPI_Comm_rank(sp, &sr); MPI_Comm_size(sp, &sn);
  int x = r; MPI_Bcast(&x, 1, MPI_INT, 0, sp);
  printf("rank %d/%d sum=%lld max=%.1f all=[%d..%d] big_ok=%d split=%d/%d bcast=%d\n", r, n, s, dm, all[0], all[n-1], (int) ok, sr, sn, x);
  MPI_Finalize();
  return ok ? 0 : 1;
}
