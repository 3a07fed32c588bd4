// mympi_b200.cpp — MPI-free multi-process runtime for the reference's process model.
//
// The reference funnels every inter-process operation through src/mympi.cpp:
// rank/size queries, small host-side reductions (sum_to_all, and_to_all, ...) and broadcasts.
// MPI is not available in this environment and the B200 build is launched as one process per
// GPU by torchrun (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the environment), so this
// file interposes those entry points with an implementation over TCP sockets (star topology
// through rank 0; payloads are a few bytes to a few kB and occur only at set-up, at
// connect_chunks and when a user asks for a flux/probe — never inside the steady-state step).
// The bulk halo data never goes through here: it moves device-to-device (step.cpp).
//
// With WORLD_SIZE unset or 1 every function degenerates to the reference's serial behaviour.
#include <arpa/inet.h>
#include <errno.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>
#include <complex>
#include <vector>

#include "meep.hpp"
#include "comm.hpp"

namespace meep_b200 {

namespace {
int g_rank = 0, g_size = 1;
bool g_init = false;
std::vector<int> g_peer; // rank 0: socket to every other rank; others: [0] = socket to rank 0

void die(const char *what) {
  fprintf(stderr, "meep_b200 comm (rank %d): %s: %s\n", g_rank, what, strerror(errno));
  abort();
}

void send_all(int fd, const void *buf, size_t n) {
  const char *p = (const char *)buf;
  while (n) {
    ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) {
      if (errno == EINTR) continue;
      die("send");
    }
    p += k;
    n -= (size_t)k;
  }
}

void recv_all(int fd, void *buf, size_t n) {
  char *p = (char *)buf;
  while (n) {
    ssize_t k = ::recv(fd, p, n, 0);
    if (k <= 0) {
      if (k < 0 && errno == EINTR) continue;
      die("recv (peer closed?)");
    }
    p += k;
    n -= (size_t)k;
  }
}

int env_int(const char *a, const char *b, int dflt) {
  const char *v = getenv(a);
  if (!v || !*v) v = b ? getenv(b) : NULL;
  return (v && *v) ? atoi(v) : dflt;
}
} // namespace

void comm_init() {
  if (g_init) return;
  g_init = true;
  g_size = env_int("MEEP_B200_WORLD_SIZE", "WORLD_SIZE", 1);
  g_rank = env_int("MEEP_B200_RANK", "RANK", 0);
  if (g_size <= 1) {
    g_size = 1;
    g_rank = 0;
    return;
  }
  const char *addr = getenv("MASTER_ADDR");
  if (!addr || !*addr) addr = "127.0.0.1";
  const int port = env_int("MEEP_B200_PORT", NULL, env_int("MASTER_PORT", NULL, 29500) + 37);
  if (g_rank == 0) {
    int ls = socket(AF_INET, SOCK_STREAM, 0);
    if (ls < 0) die("socket");
    int one = 1;
    setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in sa;
    memset(&sa, 0, sizeof(sa));
    sa.sin_family = AF_INET;
    sa.sin_addr.s_addr = htonl(INADDR_ANY);
    sa.sin_port = htons((uint16_t)port);
    // (a port that is busy for a moment — a short-lived client socket that happened to get this
    // number — is retried for a few seconds before giving up; the other ranks keep knocking meanwhile)
    for (int attempt = 0;; ++attempt) {
      if (bind(ls, (sockaddr *)&sa, sizeof(sa)) == 0) break;
      if (errno != EADDRINUSE || attempt >= 50) die("bind");
      usleep(100 * 1000);
    }
    if (listen(ls, g_size) < 0) die("listen");
    g_peer.assign(g_size, -1);
    for (int k = 1; k < g_size; ++k) {
      int fd = accept(ls, NULL, NULL);
      if (fd < 0) die("accept");
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
      int r = -1;
      recv_all(fd, &r, sizeof(r));
      if (r <= 0 || r >= g_size || g_peer[r] != -1) die("bad rank in handshake");
      g_peer[r] = fd;
    }
    close(ls);
  }
  else {
    addrinfo hints, *res = NULL;
    memset(&hints, 0, sizeof(hints));
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    char ps[16];
    snprintf(ps, sizeof(ps), "%d", port);
    if (getaddrinfo(addr, ps, &hints, &res) != 0 || !res) die("getaddrinfo(MASTER_ADDR)");
    int fd = -1;
    for (int attempt = 0; attempt < 600; ++attempt) { // rank 0 may not be listening yet
      fd = socket(AF_INET, SOCK_STREAM, 0);
      if (fd < 0) die("socket");
      if (connect(fd, res->ai_addr, res->ai_addrlen) == 0) break;
      close(fd);
      fd = -1;
      usleep(100000);
    }
    freeaddrinfo(res);
    if (fd < 0) die("connect to rank 0");
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    send_all(fd, &g_rank, sizeof(g_rank));
    g_peer.assign(1, fd);
  }
}

int comm_rank() {
  comm_init();
  return g_rank;
}
int comm_size() {
  comm_init();
  return g_size;
}

// element-wise reduction of `count` elements of `esize` bytes; combine(acc, in) runs on rank 0
void comm_allreduce(void *buf, size_t esize, size_t count,
                    void (*combine)(void *acc, const void *in, size_t count), bool to_all) {
  comm_init();
  if (g_size == 1) return;
  const size_t n = esize * count;
  if (g_rank == 0) {
    std::vector<char> tmp(n);
    for (int r = 1; r < g_size; ++r) {
      recv_all(g_peer[r], tmp.data(), n);
      combine(buf, tmp.data(), count);
    }
    if (to_all)
      for (int r = 1; r < g_size; ++r)
        send_all(g_peer[r], buf, n);
  }
  else {
    send_all(g_peer[0], buf, n);
    if (to_all) recv_all(g_peer[0], buf, n);
  }
}

void comm_broadcast(int from, void *buf, size_t n) {
  comm_init();
  if (g_size == 1) return;
  // relay through rank 0
  if (from != 0) {
    if (g_rank == from) send_all(g_peer[0], buf, n);
    if (g_rank == 0) recv_all(g_peer[from], buf, n);
  }
  if (g_rank == 0) {
    for (int r = 1; r < g_size; ++r)
      send_all(g_peer[r], buf, n);
  }
  else
    recv_all(g_peer[0], buf, n);
}

void comm_barrier() {
  char c = 0;
  comm_allreduce(&c, 1, 1, [](void *, const void *, size_t) {}, true);
}

// exclusive prefix: gather to rank 0, scan, scatter
void comm_exscan(const void *in, void *out, size_t esize,
                 void (*add)(void *acc, const void *in, size_t count)) {
  comm_init();
  if (g_size == 1) {
    memcpy(out, in, esize);
    return;
  }
  if (g_rank == 0) {
    std::vector<char> vals((size_t)g_size * esize), acc(esize);
    memcpy(vals.data(), in, esize);
    for (int r = 1; r < g_size; ++r)
      recv_all(g_peer[r], vals.data() + (size_t)r * esize, esize);
    // inclusive scan (MPI_Scan semantics, as the reference's partial_sum_to_all uses)
    memcpy(acc.data(), vals.data(), esize);
    memcpy(out, acc.data(), esize);
    for (int r = 1; r < g_size; ++r) {
      add(acc.data(), vals.data() + (size_t)r * esize, 1);
      send_all(g_peer[r], acc.data(), esize);
    }
  }
  else {
    send_all(g_peer[0], in, esize);
    recv_all(g_peer[0], out, esize);
  }
}

// point-to-point host payload between arbitrary ranks (used for the NCCL id / IPC handles and by
// the emulator's halo transport); relayed through rank 0.  Tag-less: both sides call in the same
// global order.
void comm_sendrecv_all(const std::vector<HostMsg> &sends, std::vector<HostMsg> &recvs) {
  comm_init();
  if (g_size == 1) return;
  // protocol: every rank ships (dst, nbytes, payload)* to rank 0, which routes.
  auto pack = [](const std::vector<HostMsg> &v, std::vector<char> &out) {
    uint64_t n = v.size();
    out.insert(out.end(), (char *)&n, (char *)&n + 8);
    for (const HostMsg &m : v) {
      int32_t peer = m.peer;
      uint64_t nb = m.bytes;
      out.insert(out.end(), (char *)&peer, (char *)&peer + 4);
      out.insert(out.end(), (char *)&nb, (char *)&nb + 8);
      out.insert(out.end(), (const char *)m.data, (const char *)m.data + nb);
    }
  };
  struct Routed {
    int src, dst;
    std::vector<char> payload;
  };
  auto deliver = [&](const std::vector<Routed> &mine) {
    // match in order per source
    std::vector<size_t> cursor(g_size, 0);
    for (HostMsg &r : recvs) {
      bool found = false;
      for (size_t k = cursor[r.peer]; k < mine.size(); ++k)
        if (mine[k].src == r.peer) {
          if (mine[k].payload.size() != r.bytes) die("message size mismatch");
          memcpy(r.data, mine[k].payload.data(), r.bytes);
          cursor[r.peer] = k + 1;
          found = true;
          break;
        }
      if (!found) die("expected message not received");
    }
  };
  std::vector<char> mybuf;
  pack(sends, mybuf);
  if (g_rank == 0) {
    std::vector<std::vector<Routed> > inbox(g_size);
    auto unpack = [&](int src, const std::vector<char> &buf) {
      size_t o = 0;
      uint64_t n;
      memcpy(&n, buf.data(), 8);
      o = 8;
      for (uint64_t k = 0; k < n; ++k) {
        int32_t dst;
        uint64_t nb;
        memcpy(&dst, buf.data() + o, 4);
        memcpy(&nb, buf.data() + o + 4, 8);
        o += 12;
        Routed r;
        r.src = src;
        r.dst = dst;
        r.payload.assign(buf.begin() + o, buf.begin() + o + nb);
        o += nb;
        inbox[dst].push_back(std::move(r));
      }
    };
    unpack(0, mybuf);
    for (int r = 1; r < g_size; ++r) {
      uint64_t nb;
      recv_all(g_peer[r], &nb, 8);
      std::vector<char> buf(nb);
      recv_all(g_peer[r], buf.data(), nb);
      unpack(r, buf);
    }
    for (int r = 1; r < g_size; ++r) {
      std::vector<char> out;
      uint64_t n = inbox[r].size();
      out.insert(out.end(), (char *)&n, (char *)&n + 8);
      for (const Routed &m : inbox[r]) {
        int32_t src = m.src;
        uint64_t nb = m.payload.size();
        out.insert(out.end(), (char *)&src, (char *)&src + 4);
        out.insert(out.end(), (char *)&nb, (char *)&nb + 8);
        out.insert(out.end(), m.payload.begin(), m.payload.end());
      }
      uint64_t tot = out.size();
      send_all(g_peer[r], &tot, 8);
      send_all(g_peer[r], out.data(), tot);
    }
    deliver(inbox[0]);
  }
  else {
    uint64_t nb = mybuf.size();
    send_all(g_peer[0], &nb, 8);
    send_all(g_peer[0], mybuf.data(), nb);
    uint64_t tot;
    recv_all(g_peer[0], &tot, 8);
    std::vector<char> buf(tot);
    recv_all(g_peer[0], buf.data(), tot);
    std::vector<Routed> mine;
    size_t o = 8;
    uint64_t n;
    memcpy(&n, buf.data(), 8);
    for (uint64_t k = 0; k < n; ++k) {
      int32_t src;
      uint64_t nb2;
      memcpy(&src, buf.data() + o, 4);
      memcpy(&nb2, buf.data() + o + 4, 8);
      o += 12;
      Routed r;
      r.src = src;
      r.dst = g_rank;
      r.payload.assign(buf.begin() + o, buf.begin() + o + nb2);
      o += nb2;
      mine.push_back(std::move(r));
    }
    deliver(mine);
  }
}

} // namespace meep_b200

// ---- interposed mympi.cpp entry points (reference src/mympi.cpp; declarations
//      src/meep/mympi.hpp:51-120) ------------------------------------------------------------------
namespace meep {
using namespace meep_b200;

template <typename T> static void add_fn(void *a, const void *b, size_t n) {
  T *x = (T *)a;
  const T *y = (const T *)b;
  for (size_t i = 0; i < n; ++i)
    x[i] += y[i];
}
template <typename T> static void max_fn(void *a, const void *b, size_t n) {
  T *x = (T *)a;
  const T *y = (const T *)b;
  for (size_t i = 0; i < n; ++i)
    if (y[i] > x[i]) x[i] = y[i];
}
template <typename T> static void min_fn(void *a, const void *b, size_t n) {
  T *x = (T *)a;
  const T *y = (const T *)b;
  for (size_t i = 0; i < n; ++i)
    if (y[i] < x[i]) x[i] = y[i];
}
template <typename T> static void or_fn(void *a, const void *b, size_t n) {
  T *x = (T *)a;
  const T *y = (const T *)b;
  for (size_t i = 0; i < n; ++i)
    x[i] = x[i] | y[i];
}
static void lor_fn(void *a, const void *b, size_t n) {
  int *x = (int *)a;
  const int *y = (const int *)b;
  for (size_t i = 0; i < n; ++i)
    x[i] = (x[i] || y[i]) ? 1 : 0;
}
static void land_fn(void *a, const void *b, size_t n) {
  int *x = (int *)a;
  const int *y = (const int *)b;
  for (size_t i = 0; i < n; ++i)
    x[i] = (x[i] && y[i]) ? 1 : 0;
}

template <typename TI, typename TO>
static void reduce_vec(const TI *in, TO *out, int size, void (*fn)(void *, const void *, size_t),
                       bool to_all) {
  std::vector<TO> tmp(size);
  for (int i = 0; i < size; ++i)
    tmp[i] = (TO)in[i];
  comm_allreduce(tmp.data(), sizeof(TO), (size_t)size, fn, to_all);
  for (int i = 0; i < size; ++i)
    out[i] = tmp[i];
}

void all_wait() { comm_barrier(); }
int count_processors() { return comm_size(); }
int my_rank() { return comm_rank(); }
bool am_really_master() { return comm_rank() == 0; }
int my_global_rank() { return comm_rank(); }
bool with_mpi() { return comm_size() > 1; }

void send(int from, int to, double *data, int size) {
  if (from == to) return;
  std::vector<HostMsg> s, r;
  if (comm_rank() == from) s.push_back(HostMsg{to, data, sizeof(double) * (size_t)size});
  if (comm_rank() == to) r.push_back(HostMsg{from, data, sizeof(double) * (size_t)size});
  comm_sendrecv_all(s, r);
}

void broadcast(int from, float *data, int size) { comm_broadcast(from, data, sizeof(float) * size); }
void broadcast(int from, double *data, int size) { comm_broadcast(from, data, sizeof(double) * size); }
void broadcast(int from, char *data, int size) { comm_broadcast(from, data, (size_t)size); }
void broadcast(int from, int *data, int size) { comm_broadcast(from, data, sizeof(int) * size); }
void broadcast(int from, size_t *data, int size) { comm_broadcast(from, data, sizeof(size_t) * size); }
void broadcast(int from, std::complex<double> *data, int size) {
  comm_broadcast(from, data, sizeof(std::complex<double>) * size);
}
std::complex<double> broadcast(int from, std::complex<double> data) {
  comm_broadcast(from, &data, sizeof(data));
  return data;
}
double broadcast(int from, double data) {
  comm_broadcast(from, &data, sizeof(data));
  return data;
}
int broadcast(int from, int data) {
  comm_broadcast(from, &data, sizeof(data));
  return data;
}
bool broadcast(int from, bool b) {
  int v = b;
  comm_broadcast(from, &v, sizeof(v));
  return v != 0;
}

double max_to_master(double in) {
  comm_allreduce(&in, sizeof(double), 1, max_fn<double>, false);
  return in;
}
double max_to_all(double in) {
  comm_allreduce(&in, sizeof(double), 1, max_fn<double>, true);
  return in;
}
int max_to_all(int in) {
  comm_allreduce(&in, sizeof(int), 1, max_fn<int>, true);
  return in;
}
int min_to_all(int in) {
  comm_allreduce(&in, sizeof(int), 1, min_fn<int>, true);
  return in;
}
float sum_to_master(float in) {
  comm_allreduce(&in, sizeof(float), 1, add_fn<float>, false);
  return in;
}
double sum_to_master(double in) {
  comm_allreduce(&in, sizeof(double), 1, add_fn<double>, false);
  return in;
}
double sum_to_all(double in) {
  comm_allreduce(&in, sizeof(double), 1, add_fn<double>, true);
  return in;
}
void sum_to_all(const float *in, float *out, int size) { reduce_vec(in, out, size, add_fn<float>, true); }
void sum_to_all(const double *in, double *out, int size) { reduce_vec(in, out, size, add_fn<double>, true); }
void sum_to_master(const float *in, float *out, int size) { reduce_vec(in, out, size, add_fn<float>, false); }
void sum_to_master(const double *in, double *out, int size) { reduce_vec(in, out, size, add_fn<double>, false); }
void sum_to_all(const float *in, double *out, int size) { reduce_vec(in, out, size, add_fn<double>, true); }
void sum_to_all(const std::complex<float> *in, std::complex<double> *out, int size) {
  reduce_vec((const float *)in, (double *)out, 2 * size, add_fn<double>, true);
}
void sum_to_all(const std::complex<double> *in, std::complex<double> *out, int size) {
  reduce_vec((const double *)in, (double *)out, 2 * size, add_fn<double>, true);
}
void sum_to_all(const std::complex<float> *in, std::complex<float> *out, int size) {
  reduce_vec((const float *)in, (float *)out, 2 * size, add_fn<float>, true);
}
void sum_to_master(const std::complex<float> *in, std::complex<float> *out, int size) {
  reduce_vec((const float *)in, (float *)out, 2 * size, add_fn<float>, false);
}
void sum_to_master(const std::complex<double> *in, std::complex<double> *out, int size) {
  reduce_vec((const double *)in, (double *)out, 2 * size, add_fn<double>, false);
}
long double sum_to_all(long double in) {
  comm_allreduce(&in, sizeof(long double), 1, add_fn<long double>, true);
  return in;
}
std::complex<double> sum_to_all(std::complex<double> in) {
  comm_allreduce(&in, sizeof(double), 2, add_fn<double>, true);
  return in;
}
std::complex<long double> sum_to_all(std::complex<long double> in) {
  comm_allreduce(&in, sizeof(long double), 2, add_fn<long double>, true);
  return in;
}
int sum_to_all(int in) {
  comm_allreduce(&in, sizeof(int), 1, add_fn<int>, true);
  return in;
}
int partial_sum_to_all(int in) {
  int out = in;
  comm_exscan(&in, &out, sizeof(int), add_fn<int>);
  return out;
}
size_t sum_to_all(size_t in) {
  comm_allreduce(&in, sizeof(size_t), 1, add_fn<size_t>, true);
  return in;
}
size_t partial_sum_to_all(size_t in) {
  size_t out = in;
  comm_exscan(&in, &out, sizeof(size_t), add_fn<size_t>);
  return out;
}
void sum_to_all(const size_t *in, size_t *out, int size) { reduce_vec(in, out, size, add_fn<size_t>, true); }
void sum_to_master(const size_t *in, size_t *out, int size) { reduce_vec(in, out, size, add_fn<size_t>, false); }
bool or_to_all(bool in) {
  int v = in;
  comm_allreduce(&v, sizeof(int), 1, lor_fn, true);
  return v != 0;
}
void or_to_all(const int *in, int *out, int size) { reduce_vec(in, out, size, lor_fn, true); }
void bw_or_to_all(const size_t *in, size_t *out, int size) { reduce_vec(in, out, size, or_fn<size_t>, true); }
bool and_to_all(bool in) {
  int v = in;
  comm_allreduce(&v, sizeof(int), 1, land_fn, true);
  return v != 0;
}
void and_to_all(const int *in, int *out, int size) { reduce_vec(in, out, size, land_fn, true); }

void begin_critical_section(int) {
  // ranks enter one after the other: wait until all lower ranks have left
  for (int r = 0; r < comm_rank(); ++r)
    comm_barrier();
}
void end_critical_section(int) {
  for (int r = comm_rank(); r < comm_size(); ++r)
    comm_barrier();
}

int divide_parallel_processes(int numgroups) {
  if (numgroups > 1)
    meep::abort("meep_b200: divide_parallel_processes is not supported by the MPI-free runtime");
  return 0;
}
void begin_global_communications(void) {}
void end_global_communications(void) {}
void end_divide_parallel(void) {}

} // namespace meep
