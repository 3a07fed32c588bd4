// comm.hpp — MPI-free process runtime (implementation: mympi_b200.cpp).  One process per GPU,
// launched by torchrun: RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT from the environment.
#ifndef MEEP_B200_COMM_HPP
#define MEEP_B200_COMM_HPP
#include <cstddef>
#include <vector>

namespace meep_b200 {

struct HostMsg {
  int peer;     // destination rank (sends) / source rank (recvs)
  void *data;
  size_t bytes;
};

void comm_init();
int comm_rank();
int comm_size();
void comm_barrier();
void comm_broadcast(int from, void *buf, size_t nbytes);
void comm_allreduce(void *buf, size_t esize, size_t count,
                    void (*combine)(void *acc, const void *in, size_t count), bool to_all);
void comm_exscan(const void *in, void *out, size_t esize,
                 void (*add)(void *acc, const void *in, size_t count));
// all ranks call this together; messages between a pair of ranks are matched in order
void comm_sendrecv_all(const std::vector<HostMsg> &sends, std::vector<HostMsg> &recvs);

} // namespace meep_b200
#endif
