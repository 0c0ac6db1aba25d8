// rendezvous.h — control-plane exchange between the processes of one probe
// domain (world_size > 1: one process per GPU, as bench.py runs under torchrun).
//
// A star over an abstract unix-domain socket named after the session: rank 0
// is the hub.  It carries (a) cuMem POSIX file-descriptor handles via
// SCM_RIGHTS, (b) small fixed-size blobs (mapping status, fabric handles,
// result rows) and (c) a host barrier.  A session of the form
// "tcp:<host>:<port>" runs the same star over TCP — groundwork for cross-node
// domains: blobs and the barrier only, never fds.  No data-path bytes travel here; the data path is
// NVLink P2P between the mapped allocations.  NCCL is not used (north_star).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace cdp {

class Rendezvous {
 public:
  Rendezvous() = default;
  ~Rendezvous();
  Rendezvous(const Rendezvous&) = delete;
  Rendezvous& operator=(const Rendezvous&) = delete;

  // Collective. Returns 0 or -errno; `err` gets a description.
  int connect(const std::string& session, uint32_t rank, uint32_t world, uint32_t timeout_ms, std::string* err);
  // Every rank contributes `k` fds; every rank receives world * k fds in rank order
  // (its own are dup()ed so the caller owns all returned fds).
  int allgather_fds(const int* mine, uint32_t k, std::vector<int>* all, std::string* err);
  // Every rank contributes `bytes`; every rank receives world * bytes in rank order.
  int allgather(const void* mine, size_t bytes, void* all, std::string* err);
  int barrier(std::string* err);
  void close();

  uint32_t rank() const { return rank_; }
  uint32_t world() const { return world_; }
  bool is_tcp() const { return tcp_; }

 private:
  uint32_t rank_ = 0, world_ = 1;
  int listen_fd_ = -1;
  int hub_fd_ = -1;              // client: connection to rank 0
  std::vector<int> client_fd_;   // hub: connection per rank (index 0 unused)
  uint32_t timeout_ms_ = 10000;
  bool tcp_ = false;            // session "tcp:<host>:<port>": cross-node transport, blobs only
};

}  // namespace cdp
