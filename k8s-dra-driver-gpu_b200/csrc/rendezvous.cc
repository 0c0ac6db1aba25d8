// rendezvous.cc — see rendezvous.h.
#include "rendezvous.h"

#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include "../../include/cdprobe.h"

namespace cdp {
namespace {

constexpr uint32_t kHelloMagic = 0xCD9B0B01u;

double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
}

socklen_t make_addr(const std::string& session, sockaddr_un* a) {
  memset(a, 0, sizeof(*a));
  a->sun_family = AF_UNIX;
  std::string name = "cdprobe." + session;
  if (name.size() > sizeof(a->sun_path) - 2) name.resize(sizeof(a->sun_path) - 2);
  a->sun_path[0] = '\0';  // abstract namespace: no filesystem entry, vanishes with the process
  memcpy(a->sun_path + 1, name.data(), name.size());
  return (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + name.size());
}

// "tcp:<host>:<port>" — the cross-node transport: rank 0 listens on <port>, the others
// connect to <host>.  Only blobs can travel (fabric handles are 64-byte blobs); fds cannot leave a node.
bool parse_tcp(const std::string& session, std::string* host, std::string* port) {
  if (session.compare(0, 4, "tcp:") != 0) return false;
  const size_t c = session.rfind(':');
  if (c == std::string::npos || c <= 4) return false;
  *host = session.substr(4, c - 4);
  *port = session.substr(c + 1);
  return !host->empty() && !port->empty();
}

void set_timeouts(int fd, uint32_t ms) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));  // no-op (ENOTSUP) on unix sockets
  timeval tv;
  tv.tv_sec = ms / 1000;
  tv.tv_usec = (ms % 1000) * 1000;
  setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
}

int send_all(int fd, const void* buf, size_t n) {
  const char* p = static_cast<const char*>(buf);
  while (n) {
    ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k < 0) {
      if (errno == EINTR) continue;
      return -errno;
    }
    p += k;
    n -= (size_t)k;
  }
  return 0;
}

int recv_all(int fd, void* buf, size_t n) {
  char* p = static_cast<char*>(buf);
  while (n) {
    ssize_t k = ::recv(fd, p, n, 0);
    if (k == 0) return -ECONNRESET;
    if (k < 0) {
      if (errno == EINTR) continue;
      return -errno;
    }
    p += k;
    n -= (size_t)k;
  }
  return 0;
}

int send_fds(int sock, const int* fds, uint32_t n) {
  char payload = 'F';
  iovec iov{&payload, 1};
  std::vector<char> ctrl(CMSG_SPACE(sizeof(int) * n), 0);
  msghdr msg;
  memset(&msg, 0, sizeof(msg));
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl.data();
  msg.msg_controllen = ctrl.size();
  cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  cm->cmsg_level = SOL_SOCKET;
  cm->cmsg_type = SCM_RIGHTS;
  cm->cmsg_len = CMSG_LEN(sizeof(int) * n);
  memcpy(CMSG_DATA(cm), fds, sizeof(int) * n);
  for (;;) {
    ssize_t k = ::sendmsg(sock, &msg, MSG_NOSIGNAL);
    if (k < 0 && errno == EINTR) continue;
    return k < 0 ? -errno : 0;
  }
}

int recv_fds(int sock, int* fds, uint32_t n) {
  char payload = 0;
  iovec iov{&payload, 1};
  std::vector<char> ctrl(CMSG_SPACE(sizeof(int) * n), 0);
  msghdr msg;
  memset(&msg, 0, sizeof(msg));
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl.data();
  msg.msg_controllen = ctrl.size();
  ssize_t k;
  do {
    k = ::recvmsg(sock, &msg, MSG_CMSG_CLOEXEC);
  } while (k < 0 && errno == EINTR);
  if (k < 0) return -errno;
  if (k == 0) return -ECONNRESET;
  cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  if (cm == nullptr || cm->cmsg_level != SOL_SOCKET || cm->cmsg_type != SCM_RIGHTS ||
      cm->cmsg_len != CMSG_LEN(sizeof(int) * n) || (msg.msg_flags & MSG_CTRUNC)) {
    // whatever descriptors did arrive are ours now: close them, or every malformed message leaks fds
    for (cmsghdr* c = CMSG_FIRSTHDR(&msg); c != nullptr; c = CMSG_NXTHDR(&msg, c)) {
      if (c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS || c->cmsg_len < CMSG_LEN(0)) continue;
      const size_t cnt = (c->cmsg_len - CMSG_LEN(0)) / sizeof(int);
      for (size_t i = 0; i < cnt; ++i) {
        int fd;
        memcpy(&fd, CMSG_DATA(c) + i * sizeof(int), sizeof(int));
        if (fd >= 0) ::close(fd);
      }
    }
    return -EPROTO;
  }
  memcpy(fds, CMSG_DATA(cm), sizeof(int) * n);
  return 0;
}

// Unix transport: the peer must be a process of the same user.  The abstract socket name is visible to
// every process in the network namespace; without this check any of them could claim a rank and be handed
// SCM_RIGHTS descriptors of every rank's GPU allocation.
bool same_user(int fd) {
  ucred cred;
  socklen_t len = sizeof(cred);
  if (getsockopt(fd, SOL_SOCKET, SO_PEERCRED, &cred, &len) != 0 || len != sizeof(cred)) return false;
  return cred.uid == geteuid();
}

}  // namespace

Rendezvous::~Rendezvous() { close(); }

void Rendezvous::close() {
  if (hub_fd_ >= 0) ::close(hub_fd_);
  hub_fd_ = -1;
  for (int& fd : client_fd_)
    if (fd >= 0) {
      ::close(fd);
      fd = -1;
    }
  client_fd_.clear();
  if (listen_fd_ >= 0) ::close(listen_fd_);
  listen_fd_ = -1;
}

int Rendezvous::connect(const std::string& session, uint32_t rank, uint32_t world, uint32_t timeout_ms,
                        std::string* err) {
  rank_ = rank;
  world_ = world;
  timeout_ms_ = timeout_ms ? timeout_ms : 10000;
  if (world <= 1) return 0;
  if (rank >= world || session.empty()) {
    if (err) *err = "rendezvous: bad rank/world/session";
    return -EINVAL;
  }
  sockaddr_storage addr;
  socklen_t alen = 0;
  int family = AF_UNIX;
  memset(&addr, 0, sizeof(addr));
  std::string host, port;
  tcp_ = parse_tcp(session, &host, &port);
  if (session.compare(0, 4, "tcp:") == 0 && !tcp_) {
    if (err) *err = "rendezvous: expected tcp:<host>:<port>";
    return -EINVAL;
  }
  if (tcp_) {
    addrinfo hints, *res = nullptr;
    memset(&hints, 0, sizeof(hints));
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    // rank 0 binds the address the session names (the daemons' own DNS name / pod IP), not INADDR_ANY
    if (getaddrinfo(host.c_str(), port.c_str(), &hints, &res) != 0 || res == nullptr) {
      if (err) *err = "rendezvous: cannot resolve " + host + ":" + port;
      return -EHOSTUNREACH;
    }
    memcpy(&addr, res->ai_addr, res->ai_addrlen);
    alen = (socklen_t)res->ai_addrlen;
    family = AF_INET;
    freeaddrinfo(res);
  } else {
    alen = make_addr(session, reinterpret_cast<sockaddr_un*>(&addr));
  }
  const double t_end = now_ms() + timeout_ms_;
  if (rank == 0) {
    listen_fd_ = ::socket(family, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (listen_fd_ < 0) return -errno;
    if (tcp_) {
      int one = 1;
      setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    }
    if (::bind(listen_fd_, reinterpret_cast<sockaddr*>(&addr), alen) < 0 || ::listen(listen_fd_, (int)world) < 0) {
      const int e = errno;
      if (err) *err = std::string("rendezvous: bind/listen: ") + strerror(e);
      return -e;
    }
    client_fd_.assign(world, -1);
    for (uint32_t got = 1; got < world;) {
      pollfd pfd{listen_fd_, POLLIN, 0};
      const double left = t_end - now_ms();
      if (left <= 0 || ::poll(&pfd, 1, (int)left) <= 0) {
        if (err) *err = "rendezvous: timed out waiting for peers";
        return -ETIMEDOUT;
      }
      int fd = ::accept4(listen_fd_, nullptr, nullptr, SOCK_CLOEXEC);
      if (fd < 0) continue;
      if (!tcp_ && !same_user(fd)) {
        ::close(fd);
        continue;
      }
      // a peer says hello at once; a silent connection gets one second, not the whole rendezvous budget
      set_timeouts(fd, 1000);
      uint32_t hello[2] = {0, 0};
      if (recv_all(fd, hello, sizeof(hello)) != 0 || hello[0] != kHelloMagic || hello[1] == 0 || hello[1] >= world ||
          client_fd_[hello[1]] >= 0) {
        ::close(fd);
        continue;
      }
      set_timeouts(fd, timeout_ms_);
      client_fd_[hello[1]] = fd;
      ++got;
    }
  } else {
    for (;;) {
      hub_fd_ = ::socket(family, SOCK_STREAM | SOCK_CLOEXEC, 0);
      if (hub_fd_ < 0) return -errno;
      if (::connect(hub_fd_, reinterpret_cast<sockaddr*>(&addr), alen) == 0) break;
      const int e = errno;
      ::close(hub_fd_);
      hub_fd_ = -1;
      if ((e != ECONNREFUSED && e != ENOENT && e != EAGAIN) || now_ms() > t_end) {
        if (err) *err = std::string("rendezvous: connect: ") + strerror(e);
        return -e;
      }
      usleep(2000);
    }
    if (!tcp_ && !same_user(hub_fd_)) {
      if (err) *err = "rendezvous: the hub socket belongs to another user";
      return -EACCES;
    }
    set_timeouts(hub_fd_, timeout_ms_);
    const uint32_t hello[2] = {kHelloMagic, rank};
    const int rc = send_all(hub_fd_, hello, sizeof(hello));
    if (rc != 0) return rc;
  }
  return barrier(err);
}

int Rendezvous::allgather(const void* mine, size_t bytes, void* all, std::string* err) {
  if (world_ <= 1) {
    memcpy(all, mine, bytes);
    return 0;
  }
  int rc = 0;
  char* out = static_cast<char*>(all);
  if (rank_ == 0) {
    memcpy(out, mine, bytes);
    for (uint32_t r = 1; r < world_ && rc == 0; ++r) rc = recv_all(client_fd_[r], out + r * bytes, bytes);
    for (uint32_t r = 1; r < world_ && rc == 0; ++r) rc = send_all(client_fd_[r], out, bytes * world_);
  } else {
    rc = send_all(hub_fd_, mine, bytes);
    if (rc == 0) rc = recv_all(hub_fd_, out, bytes * world_);
  }
  if (rc != 0 && err) *err = std::string("rendezvous: allgather: ") + strerror(-rc);
  return rc;
}

int Rendezvous::barrier(std::string* err) {
  uint8_t token = 1;
  std::vector<uint8_t> all(world_ ? world_ : 1);
  return allgather(&token, 1, all.data(), err);
}

int Rendezvous::allgather_fds(const int* mine, uint32_t k, std::vector<int>* all, std::string* err) {
  all->assign((size_t)world_ * k, -1);
  int rc = 0;
  if (tcp_ && world_ > 1) {
    if (err) *err = "rendezvous: file descriptors cannot cross nodes (tcp: transport needs fabric handles)";
    return -ENOTSUP;
  }
  if (world_ <= 1) {
    for (uint32_t i = 0; i < k; ++i) (*all)[i] = fcntl(mine[i], F_DUPFD_CLOEXEC, 0);
    return 0;
  }
  if (rank_ == 0) {
    for (uint32_t i = 0; i < k; ++i) (*all)[i] = fcntl(mine[i], F_DUPFD_CLOEXEC, 0);
    for (uint32_t r = 1; r < world_ && rc == 0; ++r) rc = recv_fds(client_fd_[r], all->data() + (size_t)r * k, k);
    for (uint32_t r = 1; r < world_ && rc == 0; ++r) rc = send_fds(client_fd_[r], all->data(), world_ * k);
  } else {
    rc = send_fds(hub_fd_, mine, k);
    if (rc == 0) rc = recv_fds(hub_fd_, all->data(), world_ * k);
  }
  if (rc != 0) {
    for (int& fd : *all)
      if (fd >= 0) {
        ::close(fd);
        fd = -1;
      }
    if (err) *err = std::string("rendezvous: fd exchange: ") + strerror(-rc);
  }
  return rc;
}

}  // namespace cdp

// Exercises the whole exchange without CUDA: every rank shares a memfd holding
// a rank-specific pattern and checks what it receives from every other rank.
extern "C" int cdprobe_rendezvous_selftest(const char* session, uint32_t rank, uint32_t world, uint32_t timeout_ms) {
  if (session == nullptr || world == 0 || rank >= world) return CDPROBE_ERR_ARG;
  cdp::Rendezvous rdv;
  std::string err;
  if (rdv.connect(session, rank, world, timeout_ms, &err) != 0) return CDPROBE_ERR_RENDEZVOUS;
  if (rdv.is_tcp()) {
    // cross-node transport: blobs only (what carries CUmemFabricHandle), plus the barrier
    uint64_t blob[8], got[8 * CDPROBE_MAX_GPUS * 4];
    if (world > CDPROBE_MAX_GPUS * 4) return CDPROBE_ERR_ARG;
    for (int i = 0; i < 8; ++i) blob[i] = 0xFAB51C0000000000ull + ((uint64_t)rank << 8) + (uint64_t)i;
    if (rdv.allgather(blob, sizeof(blob), got, &err) != 0) return CDPROBE_ERR_RENDEZVOUS;
    for (uint32_t r = 0; r < world; ++r)
      for (int i = 0; i < 8; ++i)
        if (got[r * 8 + i] != 0xFAB51C0000000000ull + ((uint64_t)r << 8) + (uint64_t)i) return CDPROBE_ERR_INTEGRITY;
    int dummy = 0;
    std::vector<int> none;
    if (world > 1 && rdv.allgather_fds(&dummy, 1, &none, &err) != -ENOTSUP) return CDPROBE_ERR_INTEGRITY;
    return rdv.barrier(&err) == 0 ? CDPROBE_OK : CDPROBE_ERR_RENDEZVOUS;
  }
  int fd = memfd_create("cdprobe-selftest", MFD_CLOEXEC);
  if (fd < 0) return CDPROBE_ERR_RENDEZVOUS;
  uint64_t words[64];
  for (int i = 0; i < 64; ++i) words[i] = 0xC0FFEE0000000000ull + ((uint64_t)rank << 16) + (uint64_t)i;
  if (write(fd, words, sizeof(words)) != (ssize_t)sizeof(words)) {
    close(fd);
    return CDPROBE_ERR_RENDEZVOUS;
  }
  std::vector<int> all;
  int rc = rdv.allgather_fds(&fd, 1, &all, &err);
  close(fd);
  if (rc != 0) return CDPROBE_ERR_RENDEZVOUS;
  int bad = 0;
  for (uint32_t r = 0; r < world; ++r) {
    uint64_t got[64];
    if (pread(all[r], got, sizeof(got), 0) != (ssize_t)sizeof(got)) bad++;
    else
      for (int i = 0; i < 64; ++i)
        if (got[i] != 0xC0FFEE0000000000ull + ((uint64_t)r << 16) + (uint64_t)i) {
          bad++;
          break;
        }
    close(all[r]);
  }
  uint32_t mine = (uint32_t)bad, sum[CDPROBE_MAX_GPUS * 4];
  if (world > CDPROBE_MAX_GPUS * 4) return CDPROBE_ERR_ARG;
  if (rdv.allgather(&mine, sizeof(mine), sum, &err) != 0) return CDPROBE_ERR_RENDEZVOUS;
  for (uint32_t r = 0; r < world; ++r)
    if (sum[r] != 0) return CDPROBE_ERR_INTEGRITY;
  return rdv.barrier(&err) == 0 ? CDPROBE_OK : CDPROBE_ERR_RENDEZVOUS;
}
