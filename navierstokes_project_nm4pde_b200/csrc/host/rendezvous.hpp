// rendezvous.hpp -- the little the three drivers need from MPI when they run one process per GPU
// (Navier-Stokes/src/main3D.cpp:9 `Utilities::MPI::MPI_InitFinalize`, :28 `this_mpi_process`): rank / size
// from the launcher's environment and an all-gather of small fixed-size blobs -- the 128-byte NCCL id and
// the 64-byte CUDA IPC handles of the peer-memory mailboxes.  Everything that moves per time step goes
// over NVLink inside libnsb.so (csrc/halo.cu); this channel is used during setup() only.
//
// Launcher contract (what `python -m torch.distributed.run --no-python --nproc-per-node N ./navier_stokes3D`
// or scripts/nsb_launch.sh export): RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR; the TCP port is
// NSB_RDV_PORT, else MASTER_PORT + 1 (MASTER_PORT itself is taken by torchrun's store), else 29617.
// Star topology through rank 0; no MPI, no third-party code.
#pragma once
#include <arpa/inet.h>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <stdexcept>
#include <string>
#include <sys/socket.h>
#include <thread>
#include <unistd.h>
#include <vector>

class Rendezvous
{
public:
  Rendezvous()
  {
    rank_ = env("RANK", 0);
    size_ = env("WORLD_SIZE", 1);
    local_rank_ = env("LOCAL_RANK", rank_);
    if (size_ < 1 || rank_ < 0 || rank_ >= size_) throw std::runtime_error("rendezvous: bad RANK / WORLD_SIZE");
    if (size_ == 1) return;
    const char *addr_env = std::getenv("MASTER_ADDR");
    const std::string addr = addr_env ? addr_env : "127.0.0.1";
    const int port = std::getenv("NSB_RDV_PORT") ? env("NSB_RDV_PORT", 0) : std::getenv("MASTER_PORT") ? env("MASTER_PORT", 0) + 1 : 29617;
    sockaddr_in sa;
    std::memset(&sa, 0, sizeof(sa));
    sa.sin_family = AF_INET;
    sa.sin_port = htons(uint16_t(port));
    if (rank_ == 0) {
      listen_fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
      if (listen_fd_ < 0) throw std::runtime_error("rendezvous: socket()");
      int one = 1;
      ::setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
      sa.sin_addr.s_addr = htonl(INADDR_ANY);
      if (::bind(listen_fd_, reinterpret_cast<sockaddr *>(&sa), sizeof(sa)) != 0 || ::listen(listen_fd_, size_) != 0)
        throw std::runtime_error("rendezvous: cannot listen on port " + std::to_string(port));
      peers_.assign(size_t(size_), -1);
      for (int k = 1; k < size_; ++k) {
        const int fd = ::accept(listen_fd_, nullptr, nullptr);
        if (fd < 0) throw std::runtime_error("rendezvous: accept()");
        nodelay(fd);
        int32_t r = -1;
        recv_all(fd, &r, sizeof(r));
        if (r < 1 || r >= size_ || peers_[size_t(r)] >= 0) throw std::runtime_error("rendezvous: unexpected peer rank");
        peers_[size_t(r)] = fd;
      }
    } else {
      if (::inet_pton(AF_INET, addr.c_str(), &sa.sin_addr) != 1) throw std::runtime_error("rendezvous: MASTER_ADDR must be an IPv4 address");
      const auto t0 = std::chrono::steady_clock::now();
      int fd = -1;
      for (;;) { // rank 0 may still be starting
        fd = ::socket(AF_INET, SOCK_STREAM, 0);
        if (fd >= 0 && ::connect(fd, reinterpret_cast<sockaddr *>(&sa), sizeof(sa)) == 0) break;
        if (fd >= 0) ::close(fd);
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 120.0)
          throw std::runtime_error("rendezvous: cannot reach rank 0 at " + addr + ":" + std::to_string(port));
        std::this_thread::sleep_for(std::chrono::milliseconds(50));
      }
      nodelay(fd);
      const int32_t r = rank_;
      send_all(fd, &r, sizeof(r));
      peers_.assign(1, fd);
    }
  }
  ~Rendezvous()
  {
    for (int fd : peers_)
      if (fd >= 0) ::close(fd);
    if (listen_fd_ >= 0) ::close(listen_fd_);
  }
  Rendezvous(const Rendezvous &) = delete;
  Rendezvous &operator=(const Rendezvous &) = delete;

  int rank() const { return rank_; }
  int size() const { return size_; }
  int local_rank() const { return local_rank_; }

  // all[size][bytes] in rank order on every rank (MPI_Allgather)
  void allgather(const void *mine, void *all, size_t bytes)
  {
    char *out = static_cast<char *>(all);
    if (size_ == 1) { std::memcpy(out, mine, bytes); return; }
    if (rank_ == 0) {
      std::memcpy(out, mine, bytes);
      for (int r = 1; r < size_; ++r) recv_all(peers_[size_t(r)], out + size_t(r) * bytes, bytes);
      for (int r = 1; r < size_; ++r) send_all(peers_[size_t(r)], out, bytes * size_t(size_));
    } else {
      send_all(peers_[0], mine, bytes);
      recv_all(peers_[0], out, bytes * size_t(size_));
    }
  }
  void barrier()
  {
    std::vector<char> all(static_cast<size_t>(size_));
    const char c = 0;
    allgather(&c, all.data(), 1);
  }

private:
  static int env(const char *name, int def) { const char *e = std::getenv(name); return e ? std::atoi(e) : def; }
  static void nodelay(int fd) { int one = 1; ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one)); }
  static void send_all(int fd, const void *buf, size_t n)
  {
    const char *p = static_cast<const char *>(buf);
    while (n) {
      const ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
      if (k <= 0) throw std::runtime_error("rendezvous: peer closed the connection (send)");
      p += k; n -= size_t(k);
    }
  }
  static void recv_all(int fd, void *buf, size_t n)
  {
    char *p = static_cast<char *>(buf);
    while (n) {
      const ssize_t k = ::recv(fd, p, n, 0);
      if (k <= 0) throw std::runtime_error("rendezvous: peer closed the connection (recv)");
      p += k; n -= size_t(k);
    }
  }
  int rank_ = 0, size_ = 1, local_rank_ = 0, listen_fd_ = -1;
  std::vector<int> peers_;
};
