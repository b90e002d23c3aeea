// halo.cu -- multi-rank plumbing: ghost exchange and scalar all-reduce over NVLink 5.
//
// Two transports.  The product path is PEER MEMORY: every rank owns a "mailbox" (cudaMalloc, exported
// with cudaIpcGetMemHandle, mapped by all peers), and the kernels of this library store straight into
// the peers' mailboxes over NVLink / NVSwitch:
//   * ghost exchange = k_halo_push (gather + P2P stores + one release flag per neighbour) followed by
//     k_halo_wait (acquire spin on the local flags, copy the inbox into the ghost segment);
//   * all-reduce of dot products = inside the LAST BLOCK of the reduction kernel that produced the
//     partial result (kernels_linalg.cu: finish_reduce / k_multi_dot): it stores its value into every
//     peer's slot, raises the flags, waits for the peers' flags and adds the nranks partials in rank
//     order -- compute and collective are one kernel, and every rank gets bitwise the same sum.
// Mailbox buffers and flags are double-buffered by the parity of a per-kind sequence number (the SPMD
// solver issues the same sequence of exchanges on every rank), which is enough because a rank can
// only be one exchange ahead of a neighbour whose previous push it has consumed.
// NCCL (grouped send/recv, ncclAllReduce) is the fallback transport when peer mapping is unavailable
// and the bootstrap for nothing else.
//
// Replaces the Epetra_Import done before every distributed SpMV and the MPI_Allreduce behind
// every dot product / norm of the reference.  One communicator per handle,
// everything is enqueued on the handle's stream so no host synchronisation is added.
// NCCL is bound at run time (dlopen) so that the library also loads on machines without it and
// so that, inside a torch process, the already loaded libnccl is shared instead of duplicated.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>

#include "nsb_internal.hpp"

namespace nsb {

struct NcclApi {
  void *lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

static NcclApi &nccl()
{
  static NcclApi api;
  if (api.lib) return api;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (api.lib) break;
  }
  if (!api.lib) throw NcclError(std::string("cannot load libnccl: ") + dlerror());
#define NSB_SYM(field, name)                                                  \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name));    \
  if (!api.field) throw NcclError(std::string("libnccl lacks ") + name);
  NSB_SYM(GetUniqueId, "ncclGetUniqueId")
  NSB_SYM(CommInitRank, "ncclCommInitRank")
  NSB_SYM(CommDestroy, "ncclCommDestroy")
  NSB_SYM(AllReduce, "ncclAllReduce")
  NSB_SYM(Send, "ncclSend")
  NSB_SYM(Recv, "ncclRecv")
  NSB_SYM(GroupStart, "ncclGroupStart")
  NSB_SYM(GroupEnd, "ncclGroupEnd")
  NSB_SYM(GetErrorString, "ncclGetErrorString")
#undef NSB_SYM
  return api;
}

#define NSB_NCCL(call)                                                                       \
  do {                                                                                       \
    ncclResult_t r_ = (call);                                                                \
    if (r_ != ncclSuccess)                                                                   \
      throw nsb::NcclError(std::string(#call) + ": " + nccl().GetErrorString(r_));           \
  } while (0)

struct Halo {
  ncclComm_t comm = nullptr;
  int n_nb = 0;
  std::vector<int> nb_rank;
  std::vector<int> send_node_ptr, recv_node_ptr, send_p_ptr, recv_p_ptr;
  DevBuf<int> d_send_node_idx, d_send_p_idx;
  DevBuf<double> sendbuf_u, sendbuf_p;
  // ---- peer-memory transport
  bool p2p = false;
  void *mailbox = nullptr;            // this rank's mailbox (MailHdr + inboxes)
  size_t mailbox_bytes = 0;
  std::vector<void *> peer_base;      // mapped mailboxes, [nranks]; own entry = mailbox
  DevBuf<PeerTab> d_tab;
  DevBuf<unsigned> d_ticket;
  unsigned long long seq_u = 0, seq_p = 0, seq_ar = 0;
};

static inline size_t mailbox_size(long long nvals_u, long long nvals_p)
{
  return sizeof(MailHdr) + sizeof(double) * 2 * size_t(nvals_u + nvals_p);
}
static inline double *inbox_u(void *base, long long, int parity, long long nvals_u)
{
  return reinterpret_cast<double *>(static_cast<char *>(base) + sizeof(MailHdr)) + size_t(parity) * nvals_u;
}
static inline double *inbox_p(void *base, long long nvals_u, int parity, long long nvals_p)
{
  return reinterpret_cast<double *>(static_cast<char *>(base) + sizeof(MailHdr)) + 2 * size_t(nvals_u) + size_t(parity) * nvals_p;
}

void get_unique_id(void *out128)
{
  ncclUniqueId id;
  NSB_NCCL(nccl().GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  std::memcpy(out128, &id, 128);
}

void halo_create(Handle &H, const void *unique_id)
{
  H.halo = new Halo();
  if (H.nranks <= 1) return;
  if (!unique_id) throw ArgError("nsb_create: nranks > 1 needs an NCCL unique id");
  ncclUniqueId id;
  std::memcpy(&id, unique_id, 128);
  NSB_NCCL(nccl().CommInitRank(&H.halo->comm, H.nranks, id, H.rank));
}

void halo_destroy(Handle &H)
{
  if (!H.halo) return;
  Halo &h = *H.halo;
  for (size_t r = 0; r < h.peer_base.size(); ++r)
    if (h.peer_base[r] && int(r) != H.rank) cudaIpcCloseMemHandle(h.peer_base[r]);
  if (h.mailbox) cudaFree(h.mailbox);
  if (h.comm) nccl().CommDestroy(h.comm);
  delete H.halo;
  H.halo = nullptr;
}

void halo_set_plan(Handle &H, int n_nb, const int *nb_rank, const int *send_node_ptr, const int *send_node_idx,
                   const int *recv_node_cnt, const int *send_p_ptr, const int *send_p_idx, const int *recv_p_cnt)
{
  Halo &h = *H.halo;
  h.p2p = false; // a new plan: back to NCCL until nsb_p2p_export / nsb_p2p_attach are called again
  h.n_nb = n_nb;
  h.nb_rank.assign(nb_rank, nb_rank + n_nb);
  h.send_node_ptr.assign(send_node_ptr, send_node_ptr + n_nb + 1);
  h.send_p_ptr.assign(send_p_ptr, send_p_ptr + n_nb + 1);
  h.recv_node_ptr.assign(n_nb + 1, 0);
  h.recv_p_ptr.assign(n_nb + 1, 0);
  for (int k = 0; k < n_nb; ++k) {
    h.recv_node_ptr[k + 1] = h.recv_node_ptr[k] + recv_node_cnt[k];
    h.recv_p_ptr[k + 1] = h.recv_p_ptr[k] + recv_p_cnt[k];
  }
  if (h.recv_node_ptr[n_nb] != H.n_nodes - H.n_nodes_owned || h.recv_p_ptr[n_nb] != H.n_p - H.n_p_owned)
    throw ArgError("nsb_set_halo: receive counts do not add up to the number of ghosts");
  std::vector<int> sn(send_node_idx, send_node_idx + h.send_node_ptr[n_nb]);
  std::vector<int> spx(send_p_idx, send_p_idx + h.send_p_ptr[n_nb]);
  for (int v : sn) if (v < 0 || v >= H.n_nodes_owned) throw ArgError("nsb_set_halo: send node index out of range");
  for (int v : spx) if (v < 0 || v >= H.n_p_owned) throw ArgError("nsb_set_halo: send pressure index out of range");
  h.d_send_node_idx.upload(sn);
  h.d_send_p_idx.upload(spx);
  h.sendbuf_u.alloc(std::max<size_t>(1, sn.size() * H.dim));
  h.sendbuf_p.alloc(std::max<size_t>(1, spx.size()));
}

// ---------------------------------------------------------------------------------------------
// peer-memory transport
// ---------------------------------------------------------------------------------------------
template <int BS>
__global__ void k_halo_push(const PeerTab *__restrict__ T, int which, int n_send, const int *__restrict__ idx,
                            const double *__restrict__ x, int parity, unsigned long long seq, unsigned *ticket)
{
  const int *sptr = which == 0 ? T->send_ptr_u : T->send_ptr_p;
  double *const *dst = which == 0 ? T->dst_u : T->dst_p;
  const long long *stride = which == 0 ? T->stride_u : T->stride_p;
  const int n_nb = T->n_nb;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_send * BS; t += gridDim.x * blockDim.x) {
    const int e = t / BS, c = t - e * BS;
    int k = 0;
    while (k + 1 < n_nb && e >= sptr[k + 1]) ++k;
    dst[k][size_t(parity) * stride[k] + size_t(e - sptr[k]) * BS + c] = x[int64_t(BS) * idx[e] + c];
  }
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if (int(threadIdx.x) < n_nb) {
    MailHdr *peer = T->peer[T->nb_rank[threadIdx.x]];
    unsigned long long *f = which == 0 ? &peer->hu_flag[parity][T->me] : &peer->hp_flag[parity][T->me];
    st_release_sys(f, seq);
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

__global__ void k_halo_wait(const PeerTab *__restrict__ T, int which, int parity, unsigned long long seq, int nvals,
                            const double *inbox, double *__restrict__ x_ghost)
{
  if (int(threadIdx.x) < T->n_nb) {
    const MailHdr *mine = T->peer[T->me];
    const int r = T->nb_rank[threadIdx.x];
    wait_flag(which == 0 ? &mine->hu_flag[parity][r] : &mine->hp_flag[parity][r], seq);
  }
  __syncthreads();
  // the inbox was written by remote stores that land in this GPU's L2: bypass the (incoherent) L1
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nvals; i += gridDim.x * blockDim.x) x_ghost[i] = __ldcg(inbox + i);
}

// stand-alone all-reduce of n <= kArSlots device doubles (used where the producing kernel is not one
// of the fused reductions)
__global__ void k_allreduce_p2p(const PeerTab *__restrict__ T, double *dev, int n, int parity, unsigned long long seq)
{
  const int j = threadIdx.x;
  const double v = j < n ? dev[j] : 0.0;
  const double s = ar_exchange_block(T, parity, seq, v, j, n);
  if (j < n) dev[j] = s;
}

// Allocates this rank's mailbox, writes its directory (where each neighbour's ghosts go) and returns
// the IPC handle.  Needs the halo plan (nsb_set_halo).
void halo_p2p_export(Handle &H, void *handle64)
{
  if (H.nranks <= 1 || !H.halo) throw StateError("nsb_p2p_export: single-rank handle");
  if (H.nranks > kMaxRanks) throw StateError("nsb_p2p_export: more ranks than mailbox slots");
  Halo &h = *H.halo;
  if (int(h.recv_node_ptr.size()) != h.n_nb + 1) throw StateError("nsb_p2p_export before nsb_set_halo");
  const long long nvals_u = (long long)(H.dim) * (H.n_nodes - H.n_nodes_owned), nvals_p = H.n_p - H.n_p_owned;
  if (h.p2p || (h.mailbox && h.mailbox_bytes != mailbox_size(nvals_u, nvals_p))) {
    // a new mesh / halo plan: drop the old mapping (peers hold stale offsets), start the sequences again
    for (size_t r = 0; r < h.peer_base.size(); ++r)
      if (h.peer_base[r] && int(r) != H.rank) cudaIpcCloseMemHandle(h.peer_base[r]);
    h.peer_base.clear();
    if (h.mailbox) cudaFree(h.mailbox);
    h.mailbox = nullptr;
    h.p2p = false;
    h.seq_u = h.seq_p = h.seq_ar = 0;
  }
  if (!h.mailbox) {
    h.mailbox_bytes = mailbox_size(nvals_u, nvals_p);
    NSB_CUDA(cudaMalloc(&h.mailbox, h.mailbox_bytes));
    NSB_CUDA(cudaMemset(h.mailbox, 0, h.mailbox_bytes));
  }
  MailHdr hdr;
  std::memset(&hdr, 0, sizeof(hdr));
  for (int r = 0; r < kMaxRanks; ++r) hdr.dir_u[r] = hdr.dir_p[r] = -1;
  for (int k = 0; k < h.n_nb; ++k) {
    hdr.dir_u[h.nb_rank[k]] = (long long)(H.dim) * h.recv_node_ptr[k];
    hdr.dir_p[h.nb_rank[k]] = h.recv_p_ptr[k];
  }
  hdr.nvals_u = nvals_u;
  hdr.nvals_p = nvals_p;
  NSB_CUDA(cudaMemcpy(h.mailbox, &hdr, sizeof(hdr), cudaMemcpyHostToDevice));
  NSB_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t ih;
  NSB_CUDA(cudaIpcGetMemHandle(&ih, h.mailbox));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  std::memcpy(handle64, &ih, 64);
}

// Maps the peers' mailboxes (handles: nranks x 64 bytes in rank order, all exported before this call)
// and switches the transport to peer memory.
void halo_p2p_attach(Handle &H, const void *handles)
{
  if (H.nranks <= 1 || !H.halo || !H.halo->mailbox) throw StateError("nsb_p2p_attach before nsb_p2p_export");
  Halo &h = *H.halo;
  h.peer_base.assign(H.nranks, nullptr);
  PeerTab tab;
  std::memset(&tab, 0, sizeof(tab));
  tab.nranks = H.nranks; tab.me = H.rank; tab.n_nb = h.n_nb;
  for (int r = 0; r < H.nranks; ++r) {
    if (r == H.rank) h.peer_base[r] = h.mailbox;
    else {
      cudaIpcMemHandle_t ih;
      std::memcpy(&ih, static_cast<const char *>(handles) + size_t(64) * r, 64);
      NSB_CUDA(cudaIpcOpenMemHandle(&h.peer_base[r], ih, cudaIpcMemLazyEnablePeerAccess));
    }
    tab.peer[r] = static_cast<MailHdr *>(h.peer_base[r]);
  }
  for (int k = 0; k < h.n_nb; ++k) {
    const int r = h.nb_rank[k];
    MailHdr ph; // the neighbour's directory: where my data goes in its inboxes
    NSB_CUDA(cudaMemcpy(&ph, h.peer_base[r], sizeof(ph), cudaMemcpyDeviceToHost));
    const int sc_u = h.send_node_ptr[k + 1] - h.send_node_ptr[k], sc_p = h.send_p_ptr[k + 1] - h.send_p_ptr[k];
    if ((sc_u > 0 && ph.dir_u[H.rank] < 0) || (sc_p > 0 && ph.dir_p[H.rank] < 0))
      throw StateError("nsb_p2p_attach: neighbour does not expect data from this rank");
    tab.nb_rank[k] = r;
    tab.dst_u[k] = inbox_u(h.peer_base[r], 0, 0, ph.nvals_u) + std::max(0LL, ph.dir_u[H.rank]);
    tab.stride_u[k] = ph.nvals_u;
    tab.dst_p[k] = inbox_p(h.peer_base[r], ph.nvals_u, 0, ph.nvals_p) + std::max(0LL, ph.dir_p[H.rank]);
    tab.stride_p[k] = ph.nvals_p;
  }
  for (int k = 0; k <= h.n_nb; ++k) { tab.send_ptr_u[k] = h.send_node_ptr[k]; tab.send_ptr_p[k] = h.send_p_ptr[k]; }
  h.d_tab.upload(std::vector<PeerTab>(1, tab));
  h.d_ticket.alloc(4);
  h.d_ticket.zero();
  NSB_CUDA(cudaDeviceSynchronize());
  h.p2p = true;
}

bool halo_is_p2p(const Handle &H) { return H.halo && H.halo->p2p; }

// arguments for a reduction kernel whose last block performs the all-reduce itself
ArArgs halo_ar_args(Handle &H)
{
  ArArgs a{nullptr, 0, 0};
  if (H.nranks > 1 && H.halo && H.halo->p2p) {
    Halo &h = *H.halo;
    ++h.seq_ar;
    a.tab = h.d_tab.p; a.parity = int(h.seq_ar & 1); a.seq = h.seq_ar;
  }
  return a;
}

static void exchange_p2p(Handle &H, int which, int bs, const std::vector<int> &sptr, const int *d_idx, const double *x_owned,
                         double *x_ghost)
{
  Halo &h = *H.halo;
  unsigned long long &seq = which == 0 ? h.seq_u : h.seq_p;
  ++seq;
  const int parity = int(seq & 1);
  const int ns = sptr[h.n_nb];
  const long long nvals_u = (long long)(H.dim) * (H.n_nodes - H.n_nodes_owned), nvals_p = H.n_p - H.n_p_owned;
  {
    const unsigned grid = unsigned(std::max(1, std::min((ns * bs + 255) / 256, 148 * 2)));
    unsigned *ticket = h.d_ticket.p + which;
    if (bs == 1) k_halo_push<1><<<grid, 256, 0, H.stream>>>(h.d_tab.p, which, ns, d_idx, x_owned, parity, seq, ticket);
    else if (bs == 2) k_halo_push<2><<<grid, 256, 0, H.stream>>>(h.d_tab.p, which, ns, d_idx, x_owned, parity, seq, ticket);
    else k_halo_push<3><<<grid, 256, 0, H.stream>>>(h.d_tab.p, which, ns, d_idx, x_owned, parity, seq, ticket);
  }
  const int nvals = int(which == 0 ? nvals_u : nvals_p);
  const double *inbox = which == 0 ? inbox_u(h.mailbox, 0, parity, nvals_u) : inbox_p(h.mailbox, nvals_u, parity, nvals_p);
  const unsigned grid = unsigned(std::max(1, std::min((nvals + 255) / 256, 148 * 2)));
  k_halo_wait<<<grid, 256, 0, H.stream>>>(h.d_tab.p, which, parity, seq, nvals, inbox, x_ghost);
  NSB_CUDA(cudaGetLastError());
  H.launches += 2;
}

template <int BS>
__global__ void k_pack(int n, const int *__restrict__ idx, const double *__restrict__ x, double *__restrict__ buf)
{
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n * BS; k += gridDim.x * blockDim.x)
    buf[k] = x[int64_t(BS) * idx[k / BS] + (k % BS)];
}

static void exchange(Handle &H, int bs, const std::vector<int> &sptr, const std::vector<int> &rptr, const int *d_idx,
                     double *sendbuf, const double *x_owned, double *x_ghost)
{
  Halo &h = *H.halo;
  const int ns = sptr[h.n_nb];
  if (ns > 0) {
    const unsigned grid = unsigned(std::max(1, std::min((ns * bs + 255) / 256, 148 * 4)));
    if (bs == 1) k_pack<1><<<grid, 256, 0, H.stream>>>(ns, d_idx, x_owned, sendbuf);
    else if (bs == 2) k_pack<2><<<grid, 256, 0, H.stream>>>(ns, d_idx, x_owned, sendbuf);
    else k_pack<3><<<grid, 256, 0, H.stream>>>(ns, d_idx, x_owned, sendbuf);
    H.launches++;
  }
  NSB_NCCL(nccl().GroupStart());
  for (int k = 0; k < h.n_nb; ++k) {
    const int sc = sptr[k + 1] - sptr[k], rc = rptr[k + 1] - rptr[k];
    if (sc > 0)
      NSB_NCCL(nccl().Send(sendbuf + size_t(bs) * sptr[k], size_t(sc) * bs, ncclDouble, h.nb_rank[k], h.comm, H.stream));
    if (rc > 0)
      NSB_NCCL(nccl().Recv(x_ghost + size_t(bs) * rptr[k], size_t(rc) * bs, ncclDouble, h.nb_rank[k], h.comm, H.stream));
  }
  NSB_NCCL(nccl().GroupEnd());
}

void halo_exchange_u(Handle &H, double *x_u, int goff_u)
{
  if (H.nranks <= 1 || !H.halo || H.halo->n_nb == 0) return;
  Halo &h = *H.halo;
  if (h.p2p) {
    exchange_p2p(H, 0, H.dim, h.send_node_ptr, h.d_send_node_idx.p, x_u, x_u + size_t(H.dim) * H.n_nodes_owned + goff_u);
    return;
  }
  exchange(H, H.dim, h.send_node_ptr, h.recv_node_ptr, h.d_send_node_idx.p, h.sendbuf_u.p, x_u,
           x_u + size_t(H.dim) * H.n_nodes_owned + goff_u);
}

void halo_exchange_p(Handle &H, double *x_p, int goff_p)
{
  if (H.nranks <= 1 || !H.halo || H.halo->n_nb == 0) return;
  Halo &h = *H.halo;
  if (h.p2p) {
    exchange_p2p(H, 1, 1, h.send_p_ptr, h.d_send_p_idx.p, x_p, x_p + H.n_p_owned + goff_p);
    return;
  }
  exchange(H, 1, h.send_p_ptr, h.recv_p_ptr, h.d_send_p_idx.p, h.sendbuf_p.p, x_p, x_p + H.n_p_owned + goff_p);
}

void halo_allreduce(Handle &H, double *dev, int n)
{
  if (H.nranks <= 1) return;
  if (H.halo->p2p) {
    for (int j0 = 0; j0 < n; j0 += kArSlots) {
      const ArArgs a = halo_ar_args(H);
      k_allreduce_p2p<<<1, kArSlots, 0, H.stream>>>(a.tab, dev + j0, std::min(kArSlots, n - j0), a.parity, a.seq);
      H.launches++;
    }
    NSB_CUDA(cudaGetLastError());
    return;
  }
  NSB_NCCL(nccl().AllReduce(dev, dev, size_t(n), ncclDouble, ncclSum, H.halo->comm, H.stream));
}

} // namespace nsb
