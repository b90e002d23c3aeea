// halo.cu -- multi-rank plumbing: ghost exchange and scalar all-reduce over NCCL (NVLink 5).
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
};

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
  if (H.halo->comm) nccl().CommDestroy(H.halo->comm);
  delete H.halo;
  H.halo = nullptr;
}

void halo_set_plan(Handle &H, int n_nb, const int *nb_rank, const int *send_node_ptr, const int *send_node_idx,
                   const int *recv_node_cnt, const int *send_p_ptr, const int *send_p_idx, const int *recv_p_cnt)
{
  Halo &h = *H.halo;
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
  exchange(H, H.dim, h.send_node_ptr, h.recv_node_ptr, h.d_send_node_idx.p, h.sendbuf_u.p, x_u,
           x_u + size_t(H.dim) * H.n_nodes_owned + goff_u);
}

void halo_exchange_p(Handle &H, double *x_p, int goff_p)
{
  if (H.nranks <= 1 || !H.halo || H.halo->n_nb == 0) return;
  Halo &h = *H.halo;
  exchange(H, 1, h.send_p_ptr, h.recv_p_ptr, h.d_send_p_idx.p, h.sendbuf_p.p, x_p, x_p + H.n_p_owned + goff_p);
}

void halo_allreduce(Handle &H, double *dev, int n)
{
  if (H.nranks <= 1) return;
  NSB_NCCL(nccl().AllReduce(dev, dev, size_t(n), ncclDouble, ncclSum, H.halo->comm, H.stream));
}

} // namespace nsb
