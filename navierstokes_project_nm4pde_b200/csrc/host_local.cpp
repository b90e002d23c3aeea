// host_local.cpp -- subdomain description of one rank, computed on the host without communication.
//
// Replaces what the reference gets from GridTools::partition_triangulation +
// parallel::fullydistributed::Triangulation + the Epetra maps (Navier-Stokes/src/NavierStokes2D.cpp:16-19,
// 71-87): every rank holds the whole mesh (as in the reference, :8-14), cells are partitioned, a DoF
// belongs to the lowest rank among its cells (deal.II's rule), and a rank keeps its owned DoFs plus the
// ghosts of a TWO-layer cell halo so that every row it needs is assembled redundantly.
// Because the mesh and the partition are replicated, a rank can also work out what every OTHER rank
// needs from it: the exchange plan of nsb_set_halo comes out of this file with no message exchanged
// (navierstokes_project_nm4pde_b200/distributed.py builds the same plan with an all-to-all; the two are
// compared in tests/test_host_cpu.py).
#include <algorithm>
#include <cstring>
#include <numeric>

#include "../../include/nsb.h"
#include "nsb_host.hpp"

namespace nsb {

struct LocalProblem {
  int dim = 0, nranks = 1, rank = 0;
  std::vector<int> cells;        // global ids of the local cells (ascending)
  std::vector<int> cell_part;    // [n_cells global] owner rank of every cell
  std::vector<int> cell_dofs;    // [n_local_cells][dpc] reference layout, local numbering
  std::vector<double> cell_coords;
  std::vector<int> node_gid, p_gid; // local -> global, owned first, then ghosts by (owner, global id)
  int n_nodes_owned = 0, n_p_owned = 0;
  std::vector<int> g2l_node, g2l_p; // global -> local (-1: not local)
  std::vector<int> g2l_cell;
  // exchange plan (arrays of nsb_set_halo)
  std::vector<int> nb_ranks, send_node_ptr, send_node_idx, recv_node_cnt, send_p_ptr, send_p_idx, recv_p_cnt;
};

// DoF owner = lowest rank among the cells that hold it
static void owners(const Dofs &D, const std::vector<int> &part, int nranks, std::vector<int> &node_owner, std::vector<int> &p_owner)
{
  node_owner.assign(D.n_nodes, nranks);
  p_owner.assign(D.n_p, nranks);
  for (int64_t c = 0; c < D.nc; ++c) {
    for (int a = 0; a < D.n2; ++a) { int &o = node_owner[D.cell_nodes[c * D.n2 + a]]; o = std::min(o, part[c]); }
    for (int v = 0; v < D.nv1; ++v) { int &o = p_owner[D.cell_p[c * D.nv1 + v]]; o = std::min(o, part[c]); }
  }
}

// cells of rank r: layer 1 = cells touching a DoF owned by r, layer 2 = cells touching any DoF of layer 1
static void local_cells(const Dofs &D, const std::vector<int> &node_owner, const std::vector<int> &p_owner, int r,
                        std::vector<char> &in, std::vector<char> &mark_n, std::vector<char> &mark_p)
{
  in.assign(size_t(D.nc), 0);
  mark_n.assign(D.n_nodes, 0);
  mark_p.assign(D.n_p, 0);
  for (int64_t c = 0; c < D.nc; ++c) {
    bool touch = false;
    for (int a = 0; a < D.n2 && !touch; ++a) touch = node_owner[D.cell_nodes[c * D.n2 + a]] == r;
    for (int v = 0; v < D.nv1 && !touch; ++v) touch = p_owner[D.cell_p[c * D.nv1 + v]] == r;
    if (!touch) continue;
    for (int a = 0; a < D.n2; ++a) mark_n[D.cell_nodes[c * D.n2 + a]] = 1;
    for (int v = 0; v < D.nv1; ++v) mark_p[D.cell_p[c * D.nv1 + v]] = 1;
  }
  for (int64_t c = 0; c < D.nc; ++c) {
    bool touch = false;
    for (int a = 0; a < D.n2 && !touch; ++a) touch = mark_n[D.cell_nodes[c * D.n2 + a]];
    for (int v = 0; v < D.nv1 && !touch; ++v) touch = mark_p[D.cell_p[c * D.nv1 + v]];
    in[c] = touch;
  }
}

// the DoFs rank r uses (its local cells' DoFs + everything it owns), split into owned and ghosts sorted by (owner, id)
static void used_dofs(const Dofs &D, const std::vector<char> &in, const std::vector<int> &owner, bool nodes, int r,
                      std::vector<int> &own, std::vector<int> &ghost)
{
  const int n = nodes ? D.n_nodes : D.n_p, k = nodes ? D.n2 : D.nv1;
  const std::vector<int> &ids = nodes ? D.cell_nodes : D.cell_p;
  std::vector<char> used(n, 0);
  for (int64_t c = 0; c < D.nc; ++c)
    if (in[c])
      for (int a = 0; a < k; ++a) used[ids[c * k + a]] = 1;
  own.clear(); ghost.clear();
  for (int i = 0; i < n; ++i) {
    if (owner[i] == r) own.push_back(i); // every owned DoF is numbered even if no local cell uses it
    else if (used[i]) ghost.push_back(i);
  }
  std::stable_sort(ghost.begin(), ghost.end(), [&](int a, int b) { return owner[a] < owner[b]; }); // ids ascending inside an owner
}

void build_local_problem(const Mesh &M, const Dofs &D, int nranks, int rank, LocalProblem &L)
{
  const int dim = D.dim, n2 = D.n2, nv1 = D.nv1, dpc = D.dpc;
  L.dim = dim; L.nranks = nranks; L.rank = rank;
  partition_cells_rcb(M, nranks, L.cell_part);
  std::vector<int> node_owner, p_owner;
  owners(D, L.cell_part, nranks, node_owner, p_owner);
  std::vector<char> in, mn, mp;
  local_cells(D, node_owner, p_owner, rank, in, mn, mp);
  L.cells.clear();
  for (int64_t c = 0; c < D.nc; ++c)
    if (in[c]) L.cells.push_back(int(c));
  std::vector<int> own_n, gh_n, own_p, gh_p;
  used_dofs(D, in, node_owner, true, rank, own_n, gh_n);
  used_dofs(D, in, p_owner, false, rank, own_p, gh_p);
  L.n_nodes_owned = int(own_n.size());
  L.n_p_owned = int(own_p.size());
  L.node_gid = own_n; L.node_gid.insert(L.node_gid.end(), gh_n.begin(), gh_n.end());
  L.p_gid = own_p; L.p_gid.insert(L.p_gid.end(), gh_p.begin(), gh_p.end());
  L.g2l_node.assign(D.n_nodes, -1);
  L.g2l_p.assign(D.n_p, -1);
  for (size_t i = 0; i < L.node_gid.size(); ++i) L.g2l_node[L.node_gid[i]] = int(i);
  for (size_t i = 0; i < L.p_gid.size(); ++i) L.g2l_p[L.p_gid[i]] = int(i);
  L.g2l_cell.assign(size_t(D.nc), -1);
  const int nlc = int(L.cells.size()), n_u_loc = dim * int(L.node_gid.size());
  L.cell_dofs.assign(size_t(nlc) * dpc, 0);
  L.cell_coords.assign(size_t(nlc) * nv1 * dim, 0.0);
  for (int k = 0; k < nlc; ++k) {
    const int64_t c = L.cells[k];
    L.g2l_cell[c] = k;
    int *cd = &L.cell_dofs[size_t(k) * dpc];
    for (int a = 0; a < n2; ++a) {
      const int base = a < nv1 ? a * (dim + 1) : nv1 * (dim + 1) + (a - nv1) * dim;
      const int ln = L.g2l_node[D.cell_nodes[c * n2 + a]];
      for (int d = 0; d < dim; ++d) cd[base + d] = dim * ln + d;
    }
    for (int v = 0; v < nv1; ++v) cd[v * (dim + 1) + dim] = n_u_loc + L.g2l_p[D.cell_p[c * nv1 + v]];
    std::memcpy(&L.cell_coords[size_t(k) * nv1 * dim], &D.cell_coords[size_t(c) * nv1 * dim], sizeof(double) * nv1 * dim);
  }
  // ---- exchange plan.  What I receive: my ghosts, grouped by owner (that IS the ghost order).  What I send to
  // rank q: the DoFs I own among q's ghosts, in q's ghost order (ascending global id) -- q's ghost set is
  // recomputed here from the replicated mesh.
  std::vector<std::vector<int>> send_n(nranks), send_p(nranks);
  std::vector<int> recv_n(nranks, 0), recv_p(nranks, 0);
  for (int g : gh_n) recv_n[node_owner[g]]++;
  for (int g : gh_p) recv_p[p_owner[g]]++;
  for (int q = 0; q < nranks; ++q) {
    if (q == rank) continue;
    std::vector<char> inq, a, b;
    local_cells(D, node_owner, p_owner, q, inq, a, b);
    std::vector<char> un(D.n_nodes, 0), up(D.n_p, 0);
    for (int64_t c = 0; c < D.nc; ++c)
      if (inq[c]) {
        for (int t = 0; t < n2; ++t) un[D.cell_nodes[c * n2 + t]] = 1;
        for (int v = 0; v < nv1; ++v) up[D.cell_p[c * nv1 + v]] = 1;
      }
    for (int i = 0; i < D.n_nodes; ++i)
      if (un[i] && node_owner[i] == rank) send_n[q].push_back(L.g2l_node[i]);
    for (int i = 0; i < D.n_p; ++i)
      if (up[i] && p_owner[i] == rank) send_p[q].push_back(L.g2l_p[i]);
  }
  L.nb_ranks.clear();
  L.send_node_ptr.assign(1, 0); L.send_p_ptr.assign(1, 0);
  L.send_node_idx.clear(); L.send_p_idx.clear(); L.recv_node_cnt.clear(); L.recv_p_cnt.clear();
  for (int q = 0; q < nranks; ++q) {
    if (q == rank) continue;
    if (send_n[q].empty() && send_p[q].empty() && recv_n[q] == 0 && recv_p[q] == 0) continue;
    L.nb_ranks.push_back(q);
    L.send_node_idx.insert(L.send_node_idx.end(), send_n[q].begin(), send_n[q].end());
    L.send_p_idx.insert(L.send_p_idx.end(), send_p[q].begin(), send_p[q].end());
    L.send_node_ptr.push_back(int(L.send_node_idx.size()));
    L.send_p_ptr.push_back(int(L.send_p_idx.size()));
    L.recv_node_cnt.push_back(recv_n[q]);
    L.recv_p_cnt.push_back(recv_p[q]);
  }
}

} // namespace nsb

// ------------------------------------------------------------------------------------------------ C ABI
struct nsh_mesh_s;
struct nsh_dofs_s;
namespace nsb {
const Mesh &mesh_of(const nsh_mesh_s *m);
const Dofs &dofs_of(const nsh_dofs_s *d);
} // namespace nsb
struct nsh_local_s { nsb::LocalProblem L; };

extern "C" {

nsh_local nsh_local_create(nsh_mesh m, nsh_dofs d, int32_t nranks, int32_t rank)
{
  if (!m || !d || nranks < 1 || rank < 0 || rank >= nranks) return nullptr;
  try {
    auto *h = new nsh_local_s();
    nsb::build_local_problem(nsb::mesh_of(m), nsb::dofs_of(d), nranks, rank, h->L);
    return h;
  } catch (...) {
    return nullptr;
  }
}
void nsh_local_free(nsh_local l) { delete l; }
int32_t nsh_local_n_cells(nsh_local l) { return int32_t(l->L.cells.size()); }
int32_t nsh_local_n_nodes(nsh_local l) { return int32_t(l->L.node_gid.size()); }
int32_t nsh_local_n_p(nsh_local l) { return int32_t(l->L.p_gid.size()); }
int32_t nsh_local_n_nodes_owned(nsh_local l) { return l->L.n_nodes_owned; }
int32_t nsh_local_n_p_owned(nsh_local l) { return l->L.n_p_owned; }
const int32_t *nsh_local_cells(nsh_local l) { return l->L.cells.data(); }
const int32_t *nsh_local_cell_part(nsh_local l) { return l->L.cell_part.data(); }
const int32_t *nsh_local_cell_dofs(nsh_local l) { return l->L.cell_dofs.data(); }
const double *nsh_local_cell_coords(nsh_local l) { return l->L.cell_coords.data(); }
const int32_t *nsh_local_node_gid(nsh_local l) { return l->L.node_gid.data(); }
const int32_t *nsh_local_p_gid(nsh_local l) { return l->L.p_gid.data(); }
const int32_t *nsh_local_g2l_node(nsh_local l) { return l->L.g2l_node.data(); }
const int32_t *nsh_local_g2l_cell(nsh_local l) { return l->L.g2l_cell.data(); }
int32_t nsh_local_halo(nsh_local l, const int32_t **nb_ranks, const int32_t **send_node_ptr, const int32_t **send_node_idx,
                       const int32_t **recv_node_cnt, const int32_t **send_p_ptr, const int32_t **send_p_idx,
                       const int32_t **recv_p_cnt)
{
  static const int32_t zero = 0;
  auto ptr = [](const std::vector<int> &v) { return v.empty() ? &zero : v.data(); };
  const nsb::LocalProblem &L = l->L;
  *nb_ranks = ptr(L.nb_ranks); *send_node_ptr = ptr(L.send_node_ptr); *send_node_idx = ptr(L.send_node_idx);
  *recv_node_cnt = ptr(L.recv_node_cnt); *send_p_ptr = ptr(L.send_p_ptr); *send_p_idx = ptr(L.send_p_idx);
  *recv_p_cnt = ptr(L.recv_p_cnt);
  return int32_t(L.nb_ranks.size());
}

} // extern "C"
