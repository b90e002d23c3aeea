// host_dofs.cpp -- DoF numbering, sparsity patterns and partitioning (cold path, host only).
//
// Restates what the reference's setup() gets from deal.II
// (Navier-Stokes/src/NavierStokes2D.cpp:58-156):
//   * dof_handler.distribute_dofs(FESystem(P2^dim, P1)): cells in mesh order; on each cell first
//     the not-yet-numbered vertices ([u_0..u_{dim-1}, p] each), then the not-yet-numbered edges
//     ([u_0..u_{dim-1}] each);
//   * DoFRenumbering::component_wise with block_component = {0,..,0,1}: velocity DoFs first,
//     pressure after, relative order kept.
// The result is stored compactly: P2 node ids (velocity DoF = dim*node + c) and pressure ids.
//   * DoFTools::make_sparsity_pattern with the coupling table of :109-119 is the union over
//     cells of all local pairs except pressure-pressure; in compact form that is the node-node
//     graph F_s, the pressure-node graph B and its transpose.
#include "nsb_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include <sys/mman.h>

namespace nsb {

void prefault_parallel(void *p, size_t bytes)
{
#ifdef MADV_POPULATE_WRITE
  const uintptr_t page = 4096, chunk = uintptr_t(32) << 20;
  const uintptr_t a = (reinterpret_cast<uintptr_t>(p) + page - 1) & ~(page - 1);
  const uintptr_t e = (reinterpret_cast<uintptr_t>(p) + bytes) & ~(page - 1);
  if (!p || bytes < (size_t(64) << 20) || e <= a) return;
  const int64_t n = int64_t((e - a + chunk - 1) / chunk);
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < n; ++i) {
    const uintptr_t b = a + uintptr_t(i) * chunk;
    (void)madvise(reinterpret_cast<void *>(b), size_t(std::min(chunk, e - b)), MADV_POPULATE_WRITE); // a hint: failure is fine
  }
#else
  (void)p; (void)bytes;
#endif
}

void number_dofs(const Mesh &M, Dofs &D)
{
  const int dim = M.dim, nv1 = dim + 1, ne = (dim == 2) ? 3 : 6;
  D.dim = dim; D.nv1 = nv1; D.n2 = nv1 + ne; D.dpc = dpc_of(dim);
  D.nc = M.n_cells();
  const int64_t nv = M.n_vertices();
  std::vector<int> vnode(nv, -1), vp(nv, -1);
  // per-vertex singly linked edge lists: edge (a<b) stored at a
  std::vector<int> head(nv, -1);
  struct ERec { int b, node, next; };
  std::vector<ERec> pool;
  pool.reserve(size_t(D.nc) * (dim == 2 ? 2 : 2));
  reserve_prefaulted(D.cell_nodes, size_t(D.nc) * D.n2);
  D.cell_nodes.resize(size_t(D.nc) * D.n2);
  D.cell_p.resize(size_t(D.nc) * nv1);
  D.node_xyz.clear();
  D.p_xyz.clear();
  int n_nodes = 0, n_p = 0;
  for (int64_t c = 0; c < D.nc; ++c) {
    const int *v = &M.cells[c * nv1];
    for (int lv = 0; lv < nv1; ++lv) {
      const int g = v[lv];
      if (vnode[g] < 0) {
        vnode[g] = n_nodes++;
        vp[g] = n_p++;
        for (int d = 0; d < dim; ++d) {
          D.node_xyz.push_back(M.verts[size_t(g) * dim + d]);
          D.p_xyz.push_back(M.verts[size_t(g) * dim + d]);
        }
      }
      D.cell_nodes[c * D.n2 + lv] = vnode[g];
      D.cell_p[c * nv1 + lv] = vp[g];
    }
    for (int le = 0; le < ne; ++le) {
      int a = v[kEdgeA[le]], b = v[kEdgeB[le]];
      if (a > b) std::swap(a, b);
      int node = -1;
      for (int e = head[a]; e >= 0; e = pool[e].next)
        if (pool[e].b == b) { node = pool[e].node; break; }
      if (node < 0) {
        node = n_nodes++;
        pool.push_back({b, node, head[a]});
        head[a] = int(pool.size()) - 1;
        for (int d = 0; d < dim; ++d)
          D.node_xyz.push_back(0.5 * (M.verts[size_t(a) * dim + d] + M.verts[size_t(b) * dim + d]));
      }
      D.cell_nodes[c * D.n2 + nv1 + le] = node;
    }
  }
  D.n_nodes = n_nodes;
  D.n_p = n_p;
  const int n_u = dim * n_nodes;
  reserve_prefaulted(D.cell_dofs, size_t(D.nc) * D.dpc);
  reserve_prefaulted(D.cell_coords, size_t(D.nc) * nv1 * dim);
  D.cell_dofs.resize(size_t(D.nc) * D.dpc);
  D.cell_coords.resize(size_t(D.nc) * nv1 * dim);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < D.nc; ++c) {
    int *cd = &D.cell_dofs[c * D.dpc];
    const int *cn = &D.cell_nodes[c * D.n2];
    for (int lv = 0; lv < nv1; ++lv) {
      for (int k = 0; k < dim; ++k) cd[lv * (dim + 1) + k] = dim * cn[lv] + k;
      cd[lv * (dim + 1) + dim] = n_u + D.cell_p[c * nv1 + lv];
      for (int d = 0; d < dim; ++d)
        D.cell_coords[(c * nv1 + lv) * dim + d] = M.verts[size_t(M.cells[c * nv1 + lv]) * dim + d];
    }
    for (int le = 0; le < ne; ++le)
      for (int k = 0; k < dim; ++k) cd[nv1 * (dim + 1) + le * dim + k] = dim * cn[nv1 + le] + k;
  }
}

int face_local_nodes(int dim, int f, int out[6])
{
  const int nv1 = dim + 1, ne = (dim == 2) ? 3 : 6;
  int n = 0;
  for (int v = 0; v < nv1; ++v) if (v != f) out[n++] = v;
  for (int e = 0; e < ne; ++e)
    if (kEdgeA[e] != f && kEdgeB[e] != f) out[n++] = nv1 + e;
  return n;
}

// Rows of a pattern are built in chunks of kRowChunk consecutive rows by the host threads: every row is gathered,
// sorted and made unique ONCE into a chunk-local buffer, the row lengths are prefix-summed and the chunk buffers are
// copied to their place (before: two passes that sorted every row twice).  Independent of the thread count.
namespace {
constexpr int kRowChunk = 8192;
template <typename GatherRow>
void build_rows_chunked(int n_rows, Csr &out, GatherRow &&gather)
{
  out.rowptr.assign(size_t(n_rows) + 1, 0);
  const int nchunks = (n_rows + kRowChunk - 1) / kRowChunk;
  std::vector<std::vector<int>> chunk_cols(nchunks);
#pragma omp parallel
  {
    std::vector<int> buf;
#pragma omp for schedule(dynamic, 1)
    for (int ch = 0; ch < nchunks; ++ch) {
      const int r_lo = ch * kRowChunk, r_hi = std::min(n_rows, r_lo + kRowChunk);
      std::vector<int> &cols = chunk_cols[ch];
      for (int r = r_lo; r < r_hi; ++r) {
        buf.clear();
        gather(r, buf);
        std::sort(buf.begin(), buf.end());
        const int n = int(std::unique(buf.begin(), buf.end()) - buf.begin());
        out.rowptr[r + 1] = n;
        cols.insert(cols.end(), buf.begin(), buf.begin() + n);
      }
    }
  }
  for (int r = 0; r < n_rows; ++r) out.rowptr[r + 1] += out.rowptr[r];
  reserve_prefaulted(out.colind, size_t(out.rowptr[n_rows]));
  out.colind.resize(size_t(out.rowptr[n_rows]));
#pragma omp parallel for schedule(dynamic, 1)
  for (int ch = 0; ch < nchunks; ++ch) {
    const std::vector<int> &cols = chunk_cols[ch];
    if (!cols.empty()) std::memcpy(&out.colind[out.rowptr[size_t(ch) * kRowChunk]], cols.data(), sizeof(int) * cols.size());
  }
}
} // namespace

void RowCells::build(int64_t nc, const int *cell_rows, int kr, int n_rows_total)
{
  ptr.assign(size_t(n_rows_total) + 1, 0);
  for (int64_t c = 0; c < nc; ++c)
    for (int i = 0; i < kr; ++i) ptr[cell_rows[c * kr + i] + 1]++;
  for (int r = 0; r < n_rows_total; ++r) ptr[r + 1] += ptr[r];
  cells.resize(size_t(ptr[n_rows_total]));
  std::vector<int> pos(ptr.begin(), ptr.end() - 1);
  for (int64_t c = 0; c < nc; ++c)
    for (int i = 0; i < kr; ++i) cells[pos[cell_rows[c * kr + i]]++] = int(c);
}

void build_pattern(const RowCells &rc, const int *cell_cols, int kc, int n_rows_owned, int n_cols, Csr &out)
{
  out.n_rows = n_rows_owned;
  out.n_cols = n_cols;
  build_rows_chunked(n_rows_owned, out, [&](int r, std::vector<int> &buf) {
    for (int k = rc.ptr[r]; k < rc.ptr[r + 1]; ++k) {
      const int *cc = &cell_cols[int64_t(rc.cells[k]) * kc];
      buf.insert(buf.end(), cc, cc + kc);
    }
  });
}

void build_pattern(int64_t nc, const int *cell_rows, int kr, const int *cell_cols, int kc, int n_rows_total,
                   int n_rows_owned, int n_cols, Csr &out)
{
  RowCells rc;
  rc.build(nc, cell_rows, kr, n_rows_total);
  build_pattern(rc, cell_cols, kc, n_rows_owned, n_cols, out);
}

void symbolic_product(const Csr &A, const Csr &B, Csr &out)
{
  out.n_rows = A.n_rows;
  out.n_cols = B.n_cols;
  build_rows_chunked(A.n_rows, out, [&](int i, std::vector<int> &buf) {
    for (int p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) {
      const int k = A.colind[p];
      if (k >= B.n_rows) continue; // ghost row of B not available locally
      buf.insert(buf.end(), B.colind.begin() + B.rowptr[k], B.colind.begin() + B.rowptr[k + 1]);
    }
  });
}

// Recursive coordinate bisection of cell centroids (replaces GridTools::partition_triangulation /
// METIS, NavierStokes2D.cpp:16; the channel geometry is close to ideal for coordinate cuts).
static void rcb(const std::vector<double> &cen, int dim, std::vector<int> &idx, int lo, int hi, int p0, int np,
                std::vector<int> &part)
{
  if (np == 1) {
    for (int i = lo; i < hi; ++i) part[idx[i]] = p0;
    return;
  }
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = lo; i < hi; ++i)
    for (int d = 0; d < dim; ++d) {
      mn[d] = std::min(mn[d], cen[size_t(idx[i]) * dim + d]);
      mx[d] = std::max(mx[d], cen[size_t(idx[i]) * dim + d]);
    }
  int ax = 0;
  for (int d = 1; d < dim; ++d) if (mx[d] - mn[d] > mx[ax] - mn[ax]) ax = d;
  const int npl = np / 2;
  const int mid = lo + int(int64_t(hi - lo) * npl / np);
  std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int a, int b) {
    const double xa = cen[size_t(a) * dim + ax], xb = cen[size_t(b) * dim + ax];
    return xa < xb || (xa == xb && a < b);
  });
  rcb(cen, dim, idx, lo, mid, p0, npl, part);
  rcb(cen, dim, idx, mid, hi, p0 + npl, np - npl, part);
}

void partition_cells_rcb(const Mesh &M, int nparts, std::vector<int> &part)
{
  const int dim = M.dim, nv1 = dim + 1;
  const int64_t nc = M.n_cells();
  std::vector<double> cen(size_t(nc) * dim, 0.0);
  for (int64_t c = 0; c < nc; ++c)
    for (int k = 0; k < nv1; ++k)
      for (int d = 0; d < dim; ++d) cen[c * dim + d] += M.verts[size_t(M.cells[c * nv1 + k]) * dim + d] / nv1;
  std::vector<int> idx(nc);
  std::iota(idx.begin(), idx.end(), 0);
  part.assign(nc, 0);
  rcb(cen, dim, idx, 0, int(nc), 0, nparts, part);
}

} // namespace nsb
