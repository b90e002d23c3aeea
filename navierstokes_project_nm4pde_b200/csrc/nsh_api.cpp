// nsh_api.cpp -- C ABI of the host prerequisites (mesh, DoF numbering, partition); no GPU needed.
#include <cstdio>
#include <cstring>

#include "../../include/nsb.h"
#include "nsb_host.hpp"

struct nsh_mesh_s { nsb::Mesh M; };
struct nsh_dofs_s { nsb::Dofs D; };
namespace nsb { // for host_local.cpp
const Mesh &mesh_of(const nsh_mesh_s *m) { return m->M; }
const Dofs &dofs_of(const nsh_dofs_s *d) { return d->D; }
} // namespace nsb

extern "C" {

nsh_mesh nsh_mesh_cylinder2d(int s)
{
  if (s < 1) return nullptr;
  auto *m = new nsh_mesh_s();
  m->M = nsb::make_cylinder2d(s);
  return m;
}
nsh_mesh nsh_mesh_cylinder3d(int s, int nz)
{
  if (s < 1 || nz < 1) return nullptr;
  auto *m = new nsh_mesh_s();
  m->M = nsb::make_cylinder3d(s, nz);
  return m;
}
nsh_mesh nsh_mesh_cube(int n)
{
  if (n < 1) return nullptr;
  auto *m = new nsh_mesh_s();
  m->M = nsb::make_cube(n);
  return m;
}
nsh_mesh nsh_mesh_box(int dim, int nx, int ny, int nz, const double *lo, const double *hi)
{
  if ((dim != 2 && dim != 3) || nx < 1 || ny < 1 || (dim == 3 && nz < 1) || !lo || !hi) return nullptr;
  auto *m = new nsh_mesh_s();
  m->M = nsb::make_box(dim, nx, ny, nz, lo, hi);
  return m;
}
nsh_mesh nsh_mesh_read_msh(const char *path)
{
  auto *m = new nsh_mesh_s();
  std::string err;
  if (!path || !nsb::read_msh(path, m->M, err)) { delete m; return nullptr; }
  return m;
}
int nsh_mesh_write_msh(nsh_mesh m, const char *path)
{
  if (!m || !path) return NSB_ERR_ARG;
  return nsb::write_msh(m->M, path) ? NSB_OK : NSB_ERR_IO;
}
void nsh_mesh_free(nsh_mesh m) { delete m; }
int nsh_mesh_dim(nsh_mesh m) { return m ? m->M.dim : 0; }
int32_t nsh_mesh_n_vertices(nsh_mesh m) { return m ? int32_t(m->M.n_vertices()) : 0; }
int32_t nsh_mesh_n_cells(nsh_mesh m) { return m ? int32_t(m->M.n_cells()) : 0; }
int32_t nsh_mesh_n_bfaces(nsh_mesh m) { return m ? int32_t(m->M.bids.size()) : 0; }
const double *nsh_mesh_vertices(nsh_mesh m) { return m ? m->M.verts.data() : nullptr; }
const int32_t *nsh_mesh_cells(nsh_mesh m) { return m ? m->M.cells.data() : nullptr; }
const int32_t *nsh_mesh_bfaces(nsh_mesh m) { return m ? m->M.bfaces.data() : nullptr; }
const int32_t *nsh_mesh_bface_ids(nsh_mesh m) { return m ? m->M.bids.data() : nullptr; }
const int32_t *nsh_mesh_bface_cells(nsh_mesh m) { return m ? m->M.bcell.data() : nullptr; }
int nsh_mesh_reorder_cells(nsh_mesh m, int mode, int block)
{
  if (!m || mode < 0 || mode > 2 || block < 1) return NSB_ERR_ARG;
  m->M.reorder_cells(mode, block);
  return NSB_OK;
}

nsh_dofs nsh_dofs_create(nsh_mesh m)
{
  if (!m) return nullptr;
  auto *d = new nsh_dofs_s();
  nsb::number_dofs(m->M, d->D);
  return d;
}
void nsh_dofs_free(nsh_dofs d) { delete d; }
int32_t nsh_dofs_n_nodes(nsh_dofs d) { return d ? d->D.n_nodes : 0; }
int32_t nsh_dofs_n_p(nsh_dofs d) { return d ? d->D.n_p : 0; }
int32_t nsh_dofs_per_cell(nsh_dofs d) { return d ? d->D.dpc : 0; }
const int32_t *nsh_dofs_cell_dofs(nsh_dofs d) { return d ? d->D.cell_dofs.data() : nullptr; }
const double *nsh_dofs_node_xyz(nsh_dofs d) { return d ? d->D.node_xyz.data() : nullptr; }
const double *nsh_dofs_p_xyz(nsh_dofs d) { return d ? d->D.p_xyz.data() : nullptr; }
const double *nsh_dofs_cell_coords(nsh_dofs d) { return d ? d->D.cell_coords.data() : nullptr; }

int32_t nsh_dofs_boundary_nodes(nsh_dofs d, nsh_mesh m, const int32_t *ids, int32_t n_ids, int32_t *out)
{
  if (!d || !m) return -1;
  const nsb::Mesh &M = m->M;
  const nsb::Dofs &D = d->D;
  std::vector<char> seen(D.n_nodes, 0);
  int32_t n = 0;
  for (size_t b = 0; b < M.bids.size(); ++b) {
    bool want = false;
    for (int k = 0; k < n_ids; ++k) want |= (ids[k] == M.bids[b]);
    if (!want) continue;
    int loc[6];
    const int nl = nsb::face_local_nodes(D.dim, M.blocal[b], loc);
    for (int k = 0; k < nl; ++k) {
      const int node = D.cell_nodes[size_t(M.bcell[b]) * D.n2 + loc[k]];
      if (!seen[node]) {
        seen[node] = 1;
        if (out) out[n] = node;
        ++n;
      }
    }
  }
  return n;
}

int32_t nsh_dofs_boundary_faces(nsh_dofs d, nsh_mesh m, int32_t id, int32_t *face_cell, int32_t *face_local)
{
  if (!d || !m) return -1;
  const nsb::Mesh &M = m->M;
  int32_t n = 0;
  for (size_t b = 0; b < M.bids.size(); ++b)
    if (M.bids[b] == id) {
      if (face_cell) face_cell[n] = M.bcell[b];
      if (face_local) face_local[n] = M.blocal[b];
      ++n;
    }
  return n;
}

// Cell containing x: the first cell whose barycentric coordinates are all >= -1e-10 (lam[dim + 1] filled),
// or -1 when no cell holds the point (deal.II: ExcPointNotAvailableHere).
int32_t nsh_dofs_find_cell(nsh_dofs d, const double *x, double *lam)
{
  if (!d || !x || !lam) return -1;
  const nsb::Dofs &D = d->D;
  const int dim = D.dim, nv1 = D.nv1;
  for (int64_t c = 0; c < D.nc; ++c) {
    const double *X = &D.cell_coords[size_t(c) * nv1 * dim];
    // solve J lam' = x - x0 for the barycentric coordinates lam_1..lam_dim
    double J[3][3], b[3];
    for (int r = 0; r < dim; ++r) {
      b[r] = x[r] - X[r];
      for (int k = 0; k < dim; ++k) J[r][k] = X[(k + 1) * dim + r] - X[r];
    }
    if (dim == 2) {
      const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      lam[1] = (b[0] * J[1][1] - J[0][1] * b[1]) / det;
      lam[2] = (J[0][0] * b[1] - b[0] * J[1][0]) / det;
      lam[0] = 1.0 - lam[1] - lam[2];
    } else {
      const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                   c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
      // Cramer's rule, column by column
      auto det3 = [](const double a[3], const double bb[3], const double cc[3]) {
        return a[0] * (bb[1] * cc[2] - bb[2] * cc[1]) - bb[0] * (a[1] * cc[2] - a[2] * cc[1]) + cc[0] * (a[1] * bb[2] - a[2] * bb[1]);
      };
      const double c0[3] = {J[0][0], J[1][0], J[2][0]}, c1[3] = {J[0][1], J[1][1], J[2][1]}, c2[3] = {J[0][2], J[1][2], J[2][2]};
      lam[1] = det3(b, c1, c2) / det;
      lam[2] = det3(c0, b, c2) / det;
      lam[3] = det3(c0, c1, b) / det;
      lam[0] = 1.0 - lam[1] - lam[2] - lam[3];
    }
    bool inside = true;
    for (int v = 0; v < nv1; ++v) inside &= (lam[v] >= -1e-10);
    if (inside) return int32_t(c);
  }
  return -1;
}

// VectorTools::point_value of the (P2^dim, P1) solution at x: evaluates the dim velocity components and the
// pressure in the cell nsh_dofs_find_cell returns.  out[dim + 1]; returns 0, or NSB_ERR_ARG when no cell
// contains the point.
int nsh_dofs_point_value(nsh_dofs d, const double *solution, const double *x, double *out)
{
  if (!d || !solution || !x || !out) return NSB_ERR_ARG;
  const nsb::Dofs &D = d->D;
  const int dim = D.dim, nv1 = D.nv1, n2 = D.n2;
  const int64_t n_u = int64_t(dim) * D.n_nodes;
  double lam[4];
  const int32_t c = nsh_dofs_find_cell(d, x, lam);
  if (c < 0) return NSB_ERR_ARG;
  for (int k = 0; k <= dim; ++k) out[k] = 0.0;
  const int *cn = &D.cell_nodes[size_t(c) * n2];
  const int *cp = &D.cell_p[size_t(c) * nv1];
  for (int v = 0; v < nv1; ++v) {
    const double ph = lam[v] * (2.0 * lam[v] - 1.0);
    for (int k = 0; k < dim; ++k) out[k] += ph * solution[size_t(dim) * cn[v] + k];
    out[dim] += lam[v] * solution[n_u + cp[v]];
  }
  for (int e = 0; e < n2 - nv1; ++e) {
    const double ph = 4.0 * lam[nsb::kEdgeA[e]] * lam[nsb::kEdgeB[e]];
    for (int k = 0; k < dim; ++k) out[k] += ph * solution[size_t(dim) * cn[nv1 + e] + k];
  }
  return NSB_OK;
}

// Minimal stand-in for DataOut::write_vtu (src/NavierStokes2D.cpp:642-675): one ASCII .vtu with the
// linear simplices of the mesh and, per vertex, "velocity" (3 components) and "pressure".  The P2
// edge values are not written (deal.II's build_patches() without subdivisions does the same).
int nsh_write_vtu(nsh_mesh m, nsh_dofs d, const double *solution, const char *path)
{
  if (!m || !d || !solution || !path) return NSB_ERR_ARG;
  const nsb::Mesh &M = m->M;
  const nsb::Dofs &D = d->D;
  const int dim = M.dim, nv1 = dim + 1;
  const int64_t nv = M.n_vertices(), nc = M.n_cells(), n_u = int64_t(dim) * D.n_nodes;
  // vertex v <-> pressure DoF / P2 vertex node: take them from the first cell that touches v
  std::vector<int> node_of(nv, -1), p_of(nv, -1);
  for (int64_t c = 0; c < nc; ++c)
    for (int k = 0; k < nv1; ++k) {
      const int v = M.cells[c * nv1 + k];
      node_of[v] = D.cell_nodes[size_t(c) * D.n2 + k];
      p_of[v] = D.cell_p[size_t(c) * nv1 + k];
    }
  FILE *f = std::fopen(path, "w");
  if (!f) return NSB_ERR_IO;
  std::fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n");
  std::fprintf(f, "<UnstructuredGrid>\n<Piece NumberOfPoints=\"%lld\" NumberOfCells=\"%lld\">\n", (long long)nv, (long long)nc);
  std::fprintf(f, "<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n");
  for (int64_t v = 0; v < nv; ++v)
    std::fprintf(f, "%.17g %.17g %.17g\n", M.verts[v * dim], M.verts[v * dim + 1], dim == 3 ? M.verts[v * dim + 2] : 0.0);
  std::fprintf(f, "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int32\" Name=\"connectivity\" format=\"ascii\">\n");
  for (int64_t c = 0; c < nc; ++c) {
    for (int k = 0; k < nv1; ++k) std::fprintf(f, "%d ", M.cells[c * nv1 + k]);
    std::fprintf(f, "\n");
  }
  std::fprintf(f, "</DataArray>\n<DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">\n");
  for (int64_t c = 0; c < nc; ++c) std::fprintf(f, "%lld\n", (long long)(c + 1) * nv1);
  std::fprintf(f, "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n");
  for (int64_t c = 0; c < nc; ++c) std::fprintf(f, "%d\n", dim == 2 ? 5 : 10); // VTK_TRIANGLE / VTK_TETRA
  std::fprintf(f, "</DataArray>\n</Cells>\n<PointData Vectors=\"velocity\" Scalars=\"pressure\">\n");
  std::fprintf(f, "<DataArray type=\"Float64\" Name=\"velocity\" NumberOfComponents=\"3\" format=\"ascii\">\n");
  for (int64_t v = 0; v < nv; ++v) {
    const double *u = solution + size_t(dim) * node_of[v];
    std::fprintf(f, "%.17g %.17g %.17g\n", u[0], u[1], dim == 3 ? u[2] : 0.0);
  }
  std::fprintf(f, "</DataArray>\n<DataArray type=\"Float64\" Name=\"pressure\" format=\"ascii\">\n");
  for (int64_t v = 0; v < nv; ++v) std::fprintf(f, "%.17g\n", solution[n_u + p_of[v]]);
  std::fprintf(f, "</DataArray>\n</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n");
  return std::fclose(f) == 0 ? NSB_OK : NSB_ERR_IO;
}

int nsh_partition_cells(nsh_mesh m, int nparts, int32_t *part)
{
  if (!m || nparts < 1 || !part) return NSB_ERR_ARG;
  std::vector<int> p;
  nsb::partition_cells_rcb(m->M, nparts, p);
  std::memcpy(part, p.data(), sizeof(int) * p.size());
  return NSB_OK;
}

} // extern "C"
