// host_mesh.cpp -- simplex meshes for the three drivers (cold path, host only).
//
// The reference reads Gmsh files with GridIn::read_msh (Navier-Stokes/src/NavierStokes2D.cpp:5-23)
// but ships only .geo scripts and gmsh is not available, so this file provides
//   * generators with the reference's geometry and physical ids
//     (mesh/Cylinder2D.geo, mesh/Cylinder3D.geo, mesh/mesh-cube.geo),
//   * a Gmsh v2 / v4.1 ASCII reader and a v2 writer, so a real deal.II could read the same file.
// Block-structured O-grid around the cylinder, quads split into triangles; 3D meshes are
// extrusions into prisms split into 3 tetrahedra each with the smallest-vertex-index rule
// (conforming for any triangle mesh).
#include "nsb_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <unordered_map>

namespace nsb {

static inline uint64_t key2(int a, int b)
{
  if (a > b) std::swap(a, b);
  return (uint64_t(uint32_t(a)) << 32) | uint32_t(b);
}

void Mesh::fix_orientation()
{
  const int nv1 = dim + 1;
  const int64_t nc = n_cells();
  for (int64_t c = 0; c < nc; ++c) {
    int *v = &cells[c * nv1];
    double J[3][3];
    for (int r = 0; r < dim; ++r)
      for (int k = 0; k < dim; ++k) J[r][k] = verts[size_t(v[k + 1]) * dim + r] - verts[size_t(v[0]) * dim + r];
    double det;
    if (dim == 2)
      det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    else
      det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
            J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    if (det < 0) std::swap(v[0], v[1]);
  }
}

// Faces with exactly one adjacent cell; ids from the classifier.
// The (dim+1) * n_cells faces are bucketed by their smallest vertex (counting sort, linear) and every bucket --
// a few dozen faces -- is searched for faces that occur once; the buckets are independent, so the search runs
// on all host threads (a comparison sort of 18.7 M face records took 10 of the 14 s this mesh generator needed
// at 4.7 M tetrahedra).  Output order as before: by cell, then by local face.
void Mesh::build_boundary(const std::function<int(const Mesh &, const int *)> &classify, bool split_locked)
{
  const int nv1 = dim + 1;
  const int64_t nc = n_cells();
  const int64_t nf = nc * nv1;
  const size_t nv = verts.size() / dim;
  // sorted vertices of face f of cell c (v[2] = -1 in 2D)
  auto face = [&](int64_t id, int v[3]) {
    const int64_t c = id / nv1;
    const int f = int(id - c * nv1);
    int n = 0;
    v[2] = -1;
    for (int k = 0; k < nv1; ++k) if (k != f) v[n++] = cells[c * nv1 + k];
    if (v[0] > v[1]) std::swap(v[0], v[1]);
    if (dim == 3) {
      if (v[1] > v[2]) std::swap(v[1], v[2]);
      if (v[0] > v[1]) std::swap(v[0], v[1]);
    }
  };
  std::vector<int64_t> start(nv + 1, 0);
  for (int64_t id = 0; id < nf; ++id) {
    int v[3];
    face(id, v);
    start[size_t(v[0]) + 1]++;
  }
  for (size_t i = 0; i < nv; ++i) start[i + 1] += start[i];
  struct Rec { int v1, v2; int64_t id; };
  std::vector<Rec> recs;
  reserve_prefaulted(recs, size_t(nf));
  recs.resize(size_t(nf));
  {
    std::vector<int64_t> pos(start.begin(), start.end() - 1);
    for (int64_t id = 0; id < nf; ++id) { // face ids ascend inside every bucket
      int v[3];
      face(id, v);
      recs[pos[v[0]]++] = Rec{v[1], v[2], id};
    }
  }
  std::vector<unsigned char> is_single(nf, 0);
#pragma omp parallel for schedule(dynamic, 4096)
  for (int64_t i = 0; i < int64_t(nv); ++i) {
    Rec *r = &recs[start[i]];
    const int64_t n = start[i + 1] - start[i];
    std::sort(r, r + n, [](const Rec &x, const Rec &y) {
      if (x.v1 != y.v1) return x.v1 < y.v1;
      if (x.v2 != y.v2) return x.v2 < y.v2;
      return x.id < y.id;
    });
    for (int64_t a = 0; a < n;) {
      int64_t b = a + 1;
      while (b < n && r[b].v1 == r[a].v1 && r[b].v2 == r[a].v2) ++b;
      if (b - a == 1) is_single[r[a].id] = 1;
      a = b;
    }
  }
  bfaces.clear(); bids.clear(); bcell.clear(); blocal.clear();
  for (int64_t id = 0; id < nf; ++id) { // cell order, then local face
    if (!is_single[id]) continue;
    int v[3];
    face(id, v);
    for (int k = 0; k < dim; ++k) bfaces.push_back(v[k]);
    bids.push_back(classify(*this, v));
    bcell.push_back(int(id / nv1));
    blocal.push_back(int(id % nv1));
  }
  if (split_locked && split_boundary_locked_cells() > 0) build_boundary(classify, false);
}

// A cell whose edges ALL lie in boundary faces (a tetrahedron cutting off a box corner) leaves the
// pressure DoF of its corner vertex without any unconstrained velocity neighbour once those
// boundaries are Dirichlet: the Schur complement B D^-1 B^T then has a zero row and ILU(0) breaks
// down.  Such cells are split at their centroid (faces untouched => still conforming).
// Needs bfaces; the caller rebuilds the boundary afterwards.  Returns the number of split cells.
int Mesh::split_boundary_locked_cells()
{
  const int nv1 = dim + 1;
  std::unordered_map<uint64_t, char> bedge;
  const size_t nb = bids.size();
  for (size_t b = 0; b < nb; ++b) {
    const int *v = &bfaces[b * dim];
    if (dim == 2) bedge[key2(v[0], v[1])] = 1;
    else { bedge[key2(v[0], v[1])] = 1; bedge[key2(v[1], v[2])] = 1; bedge[key2(v[0], v[2])] = 1; }
  }
  const int64_t nc = n_cells();
  int n_split = 0;
  std::vector<int> extra;
  for (int64_t c = 0; c < nc; ++c) {
    int *v = &cells[c * nv1];
    bool locked = true;
    for (int a = 0; a < nv1 && locked; ++a)
      for (int b = a + 1; b < nv1; ++b)
        if (!bedge.count(key2(v[a], v[b]))) { locked = false; break; }
    if (!locked) continue;
    const int nv = int(verts.size() / dim);
    for (int d = 0; d < dim; ++d) {
      double s = 0;
      for (int k = 0; k < nv1; ++k) s += verts[size_t(v[k]) * dim + d];
      verts.push_back(s / nv1);
    }
    int orig[4];
    for (int k = 0; k < nv1; ++k) orig[k] = v[k];
    // sub-cell k replaces vertex k by the centroid; the first one overwrites the cell in place
    for (int k = 0; k < nv1; ++k) {
      int sub[4];
      for (int j = 0; j < nv1; ++j) sub[j] = (j == k) ? nv : orig[j];
      if (k == 0) for (int j = 0; j < nv1; ++j) v[j] = sub[j];
      else extra.insert(extra.end(), sub, sub + nv1);
    }
    ++n_split;
  }
  cells.insert(cells.end(), extra.begin(), extra.end());
  if (n_split) fix_orientation();
  return n_split;
}

// --------------------------------------------------------------------------------------------
// 2D block-structured channel with a circular hole
// --------------------------------------------------------------------------------------------
struct Tri2D {
  std::vector<double> xy;
  std::vector<int> tri;
  int add(double x, double y) { xy.push_back(x); xy.push_back(y); return int(xy.size() / 2) - 1; }
  void quad(int a, int b, int c, int d) { // split along a-c
    tri.insert(tri.end(), {a, b, c});
    tri.insert(tri.end(), {a, c, d});
  }
};

static Tri2D channel_with_hole(int s, double xc, double yc, double R, double H, double L)
{
  const int m = 6 * s, nr = 5 * s;
  const double x0 = xc - 0.2, x1 = x0 + H;
  const double q = std::pow(1.35, 1.0 / s);
  Tri2D T;
  auto sq = [&](int k, double &x, double &y) {
    k = ((k % (4 * m)) + 4 * m) % (4 * m);
    if (k <= m) { x = x1; y = H * k / m; }
    else if (k <= 2 * m) { x = x1 - H * (k - m) / m; y = H; }
    else if (k <= 3 * m) { x = x0; y = H - H * (k - 2 * m) / m; }
    else { x = x0 + H * (k - 3 * m) / m; y = 0.0; }
  };
  std::vector<int> ring(size_t(4 * m) * (nr + 1));
  auto R_ = [&](int k, int l) -> int & { return ring[size_t(((k % (4 * m)) + 4 * m) % (4 * m)) * (nr + 1) + l]; };
  const double qn = std::pow(q, nr) - 1.0;
  for (int k = 0; k < 4 * m; ++k) {
    const double th = -M_PI / 4 + 2 * M_PI * k / (4 * m);
    const double cx = xc + R * std::cos(th), cy = yc + R * std::sin(th);
    double sx, sy;
    sq(k, sx, sy);
    for (int l = 0; l <= nr; ++l) {
      const double g = (l == nr) ? 1.0 : (std::pow(q, l) - 1.0) / qn;
      R_(k, l) = T.add(cx + g * (sx - cx), cy + g * (sy - cy));
    }
  }
  for (int k = 0; k < 4 * m; ++k)
    for (int l = 0; l < nr; ++l) T.quad(R_(k, l), R_(k, l + 1), R_(k + 1, l + 1), R_(k + 1, l));
  // downstream block, column 0 = east side of the square (k = 0..m)
  const int nxd = std::max(1, int(std::lround((L - x1) / (H / m))));
  std::vector<int> prev(m + 1), cur(m + 1);
  for (int j = 0; j <= m; ++j) prev[j] = R_(j, nr);
  for (int i = 1; i <= nxd; ++i) {
    for (int j = 0; j <= m; ++j) cur[j] = T.add(x1 + (L - x1) * i / nxd, H * j / m);
    for (int j = 0; j < m; ++j) T.quad(prev[j], cur[j], cur[j + 1], prev[j + 1]);
    prev = cur;
  }
  // upstream block (only when the square does not start at the inlet)
  if (x0 > 1e-12) {
    const int nxu = std::max(1, int(std::lround(x0 / (H / m))));
    for (int j = 0; j <= m; ++j) prev[j] = R_(3 * m - j, nr); // west side, y = H*j/m
    for (int i = 1; i <= nxu; ++i) {
      for (int j = 0; j <= m; ++j) cur[j] = T.add(x0 - x0 * i / nxu, H * j / m);
      for (int j = 0; j < m; ++j) T.quad(cur[j], prev[j], prev[j + 1], cur[j + 1]);
      prev = cur;
    }
  }
  return T;
}

static Tri2D rectangle(int nx, int ny, double x0, double x1, double y0, double y1)
{
  Tri2D T;
  std::vector<int> id(size_t(nx + 1) * (ny + 1));
  for (int i = 0; i <= nx; ++i)
    for (int j = 0; j <= ny; ++j) id[size_t(i) * (ny + 1) + j] = T.add(x0 + (x1 - x0) * i / nx, y0 + (y1 - y0) * j / ny);
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j)
      T.quad(id[size_t(i) * (ny + 1) + j], id[size_t(i + 1) * (ny + 1) + j], id[size_t(i + 1) * (ny + 1) + j + 1],
             id[size_t(i) * (ny + 1) + j + 1]);
  return T;
}

static Mesh from_tri2d(const Tri2D &T)
{
  Mesh M;
  M.dim = 2;
  M.verts = T.xy;
  M.cells = T.tri;
  M.fix_orientation();
  return M;
}

// Prism (a,b,c | a',b',c') -> 3 tets, diagonals from the smallest global index
// (Dompierre et al., "How to subdivide pyramids, prisms and hexahedra into tetrahedra").
static void split_prism(const int V[6], std::vector<int> &out)
{
  static const int rot[6][6] = {{0, 1, 2, 3, 4, 5}, {1, 2, 0, 4, 5, 3}, {2, 0, 1, 5, 3, 4},
                                {3, 5, 4, 0, 2, 1}, {4, 3, 5, 1, 0, 2}, {5, 4, 3, 2, 1, 0}};
  int smallest = 0;
  for (int i = 1; i < 6; ++i) if (V[i] < V[smallest]) smallest = i;
  int W[6];
  for (int i = 0; i < 6; ++i) W[i] = V[rot[smallest][i]];
  if (std::min(W[1], W[5]) < std::min(W[2], W[4])) {
    out.insert(out.end(), {W[0], W[1], W[2], W[5]});
    out.insert(out.end(), {W[0], W[1], W[5], W[4]});
    out.insert(out.end(), {W[0], W[4], W[5], W[3]});
  } else {
    out.insert(out.end(), {W[0], W[1], W[2], W[4]});
    out.insert(out.end(), {W[0], W[4], W[2], W[5]});
    out.insert(out.end(), {W[0], W[4], W[5], W[3]});
  }
}

static Mesh extrude(const Tri2D &T, int nz, double z0, double z1)
{
  Mesh M;
  M.dim = 3;
  const int nv2 = int(T.xy.size() / 2);
  const int nt = int(T.tri.size() / 3);
  M.verts.resize(size_t(nv2) * (nz + 1) * 3);
  for (int k = 0; k <= nz; ++k)
    for (int v = 0; v < nv2; ++v) {
      double *p = &M.verts[(size_t(k) * nv2 + v) * 3];
      p[0] = T.xy[2 * v]; p[1] = T.xy[2 * v + 1]; p[2] = z0 + (z1 - z0) * k / nz;
    }
  M.cells.reserve(size_t(nt) * nz * 12);
  for (int t = 0; t < nt; ++t)
    for (int k = 0; k < nz; ++k) {
      int V[6];
      for (int i = 0; i < 3; ++i) {
        V[i] = k * nv2 + T.tri[3 * t + i];
        V[3 + i] = (k + 1) * nv2 + T.tri[3 * t + i];
      }
      split_prism(V, M.cells);
    }
  M.fix_orientation();
  return M;
}

static bool all_on(const Mesh &M, const int *v, int axis, double value, double tol = 1e-9)
{
  for (int k = 0; k < M.dim; ++k)
    if (std::fabs(M.verts[size_t(v[k]) * M.dim + axis] - value) > tol) return false;
  return true;
}

Mesh make_cylinder2d(int s)
{
  const double H = 0.41, L = 2.2;
  Mesh M = from_tri2d(channel_with_hole(s, 0.2, 0.2, 0.05, H, L));
  M.build_boundary([=](const Mesh &m, const int *v) {
    if (all_on(m, v, 0, 0.0)) return 0;           // Physical Line(0): inlet   (Cylinder2D.geo:40)
    if (all_on(m, v, 0, L)) return 1;             // outlet
    if (all_on(m, v, 1, 0.0) || all_on(m, v, 1, H)) return 2; // walls
    return 3;                                     // cylinder
  });
  return M;
}

Mesh make_cylinder3d(int s, int nz)
{
  const double H = 0.41, L = 2.5;
  Mesh M = extrude(channel_with_hole(s, 0.5, 0.2, 0.05, H, L), nz, 0.0, H);
  M.build_boundary([=](const Mesh &m, const int *v) {
    if (all_on(m, v, 0, 0.0)) return 0;           // Physical Surface(0): inlet (Cylinder3D.geo:126)
    if (all_on(m, v, 0, L)) return 1;
    if (all_on(m, v, 1, 0.0) || all_on(m, v, 1, H) || all_on(m, v, 2, 0.0) || all_on(m, v, 2, H)) return 2;
    return 3;
  });
  return M;
}

Mesh make_box(int dim, int nx, int ny, int nz, const double *lo, const double *hi)
{
  Tri2D T = rectangle(nx, ny, lo[0], hi[0], lo[1], hi[1]);
  Mesh M = (dim == 2) ? from_tri2d(T) : extrude(T, nz, lo[2], hi[2]);
  double l[3] = {lo[0], lo[1], dim == 3 ? lo[2] : 0}, h[3] = {hi[0], hi[1], dim == 3 ? hi[2] : 0};
  // ids as in mesh/mesh-cube.geo:16-21 (surfaces 14,22,18,26,5,27 = x-, x+, y-, y+, z-, z+)
  M.build_boundary([=](const Mesh &m, const int *v) {
    if (all_on(m, v, 0, l[0])) return 0;
    if (all_on(m, v, 0, h[0])) return 1;
    if (all_on(m, v, 1, l[1])) return 2;
    if (all_on(m, v, 1, h[1])) return 3;
    if (m.dim == 3 && all_on(m, v, 2, l[2])) return 4;
    return 5;
  });
  return M;
}

Mesh make_cube(int n)
{
  const double lo[3] = {-1, -1, -1}, hi[3] = {1, 1, 1};
  return make_box(3, n, n, n, lo, hi);
}

// --------------------------------------------------------------------------------------------
// cell reordering: mode 1 = blocks of `block` consecutive... spatial blocks ordered colour by
// colour (2^dim colours) so that the natural-order ILU dependency graph is shallow across blocks.
// --------------------------------------------------------------------------------------------
void Mesh::reorder_cells(int mode, int block)
{
  if (mode == 0) return;
  const int nv1 = dim + 1;
  const int64_t nc = n_cells();
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (size_t i = 0; i < verts.size() / dim; ++i)
    for (int d = 0; d < dim; ++d) {
      lo[d] = std::min(lo[d], verts[i * dim + d]);
      hi[d] = std::max(hi[d], verts[i * dim + d]);
    }
  // target ~`block` cells per spatial box
  const double vol_per_box = double(block) / double(nc);
  double ext[3] = {1, 1, 1};
  double vol = 1;
  for (int d = 0; d < dim; ++d) { ext[d] = hi[d] - lo[d]; vol *= ext[d]; }
  const double hbox = std::pow(vol * vol_per_box, 1.0 / dim);
  int nb[3] = {1, 1, 1};
  for (int d = 0; d < dim; ++d) nb[d] = std::max(1, int(std::ceil(ext[d] / hbox)));
  struct Key { int colour; int64_t box; int64_t cell; };
  std::vector<Key> keys(nc);
  for (int64_t c = 0; c < nc; ++c) {
    int b[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d) {
      double x = 0;
      for (int k = 0; k < nv1; ++k) x += verts[size_t(cells[c * nv1 + k]) * dim + d];
      x /= nv1;
      b[d] = std::min(nb[d] - 1, std::max(0, int((x - lo[d]) / ext[d] * nb[d])));
    }
    Key k;
    k.colour = (mode == 1) ? ((b[0] & 1) | ((b[1] & 1) << 1) | ((b[2] & 1) << 2)) : 0;
    k.box = (int64_t(b[2]) * nb[1] + b[1]) * nb[0] + b[0];
    k.cell = c;
    keys[c] = k;
  }
  std::stable_sort(keys.begin(), keys.end(), [](const Key &a, const Key &b) {
    if (a.colour != b.colour) return a.colour < b.colour;
    if (a.box != b.box) return a.box < b.box;
    return a.cell < b.cell;
  });
  std::vector<int> nc2(cells.size());
  std::vector<int> new_of_old(nc);
  for (int64_t i = 0; i < nc; ++i) {
    std::memcpy(&nc2[i * nv1], &cells[keys[i].cell * nv1], sizeof(int) * nv1);
    new_of_old[keys[i].cell] = int(i);
  }
  cells.swap(nc2);
  for (auto &c : bcell) c = new_of_old[c];
}

// --------------------------------------------------------------------------------------------
// Gmsh I/O
// --------------------------------------------------------------------------------------------
bool write_msh(const Mesh &M, const std::string &path)
{
  FILE *f = std::fopen(path.c_str(), "w");
  if (!f) return false;
  const int dim = M.dim, nv1 = dim + 1;
  const size_t nv = M.verts.size() / dim, nc = M.n_cells(), nb = M.bids.size();
  std::fprintf(f, "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%zu\n", nv);
  for (size_t i = 0; i < nv; ++i)
    std::fprintf(f, "%zu %.17g %.17g %.17g\n", i + 1, M.verts[i * dim], M.verts[i * dim + 1],
                 dim == 3 ? M.verts[i * dim + 2] : 0.0);
  std::fprintf(f, "$EndNodes\n$Elements\n%zu\n", nb + nc);
  size_t id = 1;
  for (size_t b = 0; b < nb; ++b) {
    std::fprintf(f, "%zu %d 2 %d %d", id++, dim == 2 ? 1 : 2, M.bids[b], M.bids[b] + 1);
    for (int k = 0; k < dim; ++k) std::fprintf(f, " %d", M.bfaces[b * dim + k] + 1);
    std::fprintf(f, "\n");
  }
  const int vol_tag = (dim == 2) ? 4 : 10;
  for (size_t c = 0; c < nc; ++c) {
    std::fprintf(f, "%zu %d 2 %d 1", id++, dim == 2 ? 2 : 4, vol_tag);
    for (int k = 0; k < nv1; ++k) std::fprintf(f, " %d", M.cells[c * nv1 + k] + 1);
    std::fprintf(f, "\n");
  }
  std::fprintf(f, "$EndElements\n");
  std::fclose(f);
  return true;
}

static int nodes_of_type(int t)
{
  switch (t) { case 1: return 2; case 2: return 3; case 4: return 4; case 15: return 1; case 3: return 4;
    case 5: return 8; case 8: return 3; case 9: return 6; case 11: return 10; default: return -1; }
}

bool read_msh(const std::string &path, Mesh &M, std::string &err)
{
  std::ifstream in(path);
  if (!in) { err = "cannot open " + path; return false; }
  std::string line;
  double version = 2.2;
  std::map<long, int> node_index;           // gmsh tag -> 0-based
  std::vector<double> xyz;                  // 3 per node
  struct Elem { int type; int phys; std::vector<long> nodes; };
  std::vector<Elem> elems;
  std::map<std::pair<int, int>, int> entity_phys; // (dim, tag) -> first physical tag (v4)
  while (std::getline(in, line)) {
    if (line.rfind("$MeshFormat", 0) == 0) {
      int ft, ds;
      in >> version >> ft >> ds;
      if (ft != 0) { err = "binary .msh is not supported"; return false; }
    } else if (line.rfind("$Entities", 0) == 0 && version >= 4.0) {
      size_t np, nc, ns, nvv;
      in >> np >> nc >> ns >> nvv;
      for (size_t i = 0; i < np; ++i) {
        int tag; double x, y, z; size_t nph;
        in >> tag >> x >> y >> z >> nph;
        for (size_t k = 0; k < nph; ++k) { int p; in >> p; if (k == 0) entity_phys[{0, tag}] = p; }
      }
      const size_t cnt[3] = {nc, ns, nvv};
      for (int d = 1; d <= 3; ++d)
        for (size_t i = 0; i < cnt[d - 1]; ++i) {
          int tag; double b[6]; size_t nph, nbnd;
          in >> tag;
          for (double &v : b) in >> v;
          in >> nph;
          for (size_t k = 0; k < nph; ++k) { int p; in >> p; if (k == 0) entity_phys[{d, tag}] = p; }
          in >> nbnd;
          for (size_t k = 0; k < nbnd; ++k) { int t; in >> t; }
        }
    } else if (line.rfind("$Nodes", 0) == 0) {
      if (version < 4.0) {
        size_t n; in >> n;
        for (size_t i = 0; i < n; ++i) {
          long tag; double x, y, z;
          in >> tag >> x >> y >> z;
          node_index[tag] = int(xyz.size() / 3);
          xyz.insert(xyz.end(), {x, y, z});
        }
      } else {
        size_t nblocks, nn, mn, mx;
        in >> nblocks >> nn >> mn >> mx;
        for (size_t b = 0; b < nblocks; ++b) {
          int ed, et, par; size_t nb;
          in >> ed >> et >> par >> nb;
          std::vector<long> tags(nb);
          for (auto &t : tags) in >> t;
          for (size_t i = 0; i < nb; ++i) {
            double x, y, z; in >> x >> y >> z;
            node_index[tags[i]] = int(xyz.size() / 3);
            xyz.insert(xyz.end(), {x, y, z});
          }
        }
      }
    } else if (line.rfind("$Elements", 0) == 0) {
      if (version < 4.0) {
        size_t n; in >> n;
        for (size_t i = 0; i < n; ++i) {
          long id; int type, ntags;
          in >> id >> type >> ntags;
          Elem e; e.type = type; e.phys = 0;
          for (int k = 0; k < ntags; ++k) { int t; in >> t; if (k == 0) e.phys = t; }
          const int nn = nodes_of_type(type);
          if (nn < 0) { err = "unsupported element type"; return false; }
          e.nodes.resize(nn);
          for (auto &v : e.nodes) in >> v;
          elems.push_back(std::move(e));
        }
      } else {
        size_t nblocks, ne, mn, mx;
        in >> nblocks >> ne >> mn >> mx;
        for (size_t b = 0; b < nblocks; ++b) {
          int ed, et, type; size_t nb;
          in >> ed >> et >> type >> nb;
          const int nn = nodes_of_type(type);
          if (nn < 0) { err = "unsupported element type"; return false; }
          auto it = entity_phys.find({ed, et});
          const int phys = it == entity_phys.end() ? 0 : it->second;
          for (size_t i = 0; i < nb; ++i) {
            long id; in >> id;
            Elem e; e.type = type; e.phys = phys; e.nodes.resize(nn);
            for (auto &v : e.nodes) in >> v;
            elems.push_back(std::move(e));
          }
        }
      }
    }
  }
  bool has_tet = false, has_tri = false;
  for (auto &e : elems) { has_tet |= e.type == 4; has_tri |= e.type == 2; }
  if (!has_tet && !has_tri) { err = "no simplex cells in " + path; return false; }
  M = Mesh();
  M.dim = has_tet ? 3 : 2;
  const int dim = M.dim, cell_type = has_tet ? 4 : 2, face_type = has_tet ? 2 : 1;
  M.verts.resize(xyz.size() / 3 * dim);
  for (size_t i = 0; i < xyz.size() / 3; ++i)
    for (int d = 0; d < dim; ++d) M.verts[i * dim + d] = xyz[i * 3 + d];
  std::map<std::vector<int>, int> face_id;
  for (auto &e : elems) {
    if (e.type == cell_type)
      for (long v : e.nodes) M.cells.push_back(node_index.at(v));
    else if (e.type == face_type) {
      std::vector<int> k;
      for (long v : e.nodes) k.push_back(node_index.at(v));
      std::sort(k.begin(), k.end());
      face_id[k] = e.phys;
    }
  }
  M.fix_orientation();
  M.build_boundary([&](const Mesh &m, const int *v) {
    std::vector<int> k(v, v + m.dim);
    auto it = face_id.find(k);
    return it == face_id.end() ? 0 : it->second; // deal.II default boundary id 0
  }, false);
  return true;
}

} // namespace nsb
