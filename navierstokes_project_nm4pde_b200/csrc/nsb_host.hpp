// nsb_host.hpp -- host-side (cold path) data structures shared by the mesh / DoF / pattern code.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace nsb {

struct Mesh {
  int dim = 0;
  std::vector<double> verts; // [nv][dim]
  std::vector<int> cells;    // [nc][dim+1], positively oriented
  std::vector<int> bfaces;   // [nb][dim] sorted vertex ids
  std::vector<int> bids;     // [nb] physical id
  std::vector<int> bcell;    // [nb] adjacent cell
  std::vector<int> blocal;   // [nb] local face (= local index of the opposite vertex)
  int64_t n_cells() const { return dim ? int64_t(cells.size()) / (dim + 1) : 0; }
  int64_t n_vertices() const { return dim ? int64_t(verts.size()) / dim : 0; }
  void fix_orientation();
  void build_boundary(const std::function<int(const Mesh &, const int *)> &classify, bool split_locked = true);
  int split_boundary_locked_cells();
  void reorder_cells(int mode, int block);
};

Mesh make_cylinder2d(int s);
Mesh make_cylinder3d(int s, int nz);
Mesh make_cube(int n);
Mesh make_box(int dim, int nx, int ny, int nz, const double *lo, const double *hi);
bool write_msh(const Mesh &M, const std::string &path);
bool read_msh(const std::string &path, Mesh &M, std::string &err);

// local edges of a simplex in deal.II order
static const int kEdgeA[6] = {0, 1, 2, 0, 1, 2};
static const int kEdgeB[6] = {1, 2, 0, 3, 3, 3};

inline int n2_of(int dim) { return dim == 2 ? 6 : 10; }
inline int dpc_of(int dim) { return dim * n2_of(dim) + dim + 1; }

// DoF numbering of FESystem(P2^dim, P1) after component_wise renumbering, in compact form:
// P2 node ids (velocity DoF = dim*node + c) and pressure ids (DoF = n_u + p).
struct Dofs {
  int dim = 0, n2 = 0, nv1 = 0, dpc = 0;
  int64_t nc = 0;
  int n_nodes = 0, n_p = 0;
  std::vector<int> cell_nodes;      // [nc][n2]
  std::vector<int> cell_p;          // [nc][nv1]
  std::vector<int> cell_dofs;       // [nc][dpc] reference layout
  std::vector<double> node_xyz;     // [n_nodes][dim]
  std::vector<double> p_xyz;        // [n_p][dim]
  std::vector<double> cell_coords;  // [nc][nv1][dim]
};

void number_dofs(const Mesh &M, Dofs &D);
// local P2 nodes (indices into the cell's n2 nodes) lying on local face f
int face_local_nodes(int dim, int f, int out[6]);

// Large host arrays: std::vector zero-fills on ONE thread, and for a fresh 1 GB allocation most of that time is the
// kernel handing out zeroed pages.  reserve_prefaulted() reserves the storage and lets all host threads populate its
// pages (madvise(MADV_POPULATE_WRITE): contents untouched), so the following resize() / assign() is a plain memset.
void prefault_parallel(void *p, size_t bytes);
template <typename T>
inline void reserve_prefaulted(std::vector<T> &v, size_t n)
{
  v.clear();
  v.reserve(n);
  prefault_parallel(v.data(), n * sizeof(T));
}

struct Csr {
  int n_rows = 0, n_cols = 0;
  std::vector<int> rowptr, colind;
  int64_t nnz() const { return rowptr.empty() ? 0 : rowptr.back(); }
};

// cells adjacent to every row index (node or pressure DoF), cells ascending: shared by the patterns with the same rows
struct RowCells {
  std::vector<int> ptr, cells;
  void build(int64_t nc, const int *cell_rows, int kr, int n_rows_total);
};
// rows r in [0, n_rows_owned): union over cells containing r of that cell's column entities.
// cell_rows[nc][kr], cell_cols[nc][kc]; columns sorted ascending.
void build_pattern(const RowCells &rc, const int *cell_cols, int kc, int n_rows_owned, int n_cols, Csr &out);
void build_pattern(int64_t nc, const int *cell_rows, int kr, const int *cell_cols, int kc, int n_rows_total,
                   int n_rows_owned, int n_cols, Csr &out);
// symbolic product pattern(A) * pattern(B), sorted columns
void symbolic_product(const Csr &A, const Csr &B, Csr &out);
inline int find_in_row(const Csr &A, int row, int col)
{
  int lo = A.rowptr[row], hi = A.rowptr[row + 1] - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1, c = A.colind[mid];
    if (c == col) return mid;
    if (c < col) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

void partition_cells_rcb(const Mesh &M, int nparts, std::vector<int> &part);

} // namespace nsb
