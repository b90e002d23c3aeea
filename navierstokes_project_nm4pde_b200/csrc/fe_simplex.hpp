// fe_simplex.hpp -- host-side P2-P1 simplex helpers shared by the library's post-processing entry
// points (host_post.cpp) and the C++ NavierStokes class (host/NavierStokes.cpp): quadrature tables,
// shape functions, affine geometry.  Header-only, no GPU.
#pragma once
#include <cmath>
#include <initializer_list>
#include <vector>

namespace nsb {
namespace fe {

// QGaussSimplex<dim>(3) as forwarded to Witherden-Vincent by deal.II >= 9.4 (7 / 14 points), and
// QGauss<1>(3); same tables as navierstokes_project_nm4pde_b200/quadrature.py.
struct Rule {
  std::vector<double> xi, w;
  int dim;
  int size() const { return int(w.size()); }
};

inline Rule gauss_simplex(int dim)
{
  Rule r;
  r.dim = dim;
  auto add = [&](std::initializer_list<double> p, double w) {
    for (double v : p) r.xi.push_back(v);
    r.w.push_back(w);
  };
  if (dim == 1) {
    const double g = std::sqrt(0.6);
    add({0.5 - 0.5 * g}, 5.0 / 18.0); add({0.5}, 8.0 / 18.0); add({0.5 + 0.5 * g}, 5.0 / 18.0);
  } else if (dim == 2) {
    const double s = std::sqrt(15.0);
    add({1.0 / 3.0, 1.0 / 3.0}, 0.1125);
    for (int k = 0; k < 2; ++k) {
      const double a = (k == 0 ? 6.0 - s : 6.0 + s) / 21.0, w = (k == 0 ? 155.0 - s : 155.0 + s) / 2400.0;
      add({a, a}, w); add({1.0 - 2.0 * a, a}, w); add({a, 1.0 - 2.0 * a}, w);
    }
  } else {
    const double A[2] = {0.31088591926330060980, 0.092735250310891226402};
    const double W[2] = {0.11268792571801585080 / 6.0, 0.073493043116361949544 / 6.0};
    for (int k = 0; k < 2; ++k) {
      const double a = A[k], b = 1.0 - 3.0 * a;
      add({a, a, a}, W[k]); add({b, a, a}, W[k]); add({a, b, a}, W[k]); add({a, a, b}, W[k]);
    }
    const double c = 0.045503704125649649492, d = 0.5 - c, w = 0.042546020777081466438 / 6.0;
    add({c, c, d}, w); add({c, d, c}, w); add({d, c, c}, w); add({c, d, d}, w); add({d, c, d}, w); add({d, d, c}, w);
  }
  return r;
}

static const int kEdges[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};

// P2 / P1 shape values and physical gradients at barycentric point `lam` of a simplex whose
// barycentric gradients are gl[v][d].
inline void shape_p2(int dim, const double *lam, const double gl[4][3], double *phi, double (*dphi)[3])
{
  const int nv = dim + 1, ne = dim == 2 ? 3 : 6;
  for (int v = 0; v < nv; ++v) {
    phi[v] = lam[v] * (2.0 * lam[v] - 1.0);
    for (int d = 0; d < dim; ++d) dphi[v][d] = (4.0 * lam[v] - 1.0) * gl[v][d];
  }
  for (int e = 0; e < ne; ++e) {
    const int a = kEdges[e][0], b = kEdges[e][1];
    phi[nv + e] = 4.0 * lam[a] * lam[b];
    for (int d = 0; d < dim; ++d) dphi[nv + e][d] = 4.0 * (lam[a] * gl[b][d] + lam[b] * gl[a][d]);
  }
}

// barycentric gradients and |det J| of the affine simplex X[v][d]
inline double bary_gradients(int dim, const double *X, double gl[4][3])
{
  double J[3][3] = {{0}}, Ji[3][3] = {{0}};
  for (int r = 0; r < dim; ++r)
    for (int k = 0; k < dim; ++k) J[r][k] = X[(k + 1) * dim + r] - X[r];
  double det;
  if (dim == 2) {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    Ji[0][0] = J[1][1] / det; Ji[0][1] = -J[0][1] / det; Ji[1][0] = -J[1][0] / det; Ji[1][1] = J[0][0] / det;
  } else {
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                 c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    Ji[0][0] = c00 / det; Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
    Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
    Ji[1][0] = c01 / det; Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
    Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
    Ji[2][0] = c02 / det; Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
    Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  }
  for (int d = 0; d < dim; ++d) {
    gl[0][d] = 0.0;
    for (int k = 0; k < dim; ++k) { gl[k + 1][d] = Ji[k][d]; gl[0][d] -= Ji[k][d]; }
  }
  return std::fabs(det);
}

} // namespace fe
} // namespace nsb
