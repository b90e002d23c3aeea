// host_post.cpp -- host-side post-processing entry points of the C ABI (no GPU): the face integrals of
// NavierStokes::compute_forces.  O(boundary faces) work on the downloaded solution.
#include <cmath>

#include "../../include/nsb.h"
#include "fe_simplex.hpp"

using nsb::fe::Rule;

extern "C" {

// replaces: the face loop of NavierStokes::compute_forces
//   2D  src/NavierStokes2D.cpp:752-859: QGauss<1>(3), force = (nu grad u - p I) n, n = -outward normal
//   3D  src/NavierStokes3D.cpp:744-840: QGaussSimplex<2>(3), tangential formula with t = (n_y, -n_x, 0)
// over the boundary faces with the given id.  out[0] = drag, out[1] = lift (the raw integrals; the
// coefficient scaling stays with the caller).
int nsh_boundary_forces(nsh_mesh m, nsh_dofs d, const double *solution, int32_t boundary_id, double nu, double rho,
                        double *out)
{
  if (!m || !d || !solution || !out) return NSB_ERR_ARG;
  const int dim = nsh_mesh_dim(m);
  const int32_t nf = nsh_dofs_boundary_faces(d, m, boundary_id, nullptr, nullptr);
  std::vector<int32_t> fcell(size_t(nf > 0 ? nf : 1)), flocal(size_t(nf > 0 ? nf : 1));
  if (nf > 0) nsh_dofs_boundary_faces(d, m, boundary_id, fcell.data(), flocal.data());
  const Rule q = nsb::fe::gauss_simplex(dim - 1);
  const int32_t *cd = nsh_dofs_cell_dofs(d);
  const double *cc = nsh_dofs_cell_coords(d);
  const int nv = dim + 1, n2 = dim == 2 ? 6 : 10, dpc = nsh_dofs_per_cell(d);
  double drag = 0.0, lift = 0.0;
  for (int32_t f = 0; f < nf; ++f) {
    const int c = fcell[f], lf = flocal[f];
    const double *X = cc + size_t(c) * nv * dim;
    double gl[4][3];
    nsb::fe::bary_gradients(dim, X, gl);
    // outward unit normal of the face opposite to vertex lf: -grad(lambda_lf) normalised
    double nrm = 0.0, n_out[3] = {0, 0, 0};
    for (int k = 0; k < dim; ++k) nrm += gl[lf][k] * gl[lf][k];
    nrm = std::sqrt(nrm);
    for (int k = 0; k < dim; ++k) n_out[k] = -gl[lf][k] / nrm;
    int vs[3], k = 0;
    for (int v = 0; v < nv; ++v)
      if (v != lf) vs[k++] = v;
    double meas; // |edge| in 2D, 2 * area in 3D (weights of the reference face sum to 1 resp. 1/2)
    if (dim == 2) {
      meas = std::hypot(X[vs[1] * 2] - X[vs[0] * 2], X[vs[1] * 2 + 1] - X[vs[0] * 2 + 1]);
    } else {
      double e1[3], e2[3];
      for (int t = 0; t < 3; ++t) { e1[t] = X[vs[1] * 3 + t] - X[vs[0] * 3 + t]; e2[t] = X[vs[2] * 3 + t] - X[vs[0] * 3 + t]; }
      const double cr[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
      meas = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
    }
    // nodal values of this cell (FESystem order: per vertex [u.., p], then per edge [u..])
    double U[10][3], P[4];
    for (int v = 0; v < nv; ++v) {
      for (int t = 0; t < dim; ++t) U[v][t] = solution[cd[size_t(c) * dpc + v * (dim + 1) + t]];
      P[v] = solution[cd[size_t(c) * dpc + v * (dim + 1) + dim]];
    }
    for (int e = 0; e < n2 - nv; ++e)
      for (int t = 0; t < dim; ++t) U[nv + e][t] = solution[cd[size_t(c) * dpc + nv * (dim + 1) + e * dim + t]];
    for (int iq = 0; iq < q.size(); ++iq) {
      double lam[4] = {0, 0, 0, 0};
      if (dim == 2) { lam[vs[0]] = 1.0 - q.xi[iq]; lam[vs[1]] = q.xi[iq]; }
      else { lam[vs[0]] = 1.0 - q.xi[2 * iq] - q.xi[2 * iq + 1]; lam[vs[1]] = q.xi[2 * iq]; lam[vs[2]] = q.xi[2 * iq + 1]; }
      double phi[10], dphi[10][3];
      nsb::fe::shape_p2(dim, lam, gl, phi, dphi);
      double G[3][3] = {{0}}, p = 0.0; // G[i][j] = d u_i / d x_j
      for (int a = 0; a < n2; ++a)
        for (int i = 0; i < dim; ++i)
          for (int j = 0; j < dim; ++j) G[i][j] += U[a][i] * dphi[a][j];
      for (int v = 0; v < nv; ++v) p += P[v] * lam[v];
      const double jxw = q.w[iq] * meas;
      const double n[3] = {-n_out[0], -n_out[1], -n_out[2]}; // normal_vector = -fe_face_values.normal_vector(q)
      if (dim == 2) {
        double force[2];
        for (int i = 0; i < 2; ++i) force[i] = (nu * (G[i][0] * n[0] + G[i][1] * n[1]) - p * n[i]) * jxw;
        drag += force[0];
        lift += force[1];
      } else {
        const double nx = n[0], ny = n[1];
        const double t[3] = {ny, -nx, 0.0}, t2 = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
        double ngt = 0.0; // n * grad u * (t / |t|^2)
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) ngt += n[i] * G[i][j] * t[j] / t2;
        drag += (rho * nu * ngt * ny - p * nx) * jxw;
        lift -= (rho * nu * ngt * nx + p * ny) * jxw;
      }
    }
  }
  out[0] = drag;
  out[1] = lift;
  return NSB_OK;
}

} // extern "C"
