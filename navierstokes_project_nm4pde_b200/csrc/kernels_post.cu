// kernels_post.cu -- drag / lift on the device: the face loop of NavierStokes::compute_forces.
//
// Replaces (Navier-Stokes/src/NavierStokes2D.cpp:752-859, NavierStokes3D.cpp:744-840): FEFaceValues
// over the obstacle faces (boundary id 3) of locally owned cells, velocity gradients and pressure of
// `solution` at the face quadrature points, the two force formulas, and the MPI sum of the two
// doubles.  Without it every time step copied the whole solution to the host for an O(faces) sum.
//
// One thread per face.  The affine map gives the barycentric gradients (rows of J^-1), from which
// the outward normal (-grad lambda_opp normalised), the face measure (|det J| |grad lambda_opp|) and
// the P2 / P1 shape data at a face point follow in closed form; per-face results are reduced in a
// fixed order (bitwise reproducible), then summed over the ranks (halo_allreduce).
#include "nsb_internal.hpp"

namespace nsb {

struct ForceArgs {
  int nf, nq;
  const double *X;   // [(dim+1)*dim][nf]
  const int *nodes;  // [n2][nf]
  const int *pv;     // [nv1][nf]
  const int *opp;    // [nf]
  const double *q;   // xi[nq][dim-1], then w[nq]
  const double *sol; // local vector
  int n_nodes_owned, n_p_owned, ghost_off_u, p_base, ghost_off_p;
  double nu, rho;
  double *part;      // [2][nf]
};

template <int DIM>
__global__ void __launch_bounds__(128) k_face_forces(ForceArgs a)
{
  constexpr int NV = DIM + 1, N2 = DIM == 2 ? 6 : 10, NE = N2 - NV;
  constexpr int EA[6] = {0, 1, 2, 0, 1, 2}, EB[6] = {1, 2, 0, 3, 3, 3};
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= a.nf) return;
  double X[NV][DIM];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int d = 0; d < DIM; ++d) X[v][d] = a.X[size_t(v * DIM + d) * a.nf + f];
  // gl[v][d] = d lambda_v / d x_d : rows of J^-1 for v >= 1, minus their sum for v = 0
  double J[DIM][DIM], gl[NV][DIM], det;
#pragma unroll
  for (int r = 0; r < DIM; ++r)
#pragma unroll
    for (int k = 0; k < DIM; ++k) J[r][k] = X[k + 1][r] - X[0][r];
  if constexpr (DIM == 2) {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double id = 1.0 / det;
    gl[1][0] = J[1][1] * id;  gl[1][1] = -J[0][1] * id;
    gl[2][0] = -J[1][0] * id; gl[2][1] = J[0][0] * id;
  } else {
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const double id = 1.0 / det;
    gl[1][0] = c00 * id;
    gl[1][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    gl[1][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    gl[2][0] = c01 * id;
    gl[2][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    gl[2][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    gl[3][0] = c02 * id;
    gl[3][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    gl[3][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
  }
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    double s = 0.0;
#pragma unroll
    for (int v = 1; v < NV; ++v) s += gl[v][d];
    gl[0][d] = -s;
  }
  const int opp = a.opp[f];
  double gn = 0.0, g_opp[DIM];
#pragma unroll
  for (int v = 0; v < NV; ++v)
    if (v == opp) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) g_opp[d] = gl[v][d];
    }
#pragma unroll
  for (int d = 0; d < DIM; ++d) gn += g_opp[d] * g_opp[d];
  gn = sqrt(gn);
  // n = -(outward normal) = +grad lambda_opp / |grad lambda_opp|   (normal_vector = -fe_face_values.normal_vector(q))
  double n[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) n[d] = g_opp[d] / gn;
  const double meas = fabs(det) * gn; // |edge| (2D), 2 |triangle| (3D): the rule's weights sum to 1 resp. 1/2
  // nodal values
  double U[N2][DIM], P[NV];
#pragma unroll
  for (int i = 0; i < N2; ++i) {
    const int node = a.nodes[size_t(i) * a.nf + f];
    const int64_t at = int64_t(DIM) * node + (node >= a.n_nodes_owned ? a.ghost_off_u : 0);
#pragma unroll
    for (int d = 0; d < DIM; ++d) U[i][d] = a.sol[at + d];
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int p = a.pv[size_t(v) * a.nf + f];
    P[v] = a.sol[a.p_base + p + (p >= a.n_p_owned ? a.ghost_off_p : 0)];
  }
  const double *xi = a.q, *w = a.q + size_t(a.nq) * (DIM - 1);
  double drag = 0.0, lift = 0.0;
  for (int q = 0; q < a.nq; ++q) {
    // barycentrics of the face point: the face's vertices in ascending local order take (1 - s [- t], s [, t])
    double lf[DIM];
    if constexpr (DIM == 2) { lf[0] = 1.0 - xi[q]; lf[1] = xi[q]; }
    else { lf[0] = 1.0 - xi[2 * q] - xi[2 * q + 1]; lf[1] = xi[2 * q]; lf[2] = xi[2 * q + 1]; }
    double lam[NV];
    {
      int k = 0;
#pragma unroll
      for (int v = 0; v < NV; ++v) lam[v] = (v == opp) ? 0.0 : lf[k++];
    }
    double G[DIM][DIM], p = 0.0; // G[i][j] = d u_i / d x_j
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
      for (int j = 0; j < DIM; ++j) G[i][j] = 0.0;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double c = 4.0 * lam[v] - 1.0;
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j) G[i][j] += U[v][i] * c * gl[v][j];
      p += P[v] * lam[v];
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int ea = EA[e], eb = EB[e];
#pragma unroll
      for (int j = 0; j < DIM; ++j) {
        const double g = 4.0 * (lam[eb] * gl[ea][j] + lam[ea] * gl[eb][j]);
#pragma unroll
        for (int i = 0; i < DIM; ++i) G[i][j] += U[NV + e][i] * g;
      }
    }
    const double jxw = w[q] * meas;
    if constexpr (DIM == 2) {
      drag += (a.nu * (G[0][0] * n[0] + G[0][1] * n[1]) - p * n[0]) * jxw;
      lift += (a.nu * (G[1][0] * n[0] + G[1][1] * n[1]) - p * n[1]) * jxw;
    } else {
      const double nx = n[0], ny = n[1];
      const double t[3] = {ny, -nx, 0.0};
      const double t2 = t[0] * t[0] + t[1] * t[1];
      double ngt = 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) ngt += (n[0] * G[0][j] + n[1] * G[1][j] + n[2] * G[2][j]) * (t[j] / t2);
      drag += (a.rho * a.nu * ngt * ny - p * nx) * jxw;
      lift -= (a.rho * a.nu * ngt * nx + p * ny) * jxw;
    }
  }
  a.part[f] = drag;
  a.part[size_t(a.nf) + f] = lift;
}

// out[k] = sum_f part[k][f], k = 0, 1: one block, fixed summation order
__global__ void __launch_bounds__(256) k_force_reduce(int nf, const double *__restrict__ part, double *__restrict__ out)
{
  __shared__ double s[2][256];
  double a0 = 0.0, a1 = 0.0;
  for (int f = threadIdx.x; f < nf; f += 256) { a0 += part[f]; a1 += part[size_t(nf) + f]; }
  s[0][threadIdx.x] = a0; s[1][threadIdx.x] = a1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s[0][threadIdx.x] += s[0][threadIdx.x + o]; s[1][threadIdx.x] += s[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = s[0][0]; out[1] = s[1][0]; }
}

void force_faces_set(Handle &H, int nf, const int *face_cell, const int *face_opp, int nq, const double *xi, const double *w)
{
  const int dim = H.dim, nv1 = H.nv1, n2 = H.n2;
  if (nf < 0 || nq < 1 || nq > 64) throw ArgError("nsb_set_force_faces: bad face or quadrature count");
  std::vector<double> X(size_t(nv1) * dim * nf), q(size_t(nq) * dim);
  std::vector<int> nodes(size_t(n2) * nf), pv(size_t(nv1) * nf), opp(nf + size_t(0));
  for (int f = 0; f < nf; ++f) {
    const int64_t c = face_cell[f];
    if (c < 0 || c >= H.nc || face_opp[f] < 0 || face_opp[f] > dim) throw ArgError("nsb_set_force_faces: face out of range");
    for (int k = 0; k < nv1 * dim; ++k) X[size_t(k) * nf + f] = H.h_vcoords[c * nv1 * dim + k];
    for (int i = 0; i < n2; ++i) nodes[size_t(i) * nf + f] = H.h_cell_nodes[c * n2 + i];
    for (int v = 0; v < nv1; ++v) pv[size_t(v) * nf + f] = H.h_cell_p[c * nv1 + v];
    opp[f] = face_opp[f];
  }
  for (int k = 0; k < nq * (dim - 1); ++k) q[k] = xi[k];
  for (int k = 0; k < nq; ++k) q[size_t(nq) * (dim - 1) + k] = w[k];
  H.n_force_faces = nf;
  H.force_nq = nq;
  H.d_ff_x.upload(X); H.d_ff_nodes.upload(nodes); H.d_ff_p.upload(pv); H.d_ff_opp.upload(opp); H.d_ff_q.upload(q);
  H.d_ff_part.alloc(size_t(2) * std::max(nf, 1) + 2);
  NSB_CUDA(cudaDeviceSynchronize()); // uploads ran on the default stream, the kernels use H.stream
}

// drag, lift of the current solution summed over the ranks -> out[0..1] (host)
void force_faces_compute(Handle &H, double rho, double *out)
{
  if (H.force_nq == 0) throw StateError("nsb_compute_forces before nsb_set_force_faces");
  const int nf = H.n_force_faces;
  double *res = H.d_ff_part.p + size_t(2) * std::max(nf, 1);
  if (H.nranks > 1) { // faces of owned cells touch nodes owned by lower ranks: solution = solution_owned (ghost import)
    halo_exchange_u(H, H.d_sol.p, H.ghost_off_u());
    halo_exchange_p(H, H.d_sol.p + H.nu_owned(), H.ghost_off_p());
  }
  if (nf > 0) {
    ForceArgs a{nf, H.force_nq, H.d_ff_x.p, H.d_ff_nodes.p, H.d_ff_p.p, H.d_ff_opp.p, H.d_ff_q.p, H.d_sol.p,
                H.n_nodes_owned, H.n_p_owned, H.ghost_off_u(), H.p_base(), H.ghost_off_p(), H.prm.nu, rho, H.d_ff_part.p};
    if (H.dim == 2) k_face_forces<2><<<(nf + 127) / 128, 128, 0, H.stream>>>(a);
    else k_face_forces<3><<<(nf + 127) / 128, 128, 0, H.stream>>>(a);
    k_force_reduce<<<1, 256, 0, H.stream>>>(nf, H.d_ff_part.p, res);
    H.launches += 2;
    NSB_CUDA(cudaGetLastError());
  } else
    NSB_CUDA(cudaMemsetAsync(res, 0, 2 * sizeof(double), H.stream));
  halo_allreduce(H, res, 2);
  reduce_fetch(H, res, 2, out);
}

} // namespace nsb
