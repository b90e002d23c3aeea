// NavierStokes.cpp -- the reference's `NavierStokes` class over the C ABI (see NavierStokes.hpp).
// Citations are relative to /root/reference/Navier-Stokes.
#include "NavierStokes.hpp"

#include "../fe_simplex.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <sys/stat.h>

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kEsA = kPi / 4.0, kEsB = kPi / 2.0, kEsNu = 1e-2; // Convergence3D.hpp:54-56

using nsb::fe::Rule;
using nsb::fe::gauss_simplex;
using nsb::fe::shape_p2;
using nsb::fe::bary_gradients;
using nsb::fe::kEdges;

// InletVelocity::vector_value (NavierStokes2D.hpp:26-44, NavierStokes3D.hpp:25-43)
double inlet_ux(int dim, const double *x, double t, int test_case)
{
  const double H = 0.41;
  if (dim == 2) {
    const double um = 1.5, y = x[1];
    if (test_case == 2) return 4.0 * um * y * (H - y) * std::sin(kPi * t / 8.0) / (H * H);
    if (test_case != 1) return 4.0 * um * y * (H - y) / (H * H);
    return 0.0;
  }
  const double um = 9.0, y = x[1], z = x[2];
  if (test_case == 3) return 16.0 * um * y * z * (H - z) * (H - y) * std::sin(kPi * t / 8.0) / (H * H * H * H);
  if (test_case != 1) return 16.0 * um * y * z * (H - z) * (H - y) / (H * H * H * H);
  return 0.0;
}

} // namespace

// InletVelocity::getMeanVelocity (NavierStokes2D.hpp:64-75, NavierStokes3D.hpp:64-75); the 2D class
// swaps cases 2 / 3 between profile and mean -- reproduced.
double inlet_mean_velocity(int dim, int test_case, double t)
{
  if (test_case == 1) return 0.0;
  if (dim == 2) return test_case == 3 ? 2.0 * 1.5 * std::sin(t * kPi / 8.0) / 3.0 : 2.0 * 1.5 / 3.0;
  return test_case == 3 ? 4.0 * 9.0 * std::sin(t * kPi / 8.0) / 9.0 : 4.0 * 9.0 / 9.0;
}

// ExactSolution::vector_value / gradient_tensor (Convergence3D.hpp:59-132)
void ethier_steinman(const double X[3], double t, double u[3], double &p, double g[3][3])
{
  const double a = kEsA, b = kEsB, x = X[0], y = X[1], z = X[2];
  const double e = -a * std::exp(-kEsNu * b * b * t);
  const double ex = std::exp(a * x), ey = std::exp(a * y), ez = std::exp(a * z);
  u[0] = e * (ex * std::sin(a * y + b * z) + ez * std::cos(a * x + b * y));
  u[1] = e * (ey * std::sin(a * z + b * x) + ex * std::cos(a * y + b * z));
  u[2] = e * (ez * std::sin(a * x + b * y) + ey * std::cos(a * z + b * x));
  const double f = -(a * a * std::exp(-2 * kEsNu * b * b * t)) / 2.0;
  p = f * (2.0 * std::sin(a * x + b * y) * std::cos(a * z + b * x) * std::exp(a * (y + z)) +
           2.0 * std::sin(a * y + b * z) * std::cos(a * x + b * y) * std::exp(a * (x + z)) +
           2.0 * std::sin(a * z + b * x) * std::cos(a * y + b * z) * std::exp(a * (x + y)) + std::exp(2 * a * x) +
           std::exp(2 * a * y) + std::exp(2 * a * z));
  g[0][0] = e * (a * ex * std::sin(a * y + b * z) - a * ez * std::sin(a * x + b * y));
  g[0][1] = e * (a * ex * std::cos(a * y + b * z) - b * ez * std::sin(a * x + b * y));
  g[0][2] = e * (b * ex * std::cos(a * y + b * z) + a * ez * std::cos(a * x + b * y));
  g[1][0] = e * (b * ey * std::cos(a * z + b * x) + a * ex * std::cos(a * y + b * z));
  g[1][1] = e * (a * ey * std::sin(a * z + b * x) - a * ex * std::sin(a * y + b * z));
  g[1][2] = e * (a * ey * std::cos(a * z + b * x) - b * ex * std::sin(a * y + b * z));
  g[2][0] = e * (a * ez * std::cos(a * x + b * y) - b * ey * std::sin(a * z + b * x));
  g[2][1] = e * (b * ez * std::cos(a * x + b * y) + a * ey * std::cos(a * z + b * x));
  g[2][2] = e * (a * ez * std::sin(a * x + b * y) - a * ey * std::sin(a * z + b * x));
}

NavierStokes::NavierStokes(Variant variant_, const std::string &mesh_file_name_, unsigned degree_velocity,
                           unsigned degree_pressure, double T_, double deltat_, int test_case_)
  : variant(variant_), dim(variant_ == Variant::Cylinder2D ? 2 : 3), mesh_file_name(mesh_file_name_), T(T_),
    deltat(deltat_), test_case(test_case_), nu(variant_ == Variant::Convergence3D ? 1e-2 : 1e-3)
{
  if (degree_velocity != 2 || degree_pressure != 1)
    throw std::invalid_argument("NavierStokes: only Taylor-Hood P2-P1 is on the device path");
}

NavierStokes::~NavierStokes()
{
  if (engine) nsb_destroy(engine);
  if (local) nsh_local_free(local);
  if (dofs) nsh_dofs_free(dofs);
  if (mesh) nsh_mesh_free(mesh);
}

void NavierStokes::check(int rc, const char *what) const
{
  if (rc >= 0) return;
  const char *msg = nsb_last_error(engine);
  throw std::runtime_error(std::string(what) + ": " + (msg ? msg : "error") + " (code " + std::to_string(rc) + ")");
}

void NavierStokes::set_parallel(int nranks_, int rank_, AllGather allgather_)
{
  if (engine) throw std::logic_error("NavierStokes::set_parallel must precede setup()");
  if (nranks_ < 1 || rank_ < 0 || rank_ >= nranks_) throw std::invalid_argument("NavierStokes::set_parallel: bad rank");
  if (nranks_ > 1 && !allgather_) throw std::invalid_argument("NavierStokes::set_parallel: an all-gather is required");
  nranks = nranks_;
  rank = rank_;
  allgather = std::move(allgather_);
  verbose = verbose && rank == 0; // pcout (NavierStokes2D.hpp:103)
}

const double *NavierStokes::node_xyz(int i) const { return nsh_dofs_node_xyz(dofs) + size_t(dim) * (node_gid ? node_gid[i] : i); }
const double *NavierStokes::p_xyz(int i) const { return nsh_dofs_p_xyz(dofs) + size_t(dim) * (p_gid ? p_gid[i] : i); }

// NavierStokes::setup (NavierStokes2D.cpp:2-157)
void NavierStokes::setup()
{
  if (verbose) std::cout << "Initializing the mesh" << std::endl;
  if (mesh_file_name.rfind("gen:", 0) == 0) {
    std::vector<std::string> tok;
    std::stringstream ss(mesh_file_name);
    for (std::string t; std::getline(ss, t, ':');) tok.push_back(t);
    auto num = [&](size_t i, int def) { return tok.size() > i ? std::atoi(tok[i].c_str()) : def; };
    if (tok.size() > 1 && tok[1] == "cylinder2d") mesh = nsh_mesh_cylinder2d(num(2, 1));
    else if (tok.size() > 1 && tok[1] == "cylinder3d") mesh = nsh_mesh_cylinder3d(num(2, 1), num(3, 3));
    else if (tok.size() > 1 && tok[1] == "cube") mesh = nsh_mesh_cube(num(2, 4));
  } else
    mesh = nsh_mesh_read_msh(mesh_file_name.c_str());
  if (!mesh) throw std::runtime_error("cannot read mesh " + mesh_file_name);
  if (nsh_mesh_dim(mesh) != dim) throw std::runtime_error("mesh dimension does not match the problem class");
  if (verbose) std::cout << "  Number of elements = " << nsh_mesh_n_cells(mesh) << std::endl;
  dofs = nsh_dofs_create(mesh);
  if (!dofs) throw std::runtime_error("DoF numbering failed");
  const int n_nodes_global = nsh_dofs_n_nodes(dofs), n_p_global = nsh_dofs_n_p(dofs);
  N_global = dim * n_nodes_global + n_p_global;
  dpc = nsh_dofs_per_cell(dofs);
  if (verbose)
    std::cout << "  Number of DoFs = " << N_global << " (" << dim * n_nodes_global << " + " << n_p_global << ")" << std::endl;

  // Partition (NavierStokes2D.cpp:16-19) and the locally relevant DoFs (:71-87); on one rank local == global.
  int n_nodes_owned, n_p_owned;
  const int32_t *g2l_node = nullptr;
  if (nranks > 1) {
    local = nsh_local_create(mesh, dofs, nranks, rank);
    if (!local) throw std::runtime_error("partitioning failed");
    n_cells = nsh_local_n_cells(local);
    n_nodes = nsh_local_n_nodes(local);
    n_p = nsh_local_n_p(local);
    n_nodes_owned = nsh_local_n_nodes_owned(local);
    n_p_owned = nsh_local_n_p_owned(local);
    cell_dofs = nsh_local_cell_dofs(local);
    cell_coords = nsh_local_cell_coords(local);
    node_gid = nsh_local_node_gid(local);
    p_gid = nsh_local_p_gid(local);
    cell_gid = nsh_local_cells(local);
    cell_owner = nsh_local_cell_part(local);
    g2l_cell = nsh_local_g2l_cell(local);
    g2l_node = nsh_local_g2l_node(local);
  } else {
    n_cells = nsh_mesh_n_cells(mesh);
    n_nodes = n_nodes_owned = n_nodes_global;
    n_p = n_p_owned = n_p_global;
    cell_dofs = nsh_dofs_cell_dofs(dofs);
    cell_coords = nsh_dofs_cell_coords(dofs);
  }
  n_u = dim * n_nodes;
  N = n_u + n_p;

  // Dirichlet nodes (interpolate_boundary_values, NavierStokes2D.cpp:328-353; Convergence3D.cpp:364-368); ghost
  // nodes included: their rows of B^T are cleared too
  auto boundary_nodes = [&](std::initializer_list<int32_t> ids) {
    std::vector<int32_t> idv(ids), out(size_t(nsh_dofs_boundary_nodes(dofs, mesh, idv.data(), int32_t(idv.size()), nullptr)));
    nsh_dofs_boundary_nodes(dofs, mesh, idv.data(), int32_t(idv.size()), out.data());
    return out;
  };
  std::vector<int32_t> dir_global;
  std::vector<char> inlet_global;
  if (variant == Variant::Convergence3D) {
    dir_global = boundary_nodes({0, 1, 2, 4, 5});
    inlet_global.assign(dir_global.size(), 0);
  } else {
    const std::vector<int32_t> inlet = boundary_nodes({0}), walls = boundary_nodes({2, 3});
    std::vector<char> on_inlet(size_t(n_nodes_global), 0), on_wall(size_t(n_nodes_global), 0);
    for (int32_t v : inlet) on_inlet[v] = 1;
    for (int32_t v : walls) on_wall[v] = 1;
    dir_global = inlet;
    for (int32_t v : walls)
      if (!on_inlet[v]) dir_global.push_back(v);
    for (int32_t v : dir_global) inlet_global.push_back(on_inlet[v] && !on_wall[v]); // the second call overwrites with zero
  }
  for (size_t k = 0; k < dir_global.size(); ++k) {
    const int32_t v = g2l_node ? g2l_node[dir_global[k]] : dir_global[k];
    if (v < 0) continue;
    dir_nodes.push_back(v);
    dir_is_inlet.push_back(inlet_global[k]);
  }
  for (int32_t v : dir_nodes)
    for (int c = 0; c < dim; ++c) dir_rows.push_back(dim * v + c);
  // faces with boundary id 3 (the obstacle; the Neumann face of the cube): those in local cells, flagged when
  // this rank owns the cell (cell->is_locally_owned(), NavierStokes2D.cpp:781)
  {
    const int32_t nf = nsh_dofs_boundary_faces(dofs, mesh, 3, nullptr, nullptr);
    std::vector<int32_t> fc(static_cast<size_t>(nf)), fl(static_cast<size_t>(nf));
    if (nf) nsh_dofs_boundary_faces(dofs, mesh, 3, fc.data(), fl.data());
    for (int32_t f = 0; f < nf; ++f) {
      const int32_t k = g2l_cell ? g2l_cell[fc[f]] : fc[f];
      if (k < 0) continue;
      obstacle_cells.push_back(k);
      obstacle_faces.push_back(fl[f]);
      obstacle_owned.push_back(!cell_owner || cell_owner[fc[f]] == rank);
    }
  }

  unsigned char nccl_id[128] = {0};
  if (nranks > 1) { // ncclGetUniqueId on rank 0, broadcast through the all-gather
    if (rank == 0 && nsb_get_unique_id(nccl_id) < 0) throw std::runtime_error(std::string("nsb_get_unique_id: ") + nsb_last_error(nullptr));
    std::vector<unsigned char> all(size_t(128) * nranks);
    allgather(nccl_id, all.data(), 128);
    std::copy(all.begin(), all.begin() + 128, nccl_id);
  }
  int rc = nsb_create(&engine, dim, device, nranks, rank, nranks > 1 ? nccl_id : nullptr);
  if (rc < 0) throw std::runtime_error(std::string("nsb_create: ") + nsb_last_error(nullptr));
  nsb_params prm;
  check(nsb_default_params(&prm, int(variant)), "nsb_default_params");
  prm.nu = nu;
  prm.deltat = deltat;
  prm.ilu_ordering = ilu_ordering;
  prm.orthogonalisation = orthogonalisation;
  prm.ilu_ordering_schur = ilu_ordering_schur;
  check(nsb_set_params(engine, &prm), "nsb_set_params");
  check(nsb_set_mesh(engine, n_cells, cell_coords, cell_dofs, n_u, n_p, dim * n_nodes_owned, n_p_owned), "nsb_set_mesh");
  const Rule q = gauss_simplex(dim); // QGaussSimplex<dim>(fe->degree + 1), NavierStokes2D.cpp:45
  check(nsb_set_quadrature(engine, q.size(), q.xi.data(), q.w.data()), "nsb_set_quadrature");
  if (nranks > 1) {
    const int32_t *nb, *snp, *sni, *rnc, *spp, *spi, *rpc;
    const int32_t nnb = nsh_local_halo(local, &nb, &snp, &sni, &rnc, &spp, &spi, &rpc);
    check(nsb_set_halo(engine, nnb, nb, snp, sni, rnc, spp, spi, rpc), "nsb_set_halo");
  }
  check(nsb_finalize_setup(engine), "nsb_finalize_setup");
  if (nranks > 1) { // peer-memory transport: every rank maps every rank's mailbox; all ranks or none (NSB_P2P=0: NCCL)
    transport_name = "nccl";
    const char *want = std::getenv("NSB_P2P");
    if (!(want && want[0] == '0') && nranks <= 16) {
      unsigned char mine[65] = {0}; // 64-byte handle + "export worked"
      mine[64] = nsb_p2p_export(engine, mine) >= 0;
      std::vector<unsigned char> all(size_t(65) * nranks), handles(size_t(64) * nranks);
      allgather(mine, all.data(), 65);
      bool ok = true;
      for (int r = 0; r < nranks; ++r) {
        ok = ok && all[size_t(65) * r + 64];
        std::copy(all.begin() + 65 * r, all.begin() + 65 * r + 64, handles.begin() + 64 * r);
      }
      unsigned char attached = ok && nsb_p2p_attach(engine, handles.data()) >= 0;
      std::vector<unsigned char> flags(size_t(nranks), 0);
      allgather(&attached, flags.data(), 1);
      const bool all_attached = std::all_of(flags.begin(), flags.end(), [](unsigned char f) { return f != 0; });
      if (attached && !all_attached) throw std::runtime_error("peer-memory transport came up on some ranks only");
      if (all_attached) transport_name = "p2p";
    }
    if (verbose) std::cout << "  " << nranks << " ranks, transport " << transport_name << std::endl;
  }
  check(nsb_set_dirichlet(engine, int32_t(dir_rows.size()), dir_rows.data()), "nsb_set_dirichlet");
  if (variant != Variant::Convergence3D) { // compute_forces runs on the device: obstacle faces of owned cells + face rule
    std::vector<int32_t> fc, fl;
    for (size_t f = 0; f < obstacle_cells.size(); ++f)
      if (obstacle_owned[f]) { fc.push_back(obstacle_cells[f]); fl.push_back(obstacle_faces[f]); }
    const Rule qf = gauss_simplex(dim - 1); // QGauss<1>(3) (NavierStokes2D.cpp:758) / QGaussSimplex<2>(3)
    check(nsb_set_force_faces(engine, int32_t(fc.size()), fc.data(), fl.data(), qf.size(), qf.xi.data(), qf.w.data()),
          "nsb_set_force_faces");
  }
  solution.assign(size_t(N), 0.0);
}

void NavierStokes::dirichlet_values(double time, std::vector<double> &vals) const
{
  vals.assign(dir_rows.size(), 0.0);
  for (size_t k = 0; k < dir_nodes.size(); ++k) {
    const double *x = node_xyz(dir_nodes[k]);
    if (variant == Variant::Convergence3D) {
      double u[3], p, g[3][3];
      ethier_steinman(x, time, u, p, g);
      for (int c = 0; c < 3; ++c) vals[3 * k + c] = u[c];
    } else if (dir_is_inlet[k])
      vals[size_t(dim) * k] = inlet_ux(dim, x, time, test_case);
  }
}

// Neumann face term of Convergence3D.cpp:309-330: sum_q h(x_q) . phi_i JxW on faces with id 3,
// h = nu du/dn - p n with n = +e_y (FunctionH, Convergence3D.hpp:159-175)
void NavierStokes::neumann_rhs(double time, std::vector<double> &rhs) const
{
  rhs.assign(size_t(n_u), 0.0);
  const Rule q = gauss_simplex(2);
  const int32_t *cd = cell_dofs; // every face touching an owned node lies in a local cell (two-layer halo)
  const double *cc = cell_coords;
  for (size_t f = 0; f < obstacle_cells.size(); ++f) {
    const int c = obstacle_cells[f], lf = obstacle_faces[f];
    int vs[3], nv = 0;
    for (int v = 0; v < 4; ++v)
      if (v != lf) vs[nv++] = v;
    const double *X = cc + size_t(c) * 12;
    double e1[3], e2[3];
    for (int d = 0; d < 3; ++d) { e1[d] = X[vs[1] * 3 + d] - X[vs[0] * 3 + d]; e2[d] = X[vs[2] * 3 + d] - X[vs[0] * 3 + d]; }
    const double cr[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const double area2 = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
    for (int iq = 0; iq < q.size(); ++iq) {
      double lam[4] = {0, 0, 0, 0};
      lam[vs[0]] = 1.0 - q.xi[2 * iq] - q.xi[2 * iq + 1];
      lam[vs[1]] = q.xi[2 * iq];
      lam[vs[2]] = q.xi[2 * iq + 1];
      double xq[3] = {0, 0, 0};
      for (int v = 0; v < 4; ++v)
        for (int d = 0; d < 3; ++d) xq[d] += lam[v] * X[v * 3 + d];
      double u[3], p, g[3][3];
      ethier_steinman(xq, time, u, p, g);
      double h[3] = {kEsNu * g[0][1], kEsNu * g[1][1] - p, kEsNu * g[2][1]};
      const double jxw = q.w[iq] * area2;
      // P2 shape values (only nodes on the face are non-zero there)
      for (int v = 0; v < 4; ++v) {
        const double ph = lam[v] * (2.0 * lam[v] - 1.0);
        if (ph == 0.0) continue;
        const int node = cd[size_t(c) * dpc + v * 4] / 3;
        for (int d = 0; d < 3; ++d) rhs[size_t(3) * node + d] += h[d] * ph * jxw;
      }
      for (int e = 0; e < 6; ++e) {
        const double ph = 4.0 * lam[kEdges[e][0]] * lam[kEdges[e][1]];
        if (ph == 0.0) continue;
        const int node = cd[size_t(c) * dpc + 16 + e * 3] / 3;
        for (int d = 0; d < 3; ++d) rhs[size_t(3) * node + d] += h[d] * ph * jxw;
      }
    }
  }
}

void NavierStokes::initial_condition(std::vector<double> &x) const
{
  x.assign(size_t(N), 0.0); // FunctionU0 = 0 for the cylinders (NavierStokes2D.hpp:140-150)
  if (variant != Variant::Convergence3D) return;
  double u[3], p, g[3][3];
  for (int n = 0; n < n_nodes; ++n) {
    ethier_steinman(node_xyz(n), 0.0, u, p, g);
    for (int c = 0; c < 3; ++c) x[size_t(3) * n + c] = u[c];
  }
  for (int v = 0; v < n_p; ++v) {
    ethier_steinman(p_xyz(v), 0.0, u, p, g);
    x[size_t(n_u) + v] = p;
  }
}

void NavierStokes::assemble(const double &time)
{
  if (verbose) std::cout << "===============================================\nAssembling the system" << std::endl;
  std::vector<double> vals;
  dirichlet_values(time, vals);
  check(nsb_set_dirichlet_values(engine, vals.data()), "nsb_set_dirichlet_values");
  check(nsb_assemble_first(engine, time), "nsb_assemble_first");
}

void NavierStokes::assemble_time_step(const double &time)
{
  std::vector<double> vals;
  dirichlet_values(time, vals);
  check(nsb_set_dirichlet_values(engine, vals.data()), "nsb_set_dirichlet_values");
  check(nsb_assemble_step(engine, time), "nsb_assemble_step");
}

void NavierStokes::solve_time_step(double)
{
  int32_t its = 0;
  double tp = 0, ts = 0;
  check(nsb_solve_step(engine, &its, &tp, &ts), "nsb_solve_step");
  time_prec.push_back(tp);
  time_solve.push_back(ts);
  gmres_iterations.push_back(its);
  if (verbose) {
    std::cout << "Time taken to initialize preconditioner: " << tp << " seconds" << std::endl;
    std::cout << "Time taken to solve Navier Stokes problem: " << ts << " seconds" << std::endl;
    std::cout << "Result:  " << its << " GMRES iterations" << std::endl;
  }
  if (write_output && rank == 0 && variant == Variant::Cylinder2D) { // gmres.csv: time, Re, iterations (NavierStokes2D.cpp:622-636)
    const int Re = int(0.1 * 1.5 * std::sin(time_now * kPi / 8.0) / .001);
    std::ofstream gm("gmres.csv", std::ios::app);
    if (gm.is_open()) gm << time_now << ',' << Re << ',' << its << "\n";
  }
  solution_stale = true; // solution = solution_owned (:637) stays on the device until a host consumer asks
}

void NavierStokes::sync_solution() const
{
  if (!solution_stale) return;
  check(nsb_get_solution(engine, solution.data()), "nsb_get_solution");
  solution_stale = false;
}

// NavierStokes2D.cpp:752-859 (QGauss<1>(3) on the cylinder edges, force = (nu grad u - p I) n) and
// NavierStokes3D.cpp:744-840 (QGaussSimplex<2>(3), tangential formula).
std::vector<double> NavierStokes::compute_forces()
{
  double fl[2]; // face integrals + sum over ranks on the device: no solution download
  check(nsb_compute_forces(engine, rho, fl), "nsb_compute_forces");
  const double drag = fl[0], lift = fl[1];
  const double mean_v = inlet_mean_velocity(dim, test_case, time_now), D = 0.1, H = 0.41;
  const double den = dim == 2 ? mean_v * mean_v * D : rho * mean_v * mean_v * D * H;
  const double c_d = 2.0 * drag / den, c_l = 2.0 * lift / den;
  if (verbose) std::cout << "Coeff:\t " << c_d << " Coeff:\t " << c_l << std::endl;
  coefficients_history.push_back({c_d, c_l});
  return {c_d, c_l};
}

// NavierStokes::output (NavierStokes2D.cpp:642-695, NavierStokes3D.cpp:643-692): one .vtu per call in
// ./output2D_1/ (2D) or ./outputConvergence/ (3D and CONV, as the reference), and in 2D the
// coefficients appended to coeff_2.csv.
void NavierStokes::output(unsigned time_step, const std::vector<double> &coeff) const
{
  const std::string dir = variant == Variant::Cylinder2D ? "./output2D_1/" : "./outputConvergence/";
  const std::string name = variant == Variant::Cylinder2D ? "output-navier-stokes-2D" : "output-navier-stokes-3D";
  mkdir(dir.c_str(), 0755);
  char num[16];
  std::snprintf(num, sizeof(num), "%03u", time_step);
  const std::string path = dir + name + "_" + num + ".vtu";
  if (nranks == 1) {
    sync_solution();
    if (nsh_write_vtu(mesh, dofs, solution.data(), path.c_str()) != 0) throw std::runtime_error("cannot write " + path);
    if (verbose) std::cout << "Output written to " << name << std::endl;
  } else if (verbose && time_step == 0)
    std::cout << "(.vtu pieces are not written in multi-rank runs: output is outside the device path)" << std::endl;
  if (rank == 0 && variant == Variant::Cylinder2D && coeff.size() >= 2) {
    std::ofstream coeff_file("coeff_2.csv", std::ios::app); // append mode as the reference (:682)
    if (coeff_file.is_open()) coeff_file << time_step << "," << coeff[0] << "," << coeff[1] << "\n";
  }
}

// NavierStokes2D.cpp:862-936: pressure at A = (0.45, 0.2[, 0.205]) minus pressure at E = (0.55, 0.2[, 0.205])
// through VectorTools::point_value; a point outside the mesh contributes 0 as in the reference.
void NavierStokes::compute_pressure_difference()
{
  const double pa[3] = {0.45, 0.2, 0.205}, pe[3] = {0.55, 0.2, 0.205};
  double va[4], ve[4];
  const double p1 = point_value(pa, va) ? va[dim] : 0.0;
  const double p2 = point_value(pe, ve) ? ve[dim] : 0.0;
  pressure_difference = p1 - p2;
  if (verbose) std::cout << "Pressure difference (P(A) - P(B)) = " << pressure_difference << std::endl;
}

// VectorTools::point_value on the distributed solution: the rank that owns the cell holding x evaluates it from
// its local vector, the others contribute zero, and the values are summed over the ranks (collective).
bool NavierStokes::point_value(const double *x, double *out) const
{
  sync_solution();
  if (nranks == 1) return nsh_dofs_point_value(dofs, solution.data(), x, out) == 0;
  double lam[4], acc[5] = {0, 0, 0, 0, 0}; // dim + 1 values and "found"
  const int32_t c = nsh_dofs_find_cell(dofs, x, lam);
  if (c >= 0 && cell_owner[c] == rank) {
    const int32_t *cd = cell_dofs + size_t(g2l_cell[c]) * dpc;
    const int nv1 = dim + 1, n2 = dim == 2 ? 6 : 10;
    for (int v = 0; v < nv1; ++v) {
      const double ph = lam[v] * (2.0 * lam[v] - 1.0);
      for (int k = 0; k < dim; ++k) acc[k] += ph * solution[cd[v * nv1 + k]];
      acc[dim] += lam[v] * solution[cd[v * nv1 + dim]];
    }
    for (int e = 0; e < n2 - nv1; ++e) {
      const double ph = 4.0 * lam[kEdges[e][0]] * lam[kEdges[e][1]];
      for (int k = 0; k < dim; ++k) acc[k] += ph * solution[cd[nv1 * nv1 + e * dim + k]];
    }
    acc[4] = 1.0;
  }
  check(nsb_allreduce_sum(engine, acc, 5), "nsb_allreduce_sum");
  for (int k = 0; k <= dim; ++k) out[k] = acc[k];
  return acc[4] > 0.5;
}

// NavierStokes::solve (NavierStokes2D.cpp:699-750, NavierStokes3D.cpp:694-742, Convergence3D.cpp:724-764)
void NavierStokes::solve()
{
  if (verbose) std::cout << "===============================================\nApplying the initial condition" << std::endl;
  initial_condition(solution);
  check(nsb_set_solution(engine, solution.data()), "nsb_set_solution");
  if (write_output) output(0, {0.0, 0.0}); // the initial solution (:712)
  double c_D_max = -999, c_L_min = 999, time = 0;
  unsigned time_step = 0;
  std::vector<double> neu;
  while (time < T - 0.5 * deltat) {
    if (variant == Variant::Convergence3D) { // function_h.set_time(time) BEFORE the increment (Convergence3D.cpp:747-750)
      neumann_rhs(time, neu);
      check(nsb_set_neumann_rhs(engine, neu.data()), "nsb_set_neumann_rhs");
    }
    time += deltat;
    ++time_step;
    time_now = time;
    if (verbose) std::cout << "n = " << std::setw(3) << time_step << ", t = " << std::setw(5) << time << ":" << std::flush;
    if (time == deltat) assemble(time);
    else assemble_time_step(time);
    solve_time_step(time);
    if (variant != Variant::Convergence3D && time == T - deltat) compute_pressure_difference(); // NavierStokes2D.cpp:735
    const bool forces = variant == Variant::Cylinder2D || (variant == Variant::Cylinder3D && time > forces_after);
    std::vector<double> c = {0.0, 0.0};
    if (forces) {
      c = compute_forces();
      c_D_max = std::max(c_D_max, c[0]);
      c_L_min = std::min(c_L_min, c[1]);
    }
    // VTU every step in 2D and CONV (:743, Convergence3D.cpp:761), every 20 steps in 3D (NavierStokes3D.cpp:734)
    if (write_output && (variant != Variant::Cylinder3D || time_step % 20 == 0)) output(time_step, c);
    if (max_steps > 0 && int(time_step) >= max_steps) break;
  }
  if (verbose && variant != Variant::Convergence3D) {
    std::cout << "===============================================\nDrag Coefficient Max ----->   " << c_D_max << "\n\n"
              << "Lift Coefficient Min ----->   " << c_L_min << "\n===============================================" << std::endl;
  }
}

// Convergence3D.cpp:766-794: VectorTools::integrate_difference over the velocity components against
// the exact solution at t = T.  The reference integrates with QGaussSimplex<3>(degree + 2); the
// 14-point degree-5 rule is used here (the integrand error is far below the discretisation error).
double NavierStokes::compute_error(const VectorTools::NormType &norm_type)
{
  sync_solution();
  const Rule q = gauss_simplex(3);
  const int32_t *cd = cell_dofs;
  const double *cc = cell_coords;
  double e2 = 0.0, h2 = 0.0;
  for (int c = 0; c < n_cells; ++c) {
    if (cell_owner && cell_owner[cell_gid[c]] != rank) continue; // locally owned cells, then the sum over ranks
    const double *X = cc + size_t(c) * 12;
    double gl[4][3];
    const double det = bary_gradients(3, X, gl);
    double U[10][3];
    for (int v = 0; v < 4; ++v)
      for (int d = 0; d < 3; ++d) U[v][d] = solution[cd[size_t(c) * dpc + v * 4 + d]];
    for (int e = 0; e < 6; ++e)
      for (int d = 0; d < 3; ++d) U[4 + e][d] = solution[cd[size_t(c) * dpc + 16 + e * 3 + d]];
    for (int iq = 0; iq < q.size(); ++iq) {
      const double lam[4] = {1.0 - q.xi[3 * iq] - q.xi[3 * iq + 1] - q.xi[3 * iq + 2], q.xi[3 * iq], q.xi[3 * iq + 1],
                             q.xi[3 * iq + 2]};
      double xq[3] = {0, 0, 0};
      for (int v = 0; v < 4; ++v)
        for (int d = 0; d < 3; ++d) xq[d] += lam[v] * X[v * 3 + d];
      double phi[10], dphi[10][3], uh[3] = {0, 0, 0}, gh[3][3] = {{0}};
      shape_p2(3, lam, gl, phi, dphi);
      for (int a = 0; a < 10; ++a)
        for (int i = 0; i < 3; ++i) {
          uh[i] += U[a][i] * phi[a];
          for (int j = 0; j < 3; ++j) gh[i][j] += U[a][i] * dphi[a][j];
        }
      double ue[3], pe, ge[3][3];
      ethier_steinman(xq, T, ue, pe, ge);
      const double jxw = q.w[iq] * det;
      for (int i = 0; i < 3; ++i) {
        e2 += jxw * (uh[i] - ue[i]) * (uh[i] - ue[i]);
        for (int j = 0; j < 3; ++j) h2 += jxw * (gh[i][j] - ge[i][j]) * (gh[i][j] - ge[i][j]);
      }
    }
  }
  double sums[2] = {e2, h2}; // Utilities::MPI::sum of the squared cell errors (Convergence3D.cpp:785-790)
  check(nsb_allreduce_sum(engine, sums, 2), "nsb_allreduce_sum");
  return norm_type == VectorTools::L2_norm ? std::sqrt(sums[0]) : std::sqrt(sums[0] + sums[1]);
}
