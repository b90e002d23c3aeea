// solver.cu -- Krylov solvers and block preconditioners driving the device kernels.
//
// Replaces (all on device vectors; the host only sees the scalars it needs for stopping tests):
//   SolverGMRES<BlockVector>::solve(system_matrix, solution_owned, system_rhs, P)
//                                     Navier-Stokes/src/NavierStokes2D.cpp:559,575,592,609
//   inner SolverGMRES / SolverCG + PreconditionILU      include/Preconditioners.hpp:157-182, 271-289, 371-405
//   PreconditionSIMPLE / aSIMPLE / Yosida / aYosida     include/Preconditioners.hpp:118-534
// deal.II's algorithms are restated (left preconditioning, 30 temporary vectors => restart 28,
// modified Gram-Schmidt with the Kelley re-orthogonalisation test every 5th vector, Givens QR,
// stopping on the preconditioned residual; CG in the g = Ax - b formulation).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <functional>

#include "nsb_internal.hpp"

namespace nsb {

void solver_alloc(Handle &H)
{
  solver_free(H);
  auto *w = new Handle::SolverWs();
  H.ws = w;
  const size_t nl = size_t(H.n_local());
  const size_t nu = size_t(H.dim) * H.n_nodes, np = size_t(H.n_p);
  const int nt = H.prm.gmres_tmp;
  w->V_outer.alloc(nl * nt);
  w->V_inner.alloc(std::max(nu, np) * nt);
  w->V_inner.zero();
  for (auto &b : w->tu) b.alloc(nu);
  for (auto &b : w->tp) b.alloc(std::max<size_t>(np, 1));
  w->scal.alloc(256);
  w->scal.zero();
  w->prec_in.alloc(nl);
  w->prec_out.alloc(nl);
}

void solver_free(Handle &H)
{
  delete H.ws;
  H.ws = nullptr;
}

void reduce_fetch(Handle &H, double *dev, int n, double *host_out)
{
  NSB_CUDA(cudaMemcpyAsync(H.h_pinned, dev, sizeof(double) * n, cudaMemcpyDeviceToHost, H.stream));
  NSB_CUDA(cudaStreamSynchronize(H.stream));
  H.cnt_sync++;
  for (int i = 0; i < n; ++i) host_out[i] = H.h_pinned[i];
}

// dot product over the owned entries, summed over ranks, result left on the device
static void dot_dev(Handle &H, int n, const double *x, const double *y, double *out)
{
  H.cnt_dot++;
  if (halo_is_p2p(H)) { vec_dot_dev(H, n, x, y, out, halo_ar_args(H)); return; } // all-reduce inside the kernel
  vec_dot_dev(H, n, x, y, out);
  if (H.nranks > 1) halo_allreduce(H, out, 1);
}
static void add_and_dot_dev(Handle &H, int n, double *vv, const double *a, double sign, const double *vp,
                            const double *vn, double *out)
{
  H.cnt_dot++;
  if (halo_is_p2p(H)) { vec_add_and_dot_dev(H, n, vv, a, sign, vp, vn, out, halo_ar_args(H)); return; }
  vec_add_and_dot_dev(H, n, vv, a, sign, vp, vn, out);
  if (H.nranks > 1) halo_allreduce(H, out, 1);
}
static double norm_host(Handle &H, int n, const double *x, double *slot)
{
  dot_dev(H, n, x, x, slot);
  double v;
  reduce_fetch(H, slot, 1, &v);
  return std::sqrt(v);
}

// --------------------------------------------------------------------------------------------
// SolverControl
// --------------------------------------------------------------------------------------------
struct Control {
  int maxit;
  double tol;
  int last_step = 0;
  double last_value = 0;
  int check(int step, double value)
  { // 0 iterate, 1 success, 2 failure
    last_step = step;
    last_value = value;
    if (value <= tol) return 1;
    if (step >= maxit || std::isnan(value)) return 2;
    return 0;
  }
};

struct Space {
  int n_owned;  // entries taking part in reductions / updates
  int ld;       // stride between Krylov basis vectors
  int goff_sol; // ghost offset of the caller's solution vector (passed through to A)
  // y = A x ; goff = ghost offset of x
  std::function<void(const double *x, int goff, double *y)> A;
  std::function<void(const double *r, double *z)> P;
};

static void givens_rotation(double *h, double *b, double *ci, double *si, int col)
{
  for (int i = 0; i < col; ++i) {
    const double s = si[i], c = ci[i], dummy = h[i];
    h[i] = c * dummy + s * h[i + 1];
    h[i + 1] = -s * dummy + c * h[i + 1];
  }
  const double r = 1.0 / std::sqrt(h[col] * h[col] + h[col + 1] * h[col + 1]);
  si[col] = h[col + 1] * r;
  ci[col] = h[col] * r;
  h[col] = ci[col] * h[col] + si[col] * h[col + 1];
  b[col + 1] = -si[col] * b[col];
  b[col] *= ci[col];
}

// SolverGMRES::solve.  V: n_tmp basis vectors (stride sp.ld; every vector is fully written before it
// is read, so the workspace needs no clearing between solves); scal: >= 64 device doubles.
// orth: 0 = modified Gram-Schmidt as in deal.II; 1 = batched classical Gram-Schmidt with deal.II's
// loss-of-orthogonality test applied to every vector; 2 = batched classical Gram-Schmidt, always two
// passes (CGS2).
static int gmres(Handle &H, Space &sp, double *x, const double *b, double *V, int n_tmp, double *scal, Control &ctl,
                 int orth = 0)
{
  const int n = sp.n_owned;
  std::vector<double> Hm(size_t(n_tmp) * (n_tmp - 1), 0.0), gamma(n_tmp + 1, 0.0), ci(n_tmp, 0.0), si(n_tmp, 0.0),
      h(n_tmp + 2, 0.0), hh(std::max(n_tmp + 4, 64), 0.0);
  const int ldh = n_tmp - 1;
  double *v = V;
  double *p = V + size_t(n_tmp - 1) * sp.ld;
  int accumulated = 0, state = 0, dim = 0;
  bool re_orth = false;
  do {
    std::fill(h.begin(), h.end(), 0.0);
    sp.A(x, sp.goff_sol, p);
    vec_sadd(H, n, -1.0, 1.0, b, p); // p.sadd(-1, 1, b)
    sp.P(p, v);
    double rho = norm_host(H, n, v, scal);
    state = ctl.check(accumulated, rho);
    if (state != 0) break;
    gamma[0] = rho;
    vec_scale(H, n, 1.0 / rho, v);
    dim = 0;
    for (int inner = 0; inner < n_tmp - 2 && state == 0; ++inner) {
      ++accumulated;
      double *vv = V + size_t(inner + 1) * sp.ld;
      sp.A(V + size_t(inner) * sp.ld, 0, p);
      sp.P(p, vv);
      dim = inner + 1;
      if (orth != 0) {
        // batched classical Gram-Schmidt: hd[0..dim-1] = V^T vv and hd[dim] = |vv|^2 from one fused
        // multi-dot (one all-reduce), vv -= V hd with the new |vv|^2 in hd[dim+1] (one all-reduce)
        double *hd = scal, *hd2 = scal + 32;
        const bool fused = halo_is_p2p(H); // peer-memory transport: the reductions all-reduce themselves
        vec_multi_dot_dev(H, n, vv, V, size_t(sp.ld), dim, hd, hd + dim, fused);
        H.cnt_dot++;
        if (H.nranks > 1 && !fused) halo_allreduce(H, hd, dim + 1);
        vec_multi_axpy_dev(H, n, vv, V, size_t(sp.ld), dim, hd, orth == 2 ? nullptr : hd + dim + 1, fused);
        bool second = (orth == 2);
        if (!second) {
          if (H.nranks > 1 && !fused) halo_allreduce(H, hd + dim + 1, 1);
          reduce_fetch(H, hd, dim + 2, hh.data());
          for (int i = 0; i < dim; ++i) h[i] = hh[i];
          // deal.II's test (solver_gmres.h, modified_gram_schmidt) on every vector
          second = !(std::sqrt(hh[dim + 1]) > 10.0 * std::sqrt(hh[dim]) * std::sqrt(2.220446049250313e-16));
        }
        double s2 = hh[dim + 1];
        if (second) {
          vec_multi_dot_dev(H, n, vv, V, size_t(sp.ld), dim, hd2, nullptr, fused);
          H.cnt_dot++;
          if (H.nranks > 1 && !fused) halo_allreduce(H, hd2, dim);
          vec_multi_axpy_dev(H, n, vv, V, size_t(sp.ld), dim, hd2, hd2 + dim, fused);
          if (H.nranks > 1 && !fused) halo_allreduce(H, hd2 + dim, 1);
          if (orth == 2) {
            reduce_fetch(H, scal, 64, hh.data());
            for (int i = 0; i < dim; ++i) h[i] = hh[i] + hh[32 + i];
            s2 = hh[32 + dim];
          } else {
            reduce_fetch(H, hd2, dim + 1, hh.data());
            for (int i = 0; i < dim; ++i) h[i] += hh[i];
            s2 = hh[dim];
          }
        }
        const double s = std::sqrt(s2);
        h[inner + 1] = s;
        if (s != 0.0) vec_scale(H, n, 1.0 / s, vv);
        givens_rotation(h.data(), gamma.data(), ci.data(), si.data(), inner);
        for (int i = 0; i < dim; ++i) Hm[size_t(i) * ldh + inner] = h[i];
        rho = std::fabs(gamma[dim]);
        state = ctl.check(accumulated, rho);
        continue;
      }
      // modified_gram_schmidt: all coefficients stay on the device until the single fetch below
      const bool consider = (!re_orth) && (inner % 5 == 4);
      double *hd = scal;              // hd[0..dim-1] coefficients, hd[dim] = |vv|^2 after orthogonalisation
      double *nstart = scal + n_tmp + 1;
      if (consider) dot_dev(H, n, vv, vv, nstart);
      dot_dev(H, n, vv, V, hd);
      for (int i = 1; i < dim; ++i)
        add_and_dot_dev(H, n, vv, hd + i - 1, -1.0, V + size_t(i - 1) * sp.ld, V + size_t(i) * sp.ld, hd + i);
      add_and_dot_dev(H, n, vv, hd + dim - 1, -1.0, V + size_t(dim - 1) * sp.ld, vv, hd + dim);
      reduce_fetch(H, scal, n_tmp + 2, hh.data());
      for (int i = 0; i < dim; ++i) h[i] = hh[i];
      double s = std::sqrt(hh[dim]);
      bool done = false;
      if (consider) {
        const double norm_vv_start = std::sqrt(hh[n_tmp + 1]);
        if (s > 10.0 * norm_vv_start * std::sqrt(2.220446049250313e-16)) done = true;
        else re_orth = true;
      }
      if (!done && re_orth) {
        dot_dev(H, n, vv, V, hd);
        for (int i = 1; i < dim; ++i)
          add_and_dot_dev(H, n, vv, hd + i - 1, -1.0, V + size_t(i - 1) * sp.ld, V + size_t(i) * sp.ld, hd + i);
        add_and_dot_dev(H, n, vv, hd + dim - 1, -1.0, V + size_t(dim - 1) * sp.ld, vv, hd + dim);
        reduce_fetch(H, scal, n_tmp + 2, hh.data());
        for (int i = 0; i < dim; ++i) h[i] += hh[i];
        s = std::sqrt(hh[dim]);
      }
      h[inner + 1] = s;
      if (s != 0.0) vec_scale(H, n, 1.0 / s, vv);
      givens_rotation(h.data(), gamma.data(), ci.data(), si.data(), inner);
      for (int i = 0; i < dim; ++i) Hm[size_t(i) * ldh + inner] = h[i];
      rho = std::fabs(gamma[dim]);
      state = ctl.check(accumulated, rho);
    }
    for (int i = dim - 1; i >= 0; --i) { // H1.backward(h, gamma)
      double s = gamma[i];
      for (int j = i + 1; j < dim; ++j) s -= Hm[size_t(i) * ldh + j] * h[j];
      h[i] = s / Hm[size_t(i) * ldh + i];
    }
    for (int i = 0; i < dim; ++i) vec_axpy(H, n, h[i], V + size_t(i) * sp.ld, x);
  } while (state == 0);
  return state == 1 ? 0 : -1;
}

// SolverCG::solve with device-resident alpha / beta
__global__ void k_cg_update(int n, const double *__restrict__ gh, const double *__restrict__ dh,
                            const double *__restrict__ d, const double *__restrict__ h, double *__restrict__ x,
                            double *__restrict__ g)
{
  const double alpha = (*gh) / (*dh);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    x[i] += alpha * d[i];
    g[i] += alpha * h[i];
  }
}
__global__ void k_cg_dir(int n, const double *__restrict__ gh_new, const double *__restrict__ gh_old,
                         const double *__restrict__ h, double *__restrict__ d)
{
  const double beta = (*gh_new) / (*gh_old);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) d[i] = beta * d[i] - h[i];
}

static int cg(Handle &H, Space &sp, double *x, const double *b, double *g, double *d, double *h, double *scal,
              Control &ctl)
{
  const int n = sp.n_owned;
  const unsigned grid = unsigned(std::max(1, std::min((n + 255) / 256, 148 * 8)));
  double *gh[2] = {scal, scal + 1};
  double *dh = scal + 2, *rr = scal + 3;
  int it = 0, cur = 0;
  sp.A(x, sp.goff_sol, g);
  vec_axpy(H, n, -1.0, b, g); // g = A x - b  (identical to g.equ(-1, b) when x == 0)
  double res = norm_host(H, n, g, rr);
  int state = ctl.check(0, res);
  if (state == 0) {
    sp.P(g, h);
    vec_copy(H, n, h, d);
    vec_scale(H, n, -1.0, d);
    dot_dev(H, n, g, h, gh[cur]);
    while (state == 0) {
      ++it;
      sp.A(d, 0, h);
      dot_dev(H, n, d, h, dh);
      k_cg_update<<<grid, 256, 0, H.stream>>>(n, gh[cur], dh, d, h, x, g);
      H.launches++;
      res = norm_host(H, n, g, rr);
      state = ctl.check(it, res);
      if (state != 0) break;
      sp.P(g, h);
      dot_dev(H, n, g, h, gh[cur ^ 1]);
      k_cg_dir<<<grid, 256, 0, H.stream>>>(n, gh[cur ^ 1], gh[cur], h, d);
      H.launches++;
      cur ^= 1;
    }
  }
  return state == 1 ? 0 : -1;
}

// --------------------------------------------------------------------------------------------
// operators
// --------------------------------------------------------------------------------------------
// BlockSparseMatrix::vmult on full local vectors (layout [u owned | p owned | u ghost | p ghost])
void system_vmult(Handle &H, const double *x, double *y)
{
  const int nu = H.nu_owned();
  if (H.nranks > 1) {
    halo_exchange_u(H, const_cast<double *>(x), H.ghost_off_u());
    halo_exchange_p(H, const_cast<double *>(x) + nu, H.ghost_off_p());
  }
  spmv_F(H, x, H.ghost_off_u(), x + nu, H.ghost_off_p(), y);
  spmv_B(H, x, H.ghost_off_u(), y + nu);
  H.n_vmult++;
}

static void apply_F(Handle &H, const double *x, int goff, double *y)
{
  if (H.nranks > 1) halo_exchange_u(H, const_cast<double *>(x), goff);
  spmv_F(H, x, goff, nullptr, 0, y);
}
static void apply_S(Handle &H, const double *x, int goff, double *y)
{
  if (H.nranks > 1) halo_exchange_p(H, const_cast<double *>(x), goff);
  spmv_S(H, x, goff, y);
}
static void apply_B(Handle &H, const double *x, int goff, double *y)
{
  if (H.nranks > 1) halo_exchange_u(H, const_cast<double *>(x), goff);
  spmv_B(H, x, goff, y);
}
static void apply_Bt(Handle &H, const double *x, int goff, double *y)
{
  if (H.nranks > 1) halo_exchange_p(H, const_cast<double *>(x), goff);
  spmv_Bt(H, x, goff, y);
}


// An inner SolverGMRES / SolverCG that hits maxiter (or NaN) throws SolverControl::NoConvergence out of
// Preconditioner::vmult in the reference and aborts the step; same here (-> NSB_ERR_NOCONV).
static void check_inner(int rc, const char *what, const Control &ctl)
{
  if (rc == 0) return;
  char buf[200];
  std::snprintf(buf, sizeof(buf), "SolverControl::NoConvergence: inner %s stopped at step %d with residual %.6e (tol %.6e)",
                what, ctl.last_step, ctl.last_value, ctl.tol);
  throw NoConvergence(buf);
}

static void inner_gmres_F(Handle &H, double *x, int goff_x, const double *b, double tol)
{
  auto &w = *H.ws;
  Space sp;
  sp.n_owned = H.nu_owned();
  sp.ld = H.dim * H.n_nodes;
  sp.goff_sol = goff_x;
  sp.A = [&H](const double *xx, int goff, double *y) { apply_F(H, xx, goff, y); };
  sp.P = [&H](const double *r, double *z) { ilu_solve(H, H.iluF, r, z); };
  Control ctl{H.prm.inner_maxit, tol};
  const int rc = gmres(H, sp, x, b, w.V_inner.p, H.prm.gmres_tmp, w.scal.p + 64, ctl, H.prm.orthogonalisation ? 1 : 0);
  H.n_inner_F += ctl.last_step;
  H.n_F_solves++;
  check_inner(rc, "GMRES on F", ctl);
}
static void inner_gmres_S(Handle &H, double *x, int goff_x, const double *b, double tol)
{
  auto &w = *H.ws;
  Space sp;
  sp.n_owned = H.n_p_owned;
  sp.ld = H.n_p;
  sp.goff_sol = goff_x;
  sp.A = [&H](const double *xx, int goff, double *y) { apply_S(H, xx, goff, y); };
  sp.P = [&H](const double *r, double *z) { ilu_solve(H, H.iluS, r, z); };
  Control ctl{H.prm.inner_maxit, tol};
  const int rc = gmres(H, sp, x, b, w.V_inner.p, H.prm.gmres_tmp, w.scal.p + 64, ctl, H.prm.orthogonalisation ? 1 : 0);
  H.n_inner_S += ctl.last_step;
  H.n_S_solves++;
  check_inner(rc, "GMRES on the Schur complement", ctl);
}
static void inner_cg_S(Handle &H, double *x, int goff_x, const double *b, double tol)
{
  auto &w = *H.ws;
  Space sp;
  sp.n_owned = H.n_p_owned;
  sp.ld = H.n_p;
  sp.goff_sol = goff_x;
  sp.A = [&H](const double *xx, int goff, double *y) { apply_S(H, xx, goff, y); };
  sp.P = [&H](const double *r, double *z) { ilu_solve(H, H.iluS, r, z); };
  Control ctl{H.prm.inner_maxit, tol};
  const int rc = cg(H, sp, x, b, w.tp[3].p, w.tp[4].p, w.tp[5].p, w.scal.p + 128, ctl);
  H.n_inner_S += ctl.last_step;
  H.n_S_solves++;
  check_inner(rc, "CG on the Schur complement", ctl);
}

// --------------------------------------------------------------------------------------------
// preconditioners.  src / dst: full local vectors.
// --------------------------------------------------------------------------------------------
void precond_init(Handle &H)
{
  extract_diag(H);
  if (H.nranks > 1) halo_exchange_u(H, H.d_negDinv.p, 0); // -1/D of ghost nodes for the Schur product
  spgemm_schur(H);
  ilu_factor(H, H.iluF, H.Fs.val.p);
  ilu_factor(H, H.iluS, H.S.val.p);
  H.prec_ready = true;
}

static void asimple_vmult(Handle &H, const double *src, double *dst)
{ // Preconditioners.hpp:254-311
  auto &w = *H.ws;
  const int nu = H.nu_owned(), np = H.n_p_owned;
  const double *su = src, *sp_ = src + nu;
  double *du = dst, *dp = dst + nu;
  double *tmp_u = w.tu[0].p, *tmp_p = w.tp[0].p;
  double *slot = w.scal.p + 200;
  inner_gmres_F(H, du, H.ghost_off_u(), su, H.prm.inner_rtol * norm_host(H, nu, su, slot)); // :271-273
  apply_B(H, du, H.ghost_off_u(), dp);                                                      // :280
  vec_sadd(H, np, -1.0, 1.0, sp_, dp);                                                      // :281
  vec_copy(H, np, dp, tmp_p);                                                               // :282
  inner_gmres_S(H, dp, H.ghost_off_p(), tmp_p, H.prm.inner_rtol * norm_host(H, np, tmp_p, slot)); // :287-289
  vec_pointwise(H, nu, H.d_D.p, du);                                                        // :294
  vec_scale(H, np, 1.0 / H.prm.alpha_asimple, dp);                                          // :298
  apply_Bt(H, dp, H.ghost_off_p(), tmp_u);                                                  // :304
  vec_axpy(H, nu, -1.0, tmp_u, du);                                                         // :305
  vec_pointwise(H, nu, H.d_Dinv.p, du);                                                     // :309
}

static void simple_vmult(Handle &H, const double *src, double *dst)
{ // Preconditioners.hpp:151-205
  auto &w = *H.ws;
  const int nu = H.nu_owned(), np = H.n_p_owned;
  const double *su = src, *sp_ = src + nu;
  double *du = dst, *dp = dst + nu;
  double *sol1_u = w.tu[0].p, *tmp = w.tu[1].p, *sol1_p = w.tp[0].p, *temp_1 = w.tp[1].p;
  double *slot = w.scal.p + 200;
  vec_copy(H, nu, su, sol1_u);
  vec_copy(H, np, sp_, sol1_p);
  inner_gmres_F(H, sol1_u, 0, su, H.prm.inner_rtol * norm_host(H, nu, su, slot));               // :173
  apply_B(H, sol1_u, 0, temp_1);                                                              // :175
  vec_axpy(H, np, -1.0, sp_, temp_1);                                                         // :176
  inner_cg_S(H, sol1_p, 0, temp_1, H.prm.inner_rtol * norm_host(H, np, temp_1, slot));         // :179-182
  vec_copy(H, np, sol1_p, dp);
  vec_scale(H, np, 1.0 / H.prm.alpha_simple, dp);                                             // :194-195
  vec_copy(H, nu, sol1_u, du);                                                                // :199
  apply_Bt(H, dp, H.ghost_off_p(), tmp);                                                      // :201
  vec_pointwise(H, nu, H.d_Dinv.p, tmp);                                                      // :202
  vec_axpy(H, nu, -1.0, tmp, du);                                                             // :203
}

static void yosida_vmult(Handle &H, const double *src, double *dst)
{ // Preconditioners.hpp:364-408
  auto &w = *H.ws;
  const int nu = H.nu_owned(), np = H.n_p_owned;
  const double *su = src, *sp_ = src + nu;
  double *du = dst, *dp = dst + nu;
  double *yu = w.tu[0].p, *tmp2 = w.tu[1].p, *res = w.tu[2].p, *yp = w.tp[0].p, *tmp = w.tp[1].p;
  double *slot = w.scal.p + 200;
  vec_copy(H, nu, su, yu);                                                                    // :375
  vec_copy(H, np, sp_, yp);                                                                   // :376
  inner_gmres_F(H, yu, 0, su, H.prm.inner_rtol * norm_host(H, nu, su, slot));                  // :371-382
  apply_B(H, yu, 0, tmp);                                                                     // :385
  vec_axpy(H, np, -1.0, sp_, tmp);                                                            // :386
  inner_cg_S(H, yp, 0, tmp, H.prm.inner_rtol * norm_host(H, np, tmp, slot));                   // :388-390
  vec_copy(H, np, yp, dp);                                                                    // :394
  apply_Bt(H, dp, H.ghost_off_p(), tmp2);                                                     // :398
  vec_zero(H, nu, res);                                                                       // :401
  inner_gmres_F(H, res, 0, tmp2, H.prm.inner_rtol * norm_host(H, nu, tmp2, slot));             // :403-405
  vec_copy(H, nu, yu, du);                                                                    // :402
  vec_sadd(H, nu, -1.0, 1.0, res, du);                                                        // :406 dst_u = -yu + res
}

static void ayosida_vmult(Handle &H, const double *src, double *dst)
{ // Preconditioners.hpp:474-517
  auto &w = *H.ws;
  const int nu = H.nu_owned(), np = H.n_p_owned;
  const double *su = src, *sp_ = src + nu;
  double *du = dst, *dp = dst + nu;
  double *tmp = w.tu[0].p, *yu = w.tu[1].p, *t3 = w.tu[2].p, *tmp2 = w.tp[0].p, *yp = w.tp[1].p;
  double *slot = w.scal.p + 200;
  vec_pointwise_out(H, nu, H.d_Dinv.p, su, tmp);                                              // :491-492
  vec_copy(H, nu, tmp, yu);                                                                   // :493
  vec_copy(H, np, sp_, yp);                                                                   // :487
  apply_B(H, tmp, 0, tmp2);                                                                   // :496
  vec_sadd(H, np, -1.0, 1.0, tmp2, yp);                                                       // :497
  inner_cg_S(H, dp, H.ghost_off_p(), yp, H.prm.inner_rtol * norm_host(H, np, yp, slot));       // :500-502
  vec_copy(H, np, dp, yp);                                                                    // :504
  apply_F(H, yu, 0, t3);                                                                      // :507 F->vmult(yu, yu)
  vec_copy(H, nu, t3, yu);
  apply_Bt(H, yp, 0, tmp);                                                                    // :510
  vec_sadd(H, nu, -1.0, 1.0, tmp, yu);                                                        // :511
  vec_pointwise(H, nu, H.d_Dinv.p, yu);                                                       // :514
  vec_copy(H, nu, yu, du);                                                                    // :515
}

void precond_vmult(Handle &H, const double *src, double *dst)
{
  switch (H.prm.precond_type) {
    case NSB_PREC_YOSIDA: yosida_vmult(H, src, dst); break;
    case NSB_PREC_SIMPLE: simple_vmult(H, src, dst); break;
    case NSB_PREC_AYOSIDA: ayosida_vmult(H, src, dst); break;
    case NSB_PREC_ASIMPLE: asimple_vmult(H, src, dst); break;
    default: throw ArgError("Invalid preconditioner type"); // NavierStokes2D.cpp:618
  }
}

// solver.solve(system_matrix, solution_owned, system_rhs, P)
int solve_outer(Handle &H)
{
  auto &w = *H.ws;
  Space sp;
  sp.n_owned = H.n_owned();
  sp.ld = H.n_local();
  sp.goff_sol = 0;
  sp.A = [&H](const double *x, int, double *y) { system_vmult(H, x, y); };
  sp.P = [&H](const double *r, double *z) { precond_vmult(H, r, z); };
  Control ctl{H.prm.outer_maxit, H.prm.outer_tol};
  NSB_CUDA(cudaMemsetAsync(w.V_outer.p, 0, sizeof(double) * size_t(sp.ld) * H.prm.gmres_tmp, H.stream));
  const int rc = gmres(H, sp, H.d_sol.p, H.d_rhs.p, w.V_outer.p, H.prm.gmres_tmp, w.scal.p, ctl,
                       H.prm.orthogonalisation ? 2 : 0);
  H.last_outer = ctl.last_step;
  H.last_res = ctl.last_value;
  return rc;
}

} // namespace nsb
