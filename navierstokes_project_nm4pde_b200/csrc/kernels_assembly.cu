// kernels_assembly.cu -- FP64 cell-loop assembly kernels (sm_100a).
//
// Replaces the cell loops of NavierStokes::assemble (Navier-Stokes/src/NavierStokes2D.cpp:209-313,
// NavierStokes3D.cpp / Convergence3D.cpp twins) and NavierStokes::assemble_time_step
// (NavierStokes2D.cpp:414-488), plus MatrixTools::apply_boundary_values (:354, :524).
//
// Mapping: one thread block = 32 consecutive cells x N2 local rows (N2 = 6 | 10 P2 nodes);
// a warp owns ONE local row i of 32 cells, so every per-cell load is a 32-wide coalesced
// access of the cell-interleaved arrays, reference-element tables are read as shared-memory
// broadcasts, and the Jacobian / advecting velocity live in registers.  Because mass,
// stiffness, convection and Temam terms never couple velocity components
// (NavierStokes2D.cpp:247-256) only the scalar N2 x N2 block is computed (F = I_dim (x) F_s).
// Local rows are scattered with FP64 atomics (RED.E.ADD.F64) through a precomputed position map.
#include <algorithm>
#include <cstring>

#include "nsb_internal.hpp"

namespace nsb {

__device__ __forceinline__ int dev_find(const int *__restrict__ rowptr, const int *__restrict__ colind, int row,
                                        int col)
{
  int lo = rowptr[row], hi = rowptr[row + 1] - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1, c = colind[mid];
    if (c == col) return mid;
    if (c < col) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

template <int DIM>
__device__ __forceinline__ double affine_inverse(const double (*sX)[32], int lane, double Jinv[DIM][DIM])
{
  // J[r][k] = x_{k+1}[r] - x_0[r];  sX row index = v*DIM + r
  double J[DIM][DIM];
#pragma unroll
  for (int r = 0; r < DIM; ++r)
#pragma unroll
    for (int k = 0; k < DIM; ++k) J[r][k] = sX[(k + 1) * DIM + r][lane] - sX[r][lane];
  double det;
  if constexpr (DIM == 2) {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double id = 1.0 / det;
    Jinv[0][0] = J[1][1] * id;  Jinv[0][1] = -J[0][1] * id;
    Jinv[1][0] = -J[1][0] * id; Jinv[1][1] = J[0][0] * id;
  } else {
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const double id = 1.0 / det;
    Jinv[0][0] = c00 * id;
    Jinv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    Jinv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    Jinv[1][0] = c01 * id;
    Jinv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    Jinv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    Jinv[2][0] = c02 * id;
    Jinv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    Jinv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
  }
  return det;
}

struct AsmArgs {
  const double *vcoords;
  const int *cell_nodes, *cell_p, *mapF;
  const FeTables *tab;
  const StepTensor *tensor;
  const double *sol; // local vector (solution incl. ghosts)
  int n_nodes_owned, n_p_owned, ghost_off_u;
  double inv_dt, visc;
  // outputs
  double *F, *M, *A, *rhs;
  // first step only
  const int *Fs_rowptr, *Fs_colind;
  const int *B_rowptr, *B_colind, *Bt_rowptr, *Bt_colind, *Mp_rowptr, *Mp_colind;
  double *Bv, *Btv, *Mpv;
  int conv_mult;
};

template <int DIM>
__device__ __forceinline__ void stage_cell(const AsmArgs &a, int64_t g, int lane, int i, int node_i,
                                           double (*sU)[DIM][32], double (*sX)[32])
{
  constexpr int N2 = (DIM == 2) ? 6 : 10, NV1 = DIM + 1;
  if (node_i >= 0) {
    const int64_t addr = int64_t(DIM) * node_i + (node_i >= a.n_nodes_owned ? a.ghost_off_u : 0);
#pragma unroll
    for (int d = 0; d < DIM; ++d) sU[i][d][lane] = a.sol[addr + d];
  } else {
#pragma unroll
    for (int d = 0; d < DIM; ++d) sU[i][d][lane] = 0.0;
  }
  for (int k = i; k < NV1 * DIM; k += N2) sX[k][lane] = a.vcoords[(g * (NV1 * DIM) + k) * 32 + lane];
}

// ---------------------------------------------------------------------------------------------
// assemble_time_step, quadrature-loop form (same loop nest as the reference, q outermost)
// ---------------------------------------------------------------------------------------------
template <int DIM, bool TEMAM>
__global__ void __launch_bounds__(32 * ((DIM == 2) ? 6 : 10)) assemble_step_q_kernel(AsmArgs a)
{
  constexpr int N2 = (DIM == 2) ? 6 : 10, NV1 = DIM + 1;
  __shared__ FeTables tab;
  __shared__ double sU[N2][DIM][32];
  __shared__ double sX[NV1 * DIM][32];
  const int lane = threadIdx.x, i = threadIdx.y;
  const int64_t g = blockIdx.x;
  {
    const int tid = i * 32 + lane, nthr = 32 * N2;
    const int *src = reinterpret_cast<const int *>(a.tab);
    int *dst = reinterpret_cast<int *>(&tab);
    for (int k = tid; k < int(sizeof(FeTables) / sizeof(int)); k += nthr) dst[k] = src[k];
  }
  const int node_i = a.cell_nodes[(g * N2 + i) * 32 + lane];
  stage_cell<DIM>(a, g, lane, i, node_i, sU, sX);
  __syncthreads();
  if (node_i < 0) return;
  double Jinv[DIM][DIM];
  const double det = affine_inverse<DIM>(sX, lane, Jinv);
  // reference-space nodal velocities Ut_a = J^{-1} U_a are only needed through sums; keep U in smem
  double acc[N2];
#pragma unroll
  for (int j = 0; j < N2; ++j) acc[j] = 0.0;
  double rhs[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) rhs[d] = 0.0;
  const int nq = tab.nq;
  for (int q = 0; q < nq; ++q) {
    double u[DIM], G[DIM][DIM]; // G = sum_a U_a (x) dphi_hat_a  (velocity gradient in reference coords)
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      u[d] = 0.0;
#pragma unroll
      for (int k = 0; k < DIM; ++k) G[d][k] = 0.0;
    }
#pragma unroll
    for (int n = 0; n < N2; ++n) {
      const double ph = tab.phi[n][q];
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        const double U = sU[n][d][lane];
        u[d] += U * ph;
        if (TEMAM) {
#pragma unroll
          for (int k = 0; k < DIM; ++k) G[d][k] += U * tab.dphi[n][q][k];
        }
      }
    }
    double ut[DIM]; // J^{-1} u: (grad phi_j . u) = dphi_hat_j . ut
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      ut[k] = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) ut[k] += Jinv[k][d] * u[d];
    }
    double hdiv = 0.0; // 0.5 * div u = 0.5 * sum_d sum_k Jinv[k][d] G[d][k]
    if (TEMAM) {
#pragma unroll
      for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int k = 0; k < DIM; ++k) hdiv += Jinv[k][d] * G[d][k];
      hdiv *= 0.5;
    }
    const double fi = tab.phi[i][q] * (det * tab.w[q]);
#pragma unroll
    for (int j = 0; j < N2; ++j) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < DIM; ++k) t += tab.dphi[j][q][k] * ut[k];
      if (TEMAM) t += hdiv * tab.phi[j][q];
      acc[j] += t * fi;
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) rhs[d] += u[d] * fi;
  }
  if (node_i < a.n_nodes_owned) {
    const int *map = a.mapF + (g * (N2 * N2) + i * N2) * 32 + lane;
#pragma unroll
    for (int j = 0; j < N2; ++j) {
      const int pos = map[j * 32];
      if (pos >= 0) atomicAdd(a.F + pos, acc[j]);
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) atomicAdd(a.rhs + int64_t(DIM) * node_i + d, rhs[d] * a.inv_dt);
  }
}

// ---------------------------------------------------------------------------------------------
// assemble_time_step, tensor-contracted form: the q loop is folded into a constant reference
// tensor (exactly the same finite sum, re-associated), leaving N2*DIM*N2 FMAs per local row.
// ---------------------------------------------------------------------------------------------
// Persistent: a block loads the reference tensor into shared memory ONCE and then walks over groups
// of 32 cells (ncu of the one-group-per-block version: 45% of the L1 wavefront budget went into
// 64-bit shared loads of T and every block re-packed the 24 KB tensor).  T rows are read as 128-bit
// broadcasts (two j per load).
template <int DIM>
__global__ void __launch_bounds__(32 * ((DIM == 2) ? 6 : 10), 3) assemble_step_t_kernel(AsmArgs a, int n_groups)
{
  constexpr int N2 = (DIM == 2) ? 6 : 10;
  extern __shared__ __align__(16) double smem[];
  // layout: T[i][a][k][j] (N2*N2*DIM*N2) | Mh[N2][N2] | sU[N2][DIM][32] | sUt[N2][DIM][32] | sX[(DIM+1)*DIM][32]
  double *sT = smem;
  double *sMh = sT + N2 * N2 * DIM * N2;
  double(*sU)[DIM][32] = reinterpret_cast<double(*)[DIM][32]>(sMh + N2 * N2);
  double(*sUt)[DIM][32] = reinterpret_cast<double(*)[DIM][32]>(sMh + N2 * N2 + N2 * DIM * 32);
  double(*sX)[32] = reinterpret_cast<double(*)[32]>(sMh + N2 * N2 + 2 * N2 * DIM * 32);
  const int lane = threadIdx.x, i = threadIdx.y;
  {
    const int tid = i * 32 + lane, nthr = 32 * N2;
    // StepTensor::T is [10][10][10][3] = [i][j][a][k]; re-pack to [i][a][k][j]
    for (int idx = tid; idx < N2 * N2 * DIM * N2; idx += nthr) {
      const int j = idx % N2, k = (idx / N2) % DIM, aa = (idx / (N2 * DIM)) % N2, ii = idx / (N2 * DIM * N2);
      sT[idx] = a.tensor->T[ii][j][aa][k];
    }
    for (int idx = tid; idx < N2 * N2; idx += nthr) sMh[idx] = a.tensor->Mh[idx / N2][idx % N2];
  }
  const double2 *Ti = reinterpret_cast<const double2 *>(sT + i * (N2 * DIM * N2));
  for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int node_i = a.cell_nodes[(g * N2 + i) * 32 + lane];
    __syncthreads(); // the previous group's readers of sU / sUt / sX are done (and T is in place)
    stage_cell<DIM>(a, g, lane, i, node_i, sU, sX);
    __syncthreads();
    double Jinv[DIM][DIM];
    double det = 1.0;
    if (node_i >= 0) {
      det = affine_inverse<DIM>(sX, lane, Jinv);
#pragma unroll
      for (int k = 0; k < DIM; ++k) {
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) s += Jinv[k][d] * sU[i][d][lane];
        sUt[i][k][lane] = s * det;
      }
    }
    __syncthreads();
    if (node_i < 0 || node_i >= a.n_nodes_owned) continue;
    // the scatter positions are independent of the arithmetic: request them first
    int pos[N2];
    const int *map = a.mapF + (g * (N2 * N2) + i * N2) * 32 + lane;
#pragma unroll
    for (int j = 0; j < N2; ++j) pos[j] = __ldcs(map + j * 32);
    double acc[N2];
#pragma unroll
    for (int j = 0; j < N2; ++j) acc[j] = 0.0;
#pragma unroll 2
    for (int n = 0; n < N2; ++n) {
#pragma unroll
      for (int k = 0; k < DIM; ++k) {
        const double ut = sUt[n][k][lane];
        const double2 *Tr = Ti + (n * DIM + k) * (N2 / 2);
#pragma unroll
        for (int jj = 0; jj < N2 / 2; ++jj) {
          const double2 t = Tr[jj];
          acc[2 * jj] += t.x * ut;
          acc[2 * jj + 1] += t.y * ut;
        }
      }
    }
    double rhs[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) rhs[d] = 0.0;
#pragma unroll
    for (int n = 0; n < N2; ++n) {
      const double m = sMh[i * N2 + n];
#pragma unroll
      for (int d = 0; d < DIM; ++d) rhs[d] += m * sU[n][d][lane];
    }
#pragma unroll
    for (int j = 0; j < N2; ++j)
      if (pos[j] >= 0) atomicAdd(a.F + pos[j], acc[j]);
    const double sc = det * a.inv_dt;
#pragma unroll
    for (int d = 0; d < DIM; ++d) atomicAdd(a.rhs + int64_t(DIM) * node_i + d, rhs[d] * sc);
  }
}

// ---------------------------------------------------------------------------------------------
// assemble (first step): mass/dt, stiffness, convection (+Temam, x conv_mult), B, B^T, pressure
// mass and rhs.  Runs once, so positions are found by binary search instead of a map.
// ---------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(32 * ((DIM == 2) ? 6 : 10)) assemble_first_kernel(AsmArgs a)
{
  constexpr int N2 = (DIM == 2) ? 6 : 10, NV1 = DIM + 1;
  __shared__ FeTables tab;
  __shared__ double sU[N2][DIM][32];
  __shared__ double sX[NV1 * DIM][32];
  const int lane = threadIdx.x, i = threadIdx.y;
  const int64_t g = blockIdx.x;
  {
    const int tid = i * 32 + lane, nthr = 32 * N2;
    const int *src = reinterpret_cast<const int *>(a.tab);
    int *dst = reinterpret_cast<int *>(&tab);
    for (int k = tid; k < int(sizeof(FeTables) / sizeof(int)); k += nthr) dst[k] = src[k];
  }
  const int node_i = a.cell_nodes[(g * N2 + i) * 32 + lane];
  stage_cell<DIM>(a, g, lane, i, node_i, sU, sX);
  __syncthreads();
  if (node_i < 0) return;
  double Jinv[DIM][DIM];
  const double det = affine_inverse<DIM>(sX, lane, Jinv);
  double accM[N2], accA[N2], accC[N2], accB[NV1][DIM], accP[NV1], rhs[DIM];
#pragma unroll
  for (int j = 0; j < N2; ++j) accM[j] = accA[j] = accC[j] = 0.0;
#pragma unroll
  for (int v = 0; v < NV1; ++v) {
    accP[v] = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) accB[v][d] = 0.0;
  }
#pragma unroll
  for (int d = 0; d < DIM; ++d) rhs[d] = 0.0;
  const int nq = tab.nq;
  for (int q = 0; q < nq; ++q) {
    double u[DIM], divu = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) u[d] = 0.0;
    for (int n = 0; n < N2; ++n) {
      const double ph = tab.phi[n][q];
      double gr[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        gr[d] = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) gr[d] += Jinv[k][d] * tab.dphi[n][q][k];
      }
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        const double U = sU[n][d][lane];
        u[d] += U * ph;
        divu += U * gr[d];
      }
    }
    const double JxW = det * tab.w[q];
    double gi[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      gi[d] = 0.0;
#pragma unroll
      for (int k = 0; k < DIM; ++k) gi[d] += Jinv[k][d] * tab.dphi[i][q][k];
    }
    const double vi = tab.phi[i][q];
    for (int j = 0; j < N2; ++j) {
      double gj[DIM], gg = 0.0, adv = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        gj[d] = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) gj[d] += Jinv[k][d] * tab.dphi[j][q][k];
        gg += gi[d] * gj[d];
        adv += gj[d] * u[d];
      }
      const double vv = vi * tab.phi[j][q];
      accA[j] += a.visc * gg * JxW;
      accM[j] += vv * a.inv_dt * JxW;
      accC[j] += double(a.conv_mult) * (adv * vi * JxW) + 0.5 * divu * vv * JxW;
    }
#pragma unroll
    for (int v = 0; v < NV1; ++v) {
      const double ps = tab.psi[v][q];
#pragma unroll
      for (int d = 0; d < DIM; ++d) accB[v][d] += ps * gi[d] * JxW; // psi_v * d_c phi_i
      if (i < NV1) accP[v] += tab.psi[i][q] * ps / a.visc * JxW;
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) rhs[d] += u[d] * vi * JxW * a.inv_dt;
  }
  const int *cn = a.cell_nodes + g * N2 * 32 + lane;
  const int *cp = a.cell_p + g * NV1 * 32 + lane;
  if (node_i < a.n_nodes_owned) {
    for (int j = 0; j < N2; ++j) {
      const int pos = dev_find(a.Fs_rowptr, a.Fs_colind, node_i, cn[j * 32]);
      atomicAdd(a.M + pos, accM[j]);
      atomicAdd(a.A + pos, accA[j]);
      atomicAdd(a.F + pos, accC[j]);
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) atomicAdd(a.rhs + int64_t(DIM) * node_i + d, rhs[d]);
  }
  // block (0,1): -psi_v div(phi_i)  (NavierStokes2D.cpp:259).  Rows of ghost nodes are assembled
  // too: the Schur product needs them and their cell stars are local (2-layer cell halo).
  for (int v = 0; v < NV1; ++v) {
    const int pos = dev_find(a.Bt_rowptr, a.Bt_colind, node_i, cp[v * 32]);
#pragma unroll
    for (int d = 0; d < DIM; ++d) atomicAdd(a.Btv + int64_t(pos) * DIM + d, -accB[v][d]);
  }
  for (int v = 0; v < NV1; ++v) { // block (1,0): +psi_v div(phi_j)  (NavierStokes2D.cpp:262)
    const int prow = cp[v * 32];
    if (prow >= a.n_p_owned) continue;
    const int pos = dev_find(a.B_rowptr, a.B_colind, prow, node_i);
#pragma unroll
    for (int d = 0; d < DIM; ++d) atomicAdd(a.Bv + int64_t(pos) * DIM + d, accB[v][d]);
  }
  if (i < NV1) {
    const int prow = cp[i * 32];
    if (prow < a.n_p_owned)
      for (int v = 0; v < NV1; ++v) {
        const int pos = dev_find(a.Mp_rowptr, a.Mp_colind, prow, cp[v * 32]);
        atomicAdd(a.Mpv + pos, accP[v]);
      }
  }
}

__global__ void add3_kernel(int64_t n, const double *__restrict__ M, const double *__restrict__ A,
                            double *__restrict__ K, double *__restrict__ F)
{ // K = M + A ; F = F(=C) + K
  for (int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; k < n; k += int64_t(gridDim.x) * blockDim.x) {
    const double kk = M[k] + A[k];
    K[k] = kk;
    F[k] = kk + F[k];
  }
}

__global__ void add_vec_kernel(int n, const double *__restrict__ x, double *__restrict__ y)
{
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) y[k] += x[k];
}

// MatrixTools::apply_boundary_values (Trilinos block path), one warp per constrained P2 node.
template <int DIM>
__global__ void dirichlet_kernel(int n_dir, int n_nodes_owned, const int *__restrict__ nodes, const double *__restrict__ gvals,
                                 const int *__restrict__ rowptr, const int *__restrict__ diagpos,
                                 double *__restrict__ F, const int *__restrict__ bt_rowptr, double *__restrict__ Btv,
                                 double *__restrict__ rhs, const double *__restrict__ dbar_dev, int mode,
                                 int clear_bt)
{
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_dir) return;
  const int node = nodes[w];
  if (clear_bt)
    for (int k = bt_rowptr[node] * DIM + lane; k < bt_rowptr[node + 1] * DIM; k += 32) Btv[k] = 0.0;
  if (node >= n_nodes_owned) return; // ghost node: only its (redundantly assembled) Bt row is cleared
  const int dp = diagpos[node];
  for (int k = rowptr[node] + lane; k < rowptr[node + 1]; k += 32)
    if (k != dp) F[k] = 0.0;
  if (lane == 0) {
    double diag = F[dp];
    if (mode == 1 || diag == 0.0) { diag = *dbar_dev; F[dp] = diag; }
#pragma unroll
    for (int d = 0; d < DIM; ++d) rhs[int64_t(DIM) * node + d] = gvals[int64_t(w) * DIM + d] * diag;
  }
}

__global__ void find_dbar_kernel(int n, const int *__restrict__ diagpos, const double *__restrict__ F,
                                 double *__restrict__ out)
{ // "first nonzero diagonal entry" of the locally owned range of block (0,0)
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double d = 1.0;
    for (int i = 0; i < n; ++i) {
      const double v = F[diagpos[i]];
      if (v != 0.0) { d = fabs(v); break; }
    }
    *out = d;
  }
}

static AsmArgs make_args(Handle &H)
{
  AsmArgs a{};
  a.vcoords = H.d_vcoords.p; a.cell_nodes = H.d_cell_nodes.p; a.cell_p = H.d_cell_p.p; a.mapF = H.d_mapF.p;
  a.tab = H.d_tab.p; a.tensor = H.d_step_tensor.p;
  a.sol = H.d_sol.p;
  a.n_nodes_owned = H.n_nodes_owned; a.n_p_owned = H.n_p_owned; a.ghost_off_u = H.ghost_off_u();
  a.inv_dt = 1.0 / H.prm.deltat; a.visc = H.prm.nu;
  a.F = H.Fs.val.p; a.M = H.d_M.p; a.A = H.d_A.p; a.rhs = H.d_rhs.p;
  a.Fs_rowptr = H.Fs.rowptr.p; a.Fs_colind = H.Fs.colind.p;
  a.B_rowptr = H.B.rowptr.p; a.B_colind = H.B.colind.p;
  a.Bt_rowptr = H.Bt.rowptr.p; a.Bt_colind = H.Bt.colind.p;
  a.Mp_rowptr = H.Mp.rowptr.p; a.Mp_colind = H.Mp.colind.p;
  a.Bv = H.B.val.p; a.Btv = H.Bt.val.p; a.Mpv = H.Mp.val.p;
  a.conv_mult = 1;
  return a;
}

void launch_assemble_first(Handle &H)
{
  AsmArgs a = make_args(H);
  a.conv_mult = (H.prm.variant == NSB_VARIANT_CONV) ? 2 : 1; // Convergence3D.cpp:277 + :284
  cudaStream_t s = H.stream;
  H.Fs.val.zero(s); H.d_M.zero(s); H.d_A.zero(s); H.B.val.zero(s); H.Bt.val.zero(s); H.Mp.val.zero(s);
  H.d_rhs.zero(s);
  const unsigned groups = unsigned(H.nc_pad / 32);
  if (H.dim == 2) assemble_first_kernel<2><<<groups, dim3(32, 6), 0, s>>>(a);
  else assemble_first_kernel<3><<<groups, dim3(32, 10), 0, s>>>(a);
  NSB_CUDA(cudaGetLastError());
  const int64_t nnz = H.Fs.nnz;
  add3_kernel<<<std::min<int64_t>((nnz + 255) / 256, 148 * 16), 256, 0, s>>>(nnz, H.d_M.p, H.d_A.p, H.d_K.p,
                                                                              H.Fs.val.p);
  NSB_CUDA(cudaGetLastError());
  H.launches += 2;
  if (H.have_neumann) {
    add_vec_kernel<<<148 * 4, 256, 0, s>>>(H.nu_owned(), H.d_neumann.p, H.d_rhs.p);
    H.launches++;
  }
}

// F_target = K + C(u_n) accumulated by atomics; rhs = M u_n / dt (+ Neumann)
void launch_assemble_step(Handle &H, double *F_target)
{
  AsmArgs a = make_args(H);
  a.F = F_target;
  cudaStream_t s = H.stream;
  if (F_target == H.Fs.val.p)
    NSB_CUDA(cudaMemcpyAsync(H.Fs.val.p, H.d_K.p, sizeof(double) * H.Fs.nnz, cudaMemcpyDeviceToDevice, s));
  NSB_CUDA(cudaMemsetAsync(H.d_rhs.p, 0, sizeof(double) * H.n_local(), s));
  const unsigned groups = unsigned(H.nc_pad / 32);
  const bool temam = H.prm.variant != NSB_VARIANT_3D; // NavierStokes3D.cpp:456 has no Temam term
  if (H.prm.assembly_kernel == 1) {
    if (H.dim == 2) {
      if (temam) assemble_step_q_kernel<2, true><<<groups, dim3(32, 6), 0, s>>>(a);
      else assemble_step_q_kernel<2, false><<<groups, dim3(32, 6), 0, s>>>(a);
    } else {
      if (temam) assemble_step_q_kernel<3, true><<<groups, dim3(32, 10), 0, s>>>(a);
      else assemble_step_q_kernel<3, false><<<groups, dim3(32, 10), 0, s>>>(a);
    }
  } else {
    if (H.dim == 2) {
      const size_t sm = sizeof(double) * (6 * 6 * 2 * 6 + 36 + 2 * 6 * 2 * 32 + 3 * 2 * 32);
      assemble_step_t_kernel<2><<<std::min(groups, 148u * 5u), dim3(32, 6), sm, s>>>(a, int(groups));
    } else {
      const size_t sm = sizeof(double) * (10 * 10 * 3 * 10 + 100 + 2 * 10 * 3 * 32 + 4 * 3 * 32);
      static bool attr_set = false;
      if (!attr_set) {
        NSB_CUDA(cudaFuncSetAttribute(assemble_step_t_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));
        attr_set = true;
      }
      assemble_step_t_kernel<3><<<std::min(groups, 148u * 3u), dim3(32, 10), sm, s>>>(a, int(groups));
    }
  }
  NSB_CUDA(cudaGetLastError());
  H.launches += 1;
  if (H.have_neumann) {
    add_vec_kernel<<<148 * 4, 256, 0, s>>>(H.nu_owned(), H.d_neumann.p, H.d_rhs.p);
    H.launches++;
  }
}

void launch_apply_dirichlet(Handle &H, bool clear_bt)
{
  const int n_dir = int(H.h_dir_nodes.size());
  if (n_dir == 0) return;
  cudaStream_t s = H.stream;
  double *dbar = H.d_scratch.p;
  find_dbar_kernel<<<1, 32, 0, s>>>(H.n_nodes_owned, H.d_diagF.p, H.Fs.val.p, dbar);
  const int threads = 128, blocks = (n_dir * 32 + threads - 1) / threads;
  if (H.dim == 2)
    dirichlet_kernel<2><<<blocks, threads, 0, s>>>(n_dir, H.n_nodes_owned, H.d_dir_nodes.p, H.d_dir_vals.p, H.Fs.rowptr.p,
                                                   H.d_diagF.p, H.Fs.val.p, H.Bt.rowptr.p, H.Bt.val.p, H.d_rhs.p,
                                                   dbar, H.prm.dirichlet_mode, clear_bt ? 1 : 0);
  else
    dirichlet_kernel<3><<<blocks, threads, 0, s>>>(n_dir, H.n_nodes_owned, H.d_dir_nodes.p, H.d_dir_vals.p, H.Fs.rowptr.p,
                                                   H.d_diagF.p, H.Fs.val.p, H.Bt.rowptr.p, H.Bt.val.p, H.d_rhs.p,
                                                   dbar, H.prm.dirichlet_mode, clear_bt ? 1 : 0);
  NSB_CUDA(cudaGetLastError());
  H.launches += 2;
}

// Host-side contraction of the quadrature sums into the reference tensor of assemble_step_t_kernel.
void build_step_tensor(const FeTables &tab, int dim, bool temam, StepTensor &out)
{
  const int n2 = (dim == 2) ? 6 : 10;
  std::memset(&out, 0, sizeof(out));
  for (int i = 0; i < n2; ++i)
    for (int j = 0; j < n2; ++j)
      for (int a = 0; a < n2; ++a)
        for (int k = 0; k < dim; ++k) {
          double s = 0.0;
          for (int q = 0; q < tab.nq; ++q) {
            double t = tab.dphi[j][q][k] * tab.phi[a][q];
            if (temam) t += 0.5 * tab.phi[j][q] * tab.dphi[a][q][k];
            s += tab.w[q] * tab.phi[i][q] * t;
          }
          out.T[i][j][a][k] = s;
        }
  for (int i = 0; i < n2; ++i)
    for (int a = 0; a < n2; ++a) {
      double s = 0.0;
      for (int q = 0; q < tab.nq; ++q) s += tab.w[q] * tab.phi[i][q] * tab.phi[a][q];
      out.Mh[i][a] = s;
    }
}

} // namespace nsb
