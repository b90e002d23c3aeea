// kernels_linalg.cu -- bandwidth-bound sparse / dense vector kernels (sm_100a, FP64).
//
// Replaces what the reference gets from Trilinos through deal.II's wrappers:
//   Epetra_CrsMatrix::Multiply (SparseMatrix::vmult; Preconditioners.hpp:280,304; the block
//   vmult inside SolverGMRES, NavierStokes2D.cpp:609), Epetra vector ops, Ifpack_ILU
//   (PreconditionILU, Preconditioners.hpp:250-251) and EpetraExt MatrixMatrix::Multiply
//   (SparseMatrix::mmult, Preconditioners.hpp:248,358).
// Layout: the velocity block is stored once as the scalar node graph F_s and applied to the
// `dim` interleaved components of each P2 node; B / Bt keep `dim` values per (vertex,node) pair.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "nsb_internal.hpp"

namespace nsb {

constexpr int kSM = 148;

// --------------------------------------------------------------------------------------------
// SpMV
// --------------------------------------------------------------------------------------------
template <int DIM, int LPR>
__global__ void __launch_bounds__(256) spmv_F_kernel(int n_rows, const int *__restrict__ rowptr,
                                                     const int *__restrict__ colind, const double *__restrict__ val,
                                                     const double *__restrict__ xu, int n_nodes_owned, int goff_u,
                                                     const int *__restrict__ bt_rowptr,
                                                     const int *__restrict__ bt_colind,
                                                     const double *__restrict__ bt_val, const double *__restrict__ xp,
                                                     int n_p_owned, int goff_p, double *__restrict__ yu)
{
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = tid / LPR, sl = tid % LPR;
  if (row >= n_rows) return;
  double acc[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) acc[d] = 0.0;
  const int e = rowptr[row + 1];
  for (int k = rowptr[row] + sl; k < e; k += LPR) {
    const int c = colind[k];
    const double v = val[k];
    const double *xb = xu + int64_t(DIM) * c + (c >= n_nodes_owned ? goff_u : 0);
#pragma unroll
    for (int d = 0; d < DIM; ++d) acc[d] += v * xb[d];
  }
  if (xp) {
    const int e2 = bt_rowptr[row + 1];
    for (int k = bt_rowptr[row] + sl; k < e2; k += LPR) {
      const int c = bt_colind[k];
      const double x = xp[c + (c >= n_p_owned ? goff_p : 0)];
#pragma unroll
      for (int d = 0; d < DIM; ++d) acc[d] += bt_val[int64_t(k) * DIM + d] * x;
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
    for (int d = 0; d < DIM; ++d) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o, LPR);
  if (sl == 0) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) yu[int64_t(DIM) * row + d] = acc[d];
  }
}

template <int DIM, int LPR>
__global__ void __launch_bounds__(256) spmv_B_kernel(int n_rows, const int *__restrict__ rowptr,
                                                     const int *__restrict__ colind, const double *__restrict__ val,
                                                     const double *__restrict__ xu, int n_nodes_owned, int goff_u,
                                                     double *__restrict__ yp)
{
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = tid / LPR, sl = tid % LPR;
  if (row >= n_rows) return;
  double acc = 0.0;
  const int e = rowptr[row + 1];
  for (int k = rowptr[row] + sl; k < e; k += LPR) {
    const int c = colind[k];
    const double *xb = xu + int64_t(DIM) * c + (c >= n_nodes_owned ? goff_u : 0);
#pragma unroll
    for (int d = 0; d < DIM; ++d) acc += val[int64_t(k) * DIM + d] * xb[d];
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, LPR);
  if (sl == 0) yp[row] = acc;
}

template <int LPR>
__global__ void __launch_bounds__(256) spmv_csr_kernel(int n_rows, const int *__restrict__ rowptr,
                                                       const int *__restrict__ colind, const double *__restrict__ val,
                                                       const double *__restrict__ x, int n_owned, int goff,
                                                       double *__restrict__ y)
{
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = tid / LPR, sl = tid % LPR;
  if (row >= n_rows) return;
  double acc = 0.0;
  const int e = rowptr[row + 1];
  for (int k = rowptr[row] + sl; k < e; k += LPR) {
    const int c = colind[k];
    acc += val[k] * x[c + (c >= n_owned ? goff : 0)];
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, LPR);
  if (sl == 0) y[row] = acc;
}

// 3D variant: one warp per node row, lane = d*8 + e (component d of entry e, lanes 24..31 idle).
// The three lanes that share an entry read the three consecutive doubles of x, so the gather of
// an entry costs ~1.5 L1 sectors instead of three separate 8-byte accesses (the L1/TEX pipe, not
// DRAM, limited the sub-warp-per-row kernel: ncu profiles/r01).
__global__ void __launch_bounds__(256) spmv_F3_kernel(int n_rows, const int *__restrict__ rowptr,
                                                      const int *__restrict__ colind, const double *__restrict__ val,
                                                      const double *__restrict__ xu, int n_nodes_owned, int goff_u,
                                                      const int *__restrict__ bt_rowptr,
                                                      const int *__restrict__ bt_colind,
                                                      const double *__restrict__ bt_val, const double *__restrict__ xp,
                                                      int n_p_owned, int goff_p, double *__restrict__ yu)
{
  const int lane = threadIdx.x & 31;
  const int d = lane >> 3, e = lane & 7;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += warps_per_grid) {
    double acc = 0.0;
    if (d < 3) {
      const int rs = rowptr[row], re = rowptr[row + 1];
      int k = rs + e;
      for (; k + 8 < re; k += 16) { // two independent gathers in flight
        const int c0 = __ldcs(colind + k), c1 = __ldcs(colind + k + 8);
        const double v0 = __ldcs(val + k), v1 = __ldcs(val + k + 8);
        const double x0 = xu[int64_t(3) * c0 + (c0 >= n_nodes_owned ? goff_u : 0) + d];
        const double x1 = xu[int64_t(3) * c1 + (c1 >= n_nodes_owned ? goff_u : 0) + d];
        acc += v0 * x0;
        acc += v1 * x1;
      }
      if (k < re) {
        const int c0 = __ldcs(colind + k);
        acc += __ldcs(val + k) * xu[int64_t(3) * c0 + (c0 >= n_nodes_owned ? goff_u : 0) + d];
      }
      if (xp) {
        const int e2 = bt_rowptr[row + 1];
        for (int q = bt_rowptr[row] + e; q < e2; q += 8) {
          const int c = bt_colind[q];
          acc += bt_val[int64_t(q) * 3 + d] * xp[c + (c >= n_p_owned ? goff_p : 0)];
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (e == 0 && d < 3) yu[int64_t(3) * row + d] = acc;
  }
}

static inline unsigned grid_for(int64_t threads, int block) { return unsigned((threads + block - 1) / block); }

void spmv_Bt_acc(Handle &H, const double *x_p, int goff_p, double *y_u, bool accumulate);

void spmv_F(Handle &H, const double *x_u, int goff_u, const double *x_p, int goff_p, double *y_u)
{
  if (H.n_nodes_owned == 0) return;
  stream_spmv_F(H, x_u, goff_u, y_u);
  if (x_p) spmv_Bt_acc(H, x_p, goff_p, y_u, true);
  H.cnt_spmv_F++;
}

template <int DIM, int LPR>
__global__ void __launch_bounds__(256) spmv_Bt_kernel(int n_rows, const int *__restrict__ bt_rowptr,
                                                      const int *__restrict__ bt_colind,
                                                      const double *__restrict__ bt_val, const double *__restrict__ xp,
                                                      int n_p_owned, int goff_p, double *__restrict__ yu, int accumulate)
{
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = tid / LPR, sl = tid % LPR;
  if (row >= n_rows) return;
  double acc[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) acc[d] = 0.0;
  const int e2 = bt_rowptr[row + 1];
  for (int k = bt_rowptr[row] + sl; k < e2; k += LPR) {
    const int c = bt_colind[k];
    const double x = xp[c + (c >= n_p_owned ? goff_p : 0)];
#pragma unroll
    for (int d = 0; d < DIM; ++d) acc[d] += bt_val[int64_t(k) * DIM + d] * x;
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
    for (int d = 0; d < DIM; ++d) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o, LPR);
  if (sl == 0) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      if (accumulate) yu[int64_t(DIM) * row + d] += acc[d];
      else yu[int64_t(DIM) * row + d] = acc[d];
    }
  }
}

void spmv_Bt_acc(Handle &H, const double *x_p, int goff_p, double *y_u, bool accumulate)
{
  const int n = H.n_nodes_owned;
  if (n == 0) return;
  constexpr int LPR = 4;
  const unsigned grid = grid_for(int64_t(n) * LPR, 256);
  if (H.dim == 2)
    spmv_Bt_kernel<2, LPR><<<grid, 256, 0, H.stream>>>(n, H.Bt.rowptr.p, H.Bt.colind.p, H.Bt.val.p, x_p, H.n_p_owned,
                                                       goff_p, y_u, accumulate ? 1 : 0);
  else
    spmv_Bt_kernel<3, LPR><<<grid, 256, 0, H.stream>>>(n, H.Bt.rowptr.p, H.Bt.colind.p, H.Bt.val.p, x_p, H.n_p_owned,
                                                       goff_p, y_u, accumulate ? 1 : 0);
  NSB_CUDA(cudaGetLastError());
  H.launches++;
  H.cnt_spmv_Bt++;
}

void spmv_Bt(Handle &H, const double *x_p, int goff_p, double *y_u) { spmv_Bt_acc(H, x_p, goff_p, y_u, false); }

void spmv_B(Handle &H, const double *x_u, int goff_u, double *y_p)
{
  const int n = H.n_p_owned;
  if (n == 0) return;
  constexpr int LPR = 16;
  const unsigned grid = grid_for(int64_t(n) * LPR, 256);
  if (H.dim == 2)
    spmv_B_kernel<2, LPR><<<grid, 256, 0, H.stream>>>(n, H.B.rowptr.p, H.B.colind.p, H.B.val.p, x_u, H.n_nodes_owned,
                                                      goff_u, y_p);
  else
    spmv_B_kernel<3, LPR><<<grid, 256, 0, H.stream>>>(n, H.B.rowptr.p, H.B.colind.p, H.B.val.p, x_u, H.n_nodes_owned,
                                                      goff_u, y_p);
  NSB_CUDA(cudaGetLastError());
  H.launches++;
  H.cnt_spmv_B++;
}

void spmv_S(Handle &H, const double *x_p, int goff_p, double *y_p)
{
  if (H.n_p_owned == 0) return;
  stream_spmv_S(H, x_p, goff_p, y_p);
  H.cnt_spmv_S++;
}

// --------------------------------------------------------------------------------------------
// dense vector kernels
// --------------------------------------------------------------------------------------------
#define GRID_STRIDE(i, n) for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += gridDim.x * blockDim.x)

static inline unsigned vgrid(int n) { return unsigned(std::max(1, std::min((n + 255) / 256, kSM * 8))); }

__global__ void k_copy(int n, const double *__restrict__ x, double *__restrict__ y) { GRID_STRIDE(i, n) y[i] = x[i]; }
__global__ void k_zero(int n, double *__restrict__ x) { GRID_STRIDE(i, n) x[i] = 0.0; }
__global__ void k_axpy(int n, double a, const double *__restrict__ x, double *__restrict__ y)
{ GRID_STRIDE(i, n) y[i] += a * x[i]; }
__global__ void k_axpy_dev(int n, const double *__restrict__ a, double sign, const double *__restrict__ x,
                           double *__restrict__ y)
{
  const double aa = sign * (*a);
  GRID_STRIDE(i, n) y[i] += aa * x[i];
}
__global__ void k_sadd(int n, double s, double a, const double *__restrict__ x, double *__restrict__ y)
{ GRID_STRIDE(i, n) y[i] = s * y[i] + a * x[i]; }
__global__ void k_scale(int n, double a, double *__restrict__ x) { GRID_STRIDE(i, n) x[i] *= a; }
__global__ void k_scale_inv_dev(int n, const double *__restrict__ s, double *__restrict__ x)
{
  const double a = 1.0 / (*s);
  GRID_STRIDE(i, n) x[i] *= a;
}
__global__ void k_pointwise(int n, const double *__restrict__ d, double *__restrict__ x) { GRID_STRIDE(i, n) x[i] *= d[i]; }
__global__ void k_pointwise_out(int n, const double *__restrict__ d, const double *__restrict__ x,
                                double *__restrict__ y)
{ GRID_STRIDE(i, n) y[i] = d[i] * x[i]; }

void vec_copy(Handle &H, int n, const double *x, double *y)
{
  if (n <= 0 || x == y) return;
  k_copy<<<vgrid(n), 256, 0, H.stream>>>(n, x, y); H.launches++;
}
void vec_zero(Handle &H, int n, double *x)
{
  if (n <= 0) return;
  k_zero<<<vgrid(n), 256, 0, H.stream>>>(n, x); H.launches++;
}
void vec_axpy(Handle &H, int n, double a, const double *x, double *y)
{
  if (n <= 0) return;
  k_axpy<<<vgrid(n), 256, 0, H.stream>>>(n, a, x, y); H.launches++;
}
void vec_axpy_dev(Handle &H, int n, const double *a_dev, double sign, const double *x, double *y)
{
  if (n <= 0) return;
  k_axpy_dev<<<vgrid(n), 256, 0, H.stream>>>(n, a_dev, sign, x, y); H.launches++;
}
void vec_sadd(Handle &H, int n, double s, double a, const double *x, double *y)
{
  if (n <= 0) return;
  k_sadd<<<vgrid(n), 256, 0, H.stream>>>(n, s, a, x, y); H.launches++;
}
void vec_scale(Handle &H, int n, double a, double *x)
{
  if (n <= 0) return;
  k_scale<<<vgrid(n), 256, 0, H.stream>>>(n, a, x); H.launches++;
}
void vec_scale_inv_dev(Handle &H, int n, const double *s_dev, double *x)
{
  if (n <= 0) return;
  k_scale_inv_dev<<<vgrid(n), 256, 0, H.stream>>>(n, s_dev, x); H.launches++;
}
void vec_pointwise(Handle &H, int n, const double *d, double *x)
{
  if (n <= 0) return;
  k_pointwise<<<vgrid(n), 256, 0, H.stream>>>(n, d, x); H.launches++;
}
void vec_pointwise_out(Handle &H, int n, const double *d, const double *x, double *y)
{
  if (n <= 0) return;
  k_pointwise_out<<<vgrid(n), 256, 0, H.stream>>>(n, d, x, y); H.launches++;
}

// Deterministic two-stage reduction: fixed grid, fixed per-block order, the last block to finish
// (ticket) adds the block partials in index order.  scratch layout: [0..1023] partials, then ticket.
constexpr int kRedBlocks = 888; // 148 * 6 (<= 1024 partial slots)
__device__ __forceinline__ double block_sum(double v)
{
  __shared__ double sh[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x < 32) {
    s = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  return s; // valid in thread 0
}

// ar.tab != nullptr (multi-rank, peer-memory transport): the last block also performs the all-reduce
// over the ranks -- it stores the local sum into every peer's mailbox slot over NVLink, waits for the
// peers and adds the partials in rank order (ar_exchange_block, nsb_internal.hpp).
__device__ __forceinline__ void finish_reduce(double part, double *partials, unsigned *ticket, double *out, ArArgs ar)
{
  __shared__ bool last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = part;
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    double v = 0.0;
    // fixed order: thread t sums partials t, t+256, ...; then block tree
    for (int b = threadIdx.x; b < int(gridDim.x); b += blockDim.x) v += __ldcg(partials + b);
    double s = block_sum(v);
    if (ar.tab) s = ar_exchange_block(ar.tab, ar.parity, ar.seq, s, threadIdx.x, 1);
    if (threadIdx.x == 0) { *out = s; *ticket = 0u; }
  }
}

// Both reductions keep 4 independent element groups in flight per thread (the loads of one group
// do not depend on the previous one), which is what a bandwidth-bound 16-24 B/element stream needs
// to cover the HBM latency at full occupancy; the summation order is fixed by (grid, block) only.
__global__ void __launch_bounds__(256) k_dot(int n, const double *__restrict__ x, const double *__restrict__ y,
                                             double *partials, unsigned *ticket, double *out, ArArgs ar)
{
  double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
  const int stride = gridDim.x * blockDim.x;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const double a0 = x[i], a1 = x[i + stride], a2 = x[i + 2 * stride], a3 = x[i + 3 * stride];
    const double b0 = y[i], b1 = y[i + stride], b2 = y[i + 2 * stride], b3 = y[i + 3 * stride];
    v0 += a0 * b0; v1 += a1 * b1; v2 += a2 * b2; v3 += a3 * b3;
  }
  for (; i < n; i += stride) v0 += x[i] * y[i];
  const double s = block_sum((v0 + v1) + (v2 + v3));
  finish_reduce(s, partials, ticket, out, ar);
}

__global__ void __launch_bounds__(256) k_add_and_dot(int n, double *vv, const double *__restrict__ a,
                                                     double sign, const double *__restrict__ vp,
                                                     const double *vn, double *partials,
                                                     unsigned *ticket, double *out, ArArgs ar)
{
  const double aa = sign * (*a);
  const bool self = (vn == vv);
  double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
  const int stride = gridDim.x * blockDim.x;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const double w0 = vv[i], w1 = vv[i + stride], w2 = vv[i + 2 * stride], w3 = vv[i + 3 * stride];
    const double p0 = vp[i], p1 = vp[i + stride], p2 = vp[i + 2 * stride], p3 = vp[i + 3 * stride];
    const double t0 = w0 + aa * p0, t1 = w1 + aa * p1, t2 = w2 + aa * p2, t3 = w3 + aa * p3;
    double o0 = t0, o1 = t1, o2 = t2, o3 = t3;
    if (!self) { o0 = vn[i]; o1 = vn[i + stride]; o2 = vn[i + 2 * stride]; o3 = vn[i + 3 * stride]; }
    vv[i] = t0; vv[i + stride] = t1; vv[i + 2 * stride] = t2; vv[i + 3 * stride] = t3;
    v0 += t0 * o0; v1 += t1 * o1; v2 += t2 * o2; v3 += t3 * o3;
  }
  for (; i < n; i += stride) {
    const double t = vv[i] + aa * vp[i];
    const double o = self ? t : vn[i];
    vv[i] = t;
    v0 += t * o;
  }
  const double s = block_sum((v0 + v1) + (v2 + v3));
  finish_reduce(s, partials, ticket, out, ar);
}

static inline unsigned rgrid(int n) { return unsigned(std::max(1, std::min((n + 255) / 256, kRedBlocks))); }

void vec_dot_dev(Handle &H, int n, const double *x, const double *y, double *out_dev, ArArgs ar)
{
  double *partials = H.d_scratch.p + 64;
  unsigned *ticket = reinterpret_cast<unsigned *>(H.d_scratch.p + 64 + 1024);
  k_dot<<<rgrid(n), 256, 0, H.stream>>>(std::max(n, 0), x, y, partials, ticket, out_dev, ar);
  H.launches++;
}

void vec_add_and_dot_dev(Handle &H, int n, double *vv, const double *a_dev, double sign, const double *v_prev,
                         const double *v_next, double *out_dev, ArArgs ar)
{
  double *partials = H.d_scratch.p + 64;
  unsigned *ticket = reinterpret_cast<unsigned *>(H.d_scratch.p + 64 + 1024);
  k_add_and_dot<<<rgrid(n), 256, 0, H.stream>>>(std::max(n, 0), vv, a_dev, sign, v_prev, v_next, partials, ticket,
                                                out_dev, ar);
  H.launches++;
}

// --------------------------------------------------------------------------------------------
// Batched classical Gram-Schmidt (nsb_params.orthogonalisation = 1): the coefficients of up to 8
// basis vectors, and |vv|^2, from one pass over vv; then one fused update + norm pass.
// --------------------------------------------------------------------------------------------
constexpr int kMdBlocks = 592; // 148 * 4: NV + 1 independent loads per thread already cover the latency

template <int NV>
__global__ void __launch_bounds__(256) k_multi_dot(int n, const double *__restrict__ vv, const double *__restrict__ V,
                                                   size_t ld, int with_self, double *partials, unsigned *ticket,
                                                   double *out, double *out_self, ArArgs ar)
{
  double acc[NV + 1];
#pragma unroll
  for (int j = 0; j <= NV; ++j) acc[j] = 0.0;
  GRID_STRIDE(i, n) {
    const double w = vv[i];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] += w * V[size_t(j) * ld + i];
    acc[NV] += w * w;
  }
  __shared__ bool last;
#pragma unroll
  for (int j = 0; j <= NV; ++j) {
    const double s = block_sum(acc[j]);
    if (threadIdx.x == 0) partials[j * kMdBlocks + blockIdx.x] = s;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __shared__ double fin[NV + 1];
  for (int j = 0; j <= NV; ++j) {
    double v = 0.0;
    for (int b = threadIdx.x; b < int(gridDim.x); b += blockDim.x) v += __ldcg(partials + j * kMdBlocks + b);
    const double s = block_sum(v);
    if (threadIdx.x == 0) fin[j] = s;
  }
  __syncthreads();
  const int j = threadIdx.x;
  double s = j <= NV ? fin[j] : 0.0;
  // all NV + 1 values cross the ranks in ONE exchange by the same block that finished the reduction
  if (ar.tab) s = ar_exchange_block(ar.tab, ar.parity, ar.seq, s, j, NV + 1);
  if (j < NV) out[j] = s;
  else if (j == NV && with_self) *out_self = s;
  if (threadIdx.x == 0) *ticket = 0u;
}

// vv -= sum_j h[j] V_j ; optionally out_norm2 = |vv|^2 of the result
template <int NV>
__global__ void __launch_bounds__(256) k_multi_axpy(int n, double *__restrict__ vv, const double *__restrict__ V,
                                                    size_t ld, const double *__restrict__ h, int with_norm,
                                                    double *partials, unsigned *ticket, double *out_norm2, ArArgs ar)
{
  double hh[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) hh[j] = h[j];
  double acc = 0.0;
  GRID_STRIDE(i, n) {
    double w = vv[i];
#pragma unroll
    for (int j = 0; j < NV; ++j) w -= hh[j] * V[size_t(j) * ld + i];
    vv[i] = w;
    acc += w * w;
  }
  if (!with_norm) return;
  const double s = block_sum(acc);
  finish_reduce(s, partials, ticket, out_norm2, ar);
}

template <int NV>
static void multi_dot_t(Handle &H, int n, const double *vv, const double *V, size_t ld, bool with_self, double *out,
                        double *out_self, ArArgs ar)
{
  double *partials = H.d_scratch.p + 64 + 1024 + 8;
  unsigned *ticket = reinterpret_cast<unsigned *>(H.d_scratch.p + 64 + 1024 + 8 + 9 * 1024);
  const unsigned grid = unsigned(std::max(1, std::min((n + 255) / 256, kMdBlocks)));
  k_multi_dot<NV><<<grid, 256, 0, H.stream>>>(std::max(n, 0), vv, V, ld, with_self ? 1 : 0, partials, ticket, out, out_self, ar);
  H.launches++;
}
template <int NV>
static void multi_axpy_t(Handle &H, int n, double *vv, const double *V, size_t ld, const double *h, bool with_norm,
                         double *out_norm2, ArArgs ar)
{
  double *partials = H.d_scratch.p + 64;
  unsigned *ticket = reinterpret_cast<unsigned *>(H.d_scratch.p + 64 + 1024);
  k_multi_axpy<NV><<<rgrid(n), 256, 0, H.stream>>>(std::max(n, 0), vv, V, ld, h, with_norm ? 1 : 0, partials, ticket, out_norm2, ar);
  H.launches++;
}

// h[j] = vv . V_j for j < nv (V_j = V + j*ld), *self = vv . vv; groups of 8 vectors per pass over vv
void vec_multi_dot_dev(Handle &H, int n, const double *vv, const double *V, size_t ld, int nv, double *h_dev,
                       double *self_dev, bool allreduce)
{
  for (int j0 = 0; j0 < nv; j0 += 8) {
    const int g = std::min(8, nv - j0);
    const bool self = (j0 == 0) && self_dev;
    const double *Vg = V + size_t(j0) * ld;
    const ArArgs ar = allreduce ? halo_ar_args(H) : ArArgs{nullptr, 0, 0};
    switch (g) {
      case 1: multi_dot_t<1>(H, n, vv, Vg, ld, self, h_dev + j0, self_dev, ar); break;
      case 2: multi_dot_t<2>(H, n, vv, Vg, ld, self, h_dev + j0, self_dev, ar); break;
      case 3: multi_dot_t<3>(H, n, vv, Vg, ld, self, h_dev + j0, self_dev, ar); break;
      case 4: multi_dot_t<4>(H, n, vv, Vg, ld, self, h_dev + j0, self_dev, ar); break;
      case 5: multi_dot_t<5>(H, n, vv, Vg, ld, self, h_dev + j0, self_dev, ar); break;
      case 6: multi_dot_t<6>(H, n, vv, Vg, ld, self, h_dev + j0, self_dev, ar); break;
      case 7: multi_dot_t<7>(H, n, vv, Vg, ld, self, h_dev + j0, self_dev, ar); break;
      default: multi_dot_t<8>(H, n, vv, Vg, ld, self, h_dev + j0, self_dev, ar); break;
    }
  }
}

// vv -= sum_{j<nv} h[j] V_j ; *norm2_dev = |vv|^2 afterwards (computed by the last group)
void vec_multi_axpy_dev(Handle &H, int n, double *vv, const double *V, size_t ld, int nv, const double *h_dev,
                        double *norm2_dev, bool allreduce)
{
  for (int j0 = 0; j0 < nv; j0 += 8) {
    const int g = std::min(8, nv - j0);
    const bool nrm = (j0 + 8 >= nv) && norm2_dev;
    const double *Vg = V + size_t(j0) * ld;
    const ArArgs ar = (allreduce && nrm) ? halo_ar_args(H) : ArArgs{nullptr, 0, 0};
    switch (g) {
      case 1: multi_axpy_t<1>(H, n, vv, Vg, ld, h_dev + j0, nrm, norm2_dev, ar); break;
      case 2: multi_axpy_t<2>(H, n, vv, Vg, ld, h_dev + j0, nrm, norm2_dev, ar); break;
      case 3: multi_axpy_t<3>(H, n, vv, Vg, ld, h_dev + j0, nrm, norm2_dev, ar); break;
      case 4: multi_axpy_t<4>(H, n, vv, Vg, ld, h_dev + j0, nrm, norm2_dev, ar); break;
      case 5: multi_axpy_t<5>(H, n, vv, Vg, ld, h_dev + j0, nrm, norm2_dev, ar); break;
      case 6: multi_axpy_t<6>(H, n, vv, Vg, ld, h_dev + j0, nrm, norm2_dev, ar); break;
      case 7: multi_axpy_t<7>(H, n, vv, Vg, ld, h_dev + j0, nrm, norm2_dev, ar); break;
      default: multi_axpy_t<8>(H, n, vv, Vg, ld, h_dev + j0, nrm, norm2_dev, ar); break;
    }
  }
}

// --------------------------------------------------------------------------------------------
// ILU(0): Ifpack_ILU::Compute / Solve on the device, level scheduled
// --------------------------------------------------------------------------------------------
static void level_schedule(int n, const std::vector<int> &rowptr, const std::vector<int> &colind, bool forward,
                           std::vector<int> &lvl_ptr, std::vector<int> &lvl_rows)
{
  std::vector<int> level(n, 0);
  int maxl = 0;
  if (forward) {
    for (int i = 0; i < n; ++i) {
      int l = 0;
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int j = colind[k];
        if (j < i) l = std::max(l, level[j] + 1);
      }
      level[i] = l;
      maxl = std::max(maxl, l);
    }
  } else {
    for (int i = n - 1; i >= 0; --i) {
      int l = 0;
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int j = colind[k];
        if (j > i) l = std::max(l, level[j] + 1);
      }
      level[i] = l;
      maxl = std::max(maxl, l);
    }
  }
  const int nl = n ? maxl + 1 : 0;
  lvl_ptr.assign(nl + 1, 0);
  for (int i = 0; i < n; ++i) lvl_ptr[level[i] + 1]++;
  for (int l = 0; l < nl; ++l) lvl_ptr[l + 1] += lvl_ptr[l];
  lvl_rows.resize(n);
  std::vector<int> pos(lvl_ptr.begin(), lvl_ptr.end() - (nl ? 1 : 0));
  if (forward)
    for (int i = 0; i < n; ++i) lvl_rows[pos[level[i]]++] = i;
  else
    for (int i = n - 1; i >= 0; --i) lvl_rows[pos[level[i]]++] = i;
}

// Greedy distance-1 colouring of the (structurally symmetric) owned-owned graph in natural row
// order; the ILU ordering is "colour by colour, natural order inside a colour".
// Greedy multicolouring in natural order, then rows sorted by (chunk, colour, natural index) where a
// chunk is a run of `chunk_rows` consecutive rows.  Every (chunk, colour) group is an independent set
// and only depends on groups before it, so the groups are the levels of both triangular solves.
// Chunking was meant to keep the part of the vector a group gathers from resident in L2 across the
// colour sweeps (without it every sweep re-reads the gathered nodes from HBM; ncu, profiles/: 6.7 GB
// of DRAM traffic for 2.5 GB of algorithmic bytes); see ilu_chunk_rows for why it is off by default.
static void multicolour_order(int n, const Csr &A, int n_owned_cols, int chunk_rows, std::vector<int> &order,
                              std::vector<int> &colour_ptr)
{
  std::vector<int> colour(n, -1), mark;
  int ncol = 0;
  for (int i = 0; i < n; ++i) {
    mark.assign(ncol + 1, 0);
    for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
      const int j = A.colind[k];
      if (j < n_owned_cols && j != i && colour[j] >= 0) mark[colour[j]] = 1;
    }
    int c = 0;
    while (mark[c]) ++c;
    colour[i] = c;
    if (c == ncol) ++ncol;
  }
  if (chunk_rows <= 0 || chunk_rows > n) chunk_rows = std::max(n, 1);
  const int nchunks = (n + chunk_rows - 1) / chunk_rows;
  const size_t ngroups = size_t(nchunks) * ncol;
  std::vector<int> cnt(ngroups + 1, 0);
  auto group = [&](int i) { return size_t(i / chunk_rows) * ncol + colour[i]; };
  for (int i = 0; i < n; ++i) cnt[group(i) + 1]++;
  for (size_t g = 0; g < ngroups; ++g) cnt[g + 1] += cnt[g];
  std::vector<int> fill(cnt.begin(), cnt.end() - 1);
  order.resize(n);
  for (int i = 0; i < n; ++i) order[fill[group(i)]++] = i;
  // drop empty groups
  colour_ptr.assign(1, 0);
  for (size_t g = 0; g < ngroups; ++g)
    if (cnt[g + 1] > cnt[g]) colour_ptr.push_back(cnt[g + 1]);
  if (colour_ptr.size() == 1) colour_ptr.push_back(0);
}


// Block multicolour ordering (algebraic block multi-colouring, Iwashita et al.): the rows are aggregated
// into blocks of kBlkRows spatially adjacent rows (greedy breadth-first aggregation over the matrix
// graph, seeds in natural order; a block that runs out of frontier is topped up from the next seed), the
// BLOCK graph is coloured greedily, and the ILU ordering is (block colour, block, position in the
// block).  Blocks of one colour do not touch, so a colour is solved by one launch with one warp per
// block; inside a block the rows are eliminated sequentially (a lane per row, values handed on by
// shuffles, kernels_sell.cu).  Against point multicolouring this (a) keeps most couplings inside a
// block or between neighbouring blocks, so a triangular sweep gathers each vector entry from HBM ~5
// instead of ~14 times, and (b) is a much stronger ILU(0): inner iteration counts equal those of the
// natural ordering (measured with the oracle, profiles/README.md).
constexpr int kBlkRows = 32;
static void block_multicolour_order(int n, const Csr &A, int n_owned_cols, std::vector<int> &order,
                                    std::vector<int> &colour_ptr, std::vector<int> &blk_ptr,
                                    std::vector<int> &colour_blk)
{
  const int nc_lim = std::min(n, n_owned_cols);
  std::vector<int> blk_of(n, -1), seq;
  seq.reserve(n);
  std::vector<int> q;
  int nb = 0, cur = 0;
  for (int seed = 0; seed < n; ++seed) {
    if (blk_of[seed] >= 0) continue;
    q.clear();
    q.push_back(seed);
    blk_of[seed] = nb; seq.push_back(seed);
    if (++cur == kBlkRows) { ++nb; cur = 0; continue; }
    size_t head = 0;
    bool full = false;
    while (head < q.size() && !full) {
      const int u = q[head++];
      for (int k = A.rowptr[u]; k < A.rowptr[u + 1]; ++k) {
        const int v = A.colind[k];
        if (v >= nc_lim || blk_of[v] >= 0) continue;
        blk_of[v] = nb; seq.push_back(v); q.push_back(v);
        if (++cur == kBlkRows) { ++nb; cur = 0; full = true; break; }
      }
    }
  }
  if (cur > 0) ++nb;
  // rows of each block in aggregation order
  std::vector<int> bstart(nb + 1, 0);
  for (int i = 0; i < n; ++i) bstart[blk_of[i] + 1]++;
  for (int b = 0; b < nb; ++b) bstart[b + 1] += bstart[b];
  // (seq lists the blocks one after the other already: block b = seq[bstart[b] .. bstart[b+1]) )
  // greedy colouring of the block graph in block order
  std::vector<int> bcol(nb, -1), mark;
  int ncol = 0;
  for (int b = 0; b < nb; ++b) {
    mark.assign(ncol + 1, 0);
    for (int t = bstart[b]; t < bstart[b + 1]; ++t) {
      const int u = seq[t];
      for (int k = A.rowptr[u]; k < A.rowptr[u + 1]; ++k) {
        const int v = A.colind[k];
        if (v >= nc_lim) continue;
        const int bb = blk_of[v];
        if (bb != b && bcol[bb] >= 0) mark[bcol[bb]] = 1;
      }
    }
    int c = 0;
    while (mark[c]) ++c;
    bcol[b] = c;
    if (c == ncol) ++ncol;
  }
  // factor order: colour by colour, blocks in index order, rows in aggregation order
  std::vector<int> ccnt(ncol + 1, 0);
  for (int b = 0; b < nb; ++b) ccnt[bcol[b] + 1]++;
  for (int c = 0; c < ncol; ++c) ccnt[c + 1] += ccnt[c];
  colour_blk.assign(ccnt.begin(), ccnt.end());
  std::vector<int> slot(ccnt.begin(), ccnt.end() - 1), blocks(nb);
  for (int b = 0; b < nb; ++b) blocks[slot[bcol[b]]++] = b;
  order.clear();
  order.reserve(n);
  blk_ptr.assign(1, 0);
  colour_ptr.assign(1, 0);
  for (int c = 0; c < ncol; ++c) {
    for (int k = ccnt[c]; k < ccnt[c + 1]; ++k) {
      const int b = blocks[k];
      for (int t = bstart[b]; t < bstart[b + 1]; ++t) order.push_back(seq[t]);
      blk_ptr.push_back(int(order.size()));
    }
    colour_ptr.push_back(int(order.size()));
  }
  if (colour_ptr.size() == 1) colour_ptr.push_back(0);
}

// ---------------------------------------------------------------------------------------------
// Subdomain ordering (ilu_ordering = 3): a two-level ordering for triangular solves whose working
// vector lives in SHARED MEMORY (kernels_sd.cu).
//   1. The rows are cut into compact "parts" of <= leaf_max rows by recursive coordinate bisection of
//      their support points.
//   2. A row is a SEPARATOR row when it couples with a row of a lower-numbered part; all other rows
//      are INTERIOR rows, and interior rows of different parts never couple.
//   3. Factor order = [interior of part 0 | interior of part 1 | ... | separators]; inside a part and
//      inside the separator set the rows are multicoloured (greedy, natural order) and sorted by
//      (colour, row length): a colour is an independent set.
// It is an exact ILU(0) of the same matrix in yet another elimination order.  In the forward solve
// an interior row only reads rows of its own part, so one CTA solves a part start to finish with
// the part's slice of the vector in shared memory -- no re-gathering of the vector from HBM between
// colours (the point-multicolour sweeps moved 2.6x their algorithmic bytes, profiles/r01_traffic.json)
// and ONE launch for ~85% of the rows instead of one per colour; the separator rows (~15%) follow as
// colour sweeps over the global staging vector.  Backward: separators first, then the parts, which
// also stage the separator values they couple with (their "ring").
struct SdOrder {
  std::vector<int> order;          // factor row -> matrix row
  std::vector<int> part_ptr;       // [n_parts + 1] factor rows of each part's interior
  std::vector<int> pcol_ptr, pcol; // per part: factor-row boundaries of its colours (pcol[pcol_ptr[p] ..]), first = part_ptr[p]
  std::vector<int> sep_colour_ptr; // factor-row boundaries of the separator colours, first = n_interior, last = n
  std::vector<int> level_part_ptr; // [n_levels + 1] parts of each level
};

// recursive coordinate bisection of idx[lo, hi) into k leaves of (nearly) equal size
static void rcb_split(int *idx, int lo, int hi, const double *xyz, int gdim, int k, std::vector<std::pair<int, int>> &leaves)
{
  if (k <= 1 || hi - lo <= 1) {
#pragma omp critical(nsb_rcb_leaves)
    leaves.emplace_back(lo, hi);
    return;
  }
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int t = lo; t < hi; ++t)
    for (int d = 0; d < gdim; ++d) {
      const double v = xyz[size_t(idx[t]) * gdim + d];
      mn[d] = std::min(mn[d], v); mx[d] = std::max(mx[d], v);
    }
  int dd = 0;
  for (int d = 1; d < gdim; ++d)
    if (mx[d] - mn[d] > mx[dd] - mn[dd]) dd = d;
  const int k1 = k / 2;
  const int mid = lo + int(int64_t(hi - lo) * k1 / k);
  std::nth_element(idx + lo, idx + mid, idx + hi, [&](int a, int b) {
    const double xa = xyz[size_t(a) * gdim + dd], xb = xyz[size_t(b) * gdim + dd];
    return xa < xb || (xa == xb && a < b);
  });
  [[maybe_unused]] const bool big = hi - lo > 200000;
#pragma omp task default(shared) if (big)
  rcb_split(idx, lo, mid, xyz, gdim, k1, leaves);
#pragma omp task default(shared) if (big)
  rcb_split(idx, mid, hi, xyz, gdim, k - k1, leaves);
#pragma omp taskwait
}

// Multi-level: the separator rows of level l are the active rows of level l + 1 and are cut into parts
// again (they form sheets between the parts of level l, so their own separators are nearly 1D and few);
// what is left after the last level is multicoloured globally.  Parts of one level never couple; a part
// of level l couples with rows of earlier levels (forward solve: its "lower ring") and of later levels
// (backward solve: its "upper ring").
static void subdomain_order(int n, const Csr &A, int n_owned_cols, const double *xyz, int gdim, const std::vector<int> &leaf_max,
                            int min_active, SdOrder &out)
{
  const int nc_lim = std::min(n, n_owned_cols);
  auto deg = [&](int i) { return A.rowptr[i + 1] - A.rowptr[i]; };
  std::vector<int> active(n), part(n, -1), colour(n, -1);
  std::vector<char> is_active(n, 1), sep(n, 0);
  for (int i = 0; i < n; ++i) active[i] = i;
  out.order.clear();
  out.order.reserve(n);
  out.part_ptr.assign(1, 0);
  out.pcol_ptr.assign(1, 0);
  out.pcol.clear();
  out.level_part_ptr.assign(1, 0);
  for (size_t level = 0; level < leaf_max.size(); ++level) {
    const int na = int(active.size());
    if (na == 0 || (level > 0 && na < min_active)) break;
    // 1. parts of the active rows
    std::vector<int> &idx = active;
    std::vector<std::pair<int, int>> leaves;
    if (xyz) {
#pragma omp parallel
#pragma omp single
      rcb_split(idx.data(), 0, na, xyz, gdim, (na + leaf_max[level] - 1) / leaf_max[level], leaves);
      std::sort(leaves.begin(), leaves.end());
    } else // no geometry: runs of consecutive rows
      for (int lo = 0; lo < na; lo += leaf_max[level]) leaves.emplace_back(lo, std::min(na, lo + leaf_max[level]));
    const int np = int(leaves.size());
#pragma omp parallel for schedule(dynamic, 16)
    for (int p = 0; p < np; ++p) {
      std::sort(idx.begin() + leaves[p].first, idx.begin() + leaves[p].second); // natural order inside a part
      for (int k = leaves[p].first; k < leaves[p].second; ++k) part[idx[k]] = p;
    }
    // 2. separators: active rows coupling with an active row of a lower-numbered part
#pragma omp parallel for schedule(static)
    for (int t = 0; t < na; ++t) {
      const int i = idx[t];
      sep[i] = 0;
      for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
        const int j = A.colind[k];
        if (j < nc_lim && is_active[j] && part[j] < part[i]) { sep[i] = 1; break; }
      }
    }
    // 3. interiors: greedy colouring inside each part (interior rows of different parts never couple)
    std::vector<std::vector<int>> pint(np), pcb(np); // interior rows in factor order, colour boundaries (local)
#pragma omp parallel
    {
      std::vector<int> mark, rows;
#pragma omp for schedule(dynamic, 16)
      for (int p = 0; p < np; ++p) {
        rows.clear();
        int ncol = 0;
        for (int k = leaves[p].first; k < leaves[p].second; ++k) {
          const int i = idx[k];
          if (sep[i]) continue;
          mark.assign(ncol + 1, 0);
          for (int e = A.rowptr[i]; e < A.rowptr[i + 1]; ++e) {
            const int j = A.colind[e];
            // an active interior neighbour lies in the same part; its colour is only set by this thread
            if (j < nc_lim && j != i && is_active[j] && !sep[j] && part[j] == p && colour[j] >= 0) mark[colour[j]] = 1;
          }
          int c = 0;
          while (mark[c]) ++c;
          colour[i] = c;
          if (c == ncol) ++ncol;
          rows.push_back(i);
        }
        std::sort(rows.begin(), rows.end(), [&](int a, int b) {
          if (colour[a] != colour[b]) return colour[a] < colour[b];
          if (deg(a) != deg(b)) return deg(a) > deg(b);
          return a < b;
        });
        pint[p] = rows;
        pcb[p].assign(1, 0);
        for (size_t t = 1; t <= rows.size(); ++t)
          if (t == rows.size() || colour[rows[t]] != colour[rows[t - 1]]) pcb[p].push_back(int(t));
      }
    }
    // 4. append the parts of this level; the separators stay active
    std::vector<int> next;
    for (int p = 0; p < np; ++p) {
      for (int k = leaves[p].first; k < leaves[p].second; ++k)
        if (sep[idx[k]]) next.push_back(idx[k]);
      if (pint[p].empty()) continue; // a part made of separator rows only
      const int base = int(out.order.size());
      for (int b : pcb[p]) out.pcol.push_back(base + b);
      out.pcol_ptr.push_back(int(out.pcol.size()));
      out.order.insert(out.order.end(), pint[p].begin(), pint[p].end());
      out.part_ptr.push_back(int(out.order.size()));
      for (int i : pint[p]) is_active[i] = 0;
    }
    out.level_part_ptr.push_back(int(out.part_ptr.size()) - 1);
    active.swap(next);
  }
  // 5. what is left: greedy colouring of the induced subgraph, colour by colour
  {
    std::vector<int> mark;
    int ncol = 0;
    for (int i : active) colour[i] = -1;
    for (int i : active) {
      mark.assign(ncol + 1, 0);
      for (int e = A.rowptr[i]; e < A.rowptr[i + 1]; ++e) {
        const int j = A.colind[e];
        if (j < nc_lim && j != i && is_active[j] && colour[j] >= 0) mark[colour[j]] = 1;
      }
      int c = 0;
      while (mark[c]) ++c;
      colour[i] = c;
      if (c == ncol) ++ncol;
    }
    std::stable_sort(active.begin(), active.end(), [&](int a, int b) { return colour[a] < colour[b]; });
  }
  out.sep_colour_ptr.assign(1, int(out.order.size()));
  for (size_t t = 0; t < active.size(); ++t) {
    out.order.push_back(active[t]);
    if (t + 1 == active.size() || colour[active[t + 1]] != colour[active[t]]) out.sep_colour_ptr.push_back(int(out.order.size()));
  }
  if (out.sep_colour_ptr.size() == 1) out.sep_colour_ptr.push_back(int(out.order.size()));
  if (int(out.order.size()) != n) throw StateError("subdomain ordering lost rows");
}

// rows per chunk of the multicolour ordering: NSB_ILU_CHUNK (rows), default 0 = one chunk.  Measured
// at 19.9 M DoF (profiles/README.md): chunks of 0.8 / 1.25 / 2.5 M nodes make the F_s apply SLOWER
// (3.3 / 2.9 / 2.2 ms against 1.68 ms unchunked) -- the 5x more, 5x smaller sweeps are latency-bound
// and lose more than the L2 residency of the gathered vector gains -- so chunking stays off.
static int ilu_chunk_rows(int bs_rhs)
{
  const char *e = getenv("NSB_ILU_CHUNK");
  (void)bs_rhs;
  return e ? atoi(e) : 0;
}

// ordering: 0 = natural local row order (what Ifpack does in the reference), 1 = multicolour.
// The factor lives in the permuted index space: row k of the factor is row order[k] of A.
// pattern of the factors in the permuted index space: row k = row order[k] of A, columns renumbered and
// sorted, off-process columns dropped (Ifpack_LocalFilter); src[e]: position of entry e in A.
static void permute_pattern(const Csr &A, const std::vector<int> &order, int n_owned_cols, std::vector<int> &rowptr,
                            std::vector<int> &colind, std::vector<int> &src, std::vector<int> &diagpos)
{
  const int n = A.n_rows;
  const int lim = std::min(n_owned_cols, n); // Ifpack_LocalFilter: off-process columns are dropped
  std::vector<int> pos(n_owned_cols > n ? n_owned_cols : n, -1);
#pragma omp parallel for schedule(static)
  for (int k = 0; k < n; ++k) pos[order[k]] = k;
  rowptr.assign(n + 1, 0);
  diagpos.assign(n, 0);
  // rows are independent: count, prefix sum, then fill and sort every row on its own (all host threads)
  int missing_diag = 0;
#pragma omp parallel for schedule(static) reduction(+ : missing_diag)
  for (int k = 0; k < n; ++k) {
    const int i = order[k];
    int cnt = 0;
    bool have_diag = false;
    for (int e = A.rowptr[i]; e < A.rowptr[i + 1]; ++e) {
      const int j = A.colind[e];
      if (j >= lim) continue;
      if (j == i) have_diag = true;
      ++cnt;
    }
    if (!have_diag) ++missing_diag;
    rowptr[k + 1] = cnt;
  }
  if (missing_diag) throw StateError("ILU: structurally missing diagonal");
  for (int k = 0; k < n; ++k) rowptr[k + 1] += rowptr[k];
  reserve_prefaulted(colind, size_t(rowptr[n]));
  reserve_prefaulted(src, size_t(rowptr[n]));
  colind.assign(size_t(rowptr[n]), 0);
  src.assign(size_t(rowptr[n]), 0);
#pragma omp parallel
  {
    std::vector<std::pair<int, int>> row;
#pragma omp for schedule(static)
    for (int k = 0; k < n; ++k) {
      const int i = order[k];
      row.clear();
      for (int e = A.rowptr[i]; e < A.rowptr[i + 1]; ++e) {
        const int j = A.colind[e];
        if (j >= lim) continue;
        row.emplace_back(pos[j], e);
      }
      std::sort(row.begin(), row.end());
      int o = rowptr[k];
      for (auto &pe : row) {
        if (pe.first == k) diagpos[k] = o;
        colind[o] = pe.first;
        src[o] = pe.second;
        ++o;
      }
    }
  }
}

// CPU-only check of the subdomain ordering and its packed storage (tests/test_host_cpu.py)
double sd_host_check(int n, const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
                     const std::vector<int> &order, const std::vector<int> &part_ptr, const std::vector<int> &pcol_ptr,
                     const std::vector<int> &pcol, const std::vector<int> &sep_colour_ptr, const std::vector<int> &level_part_ptr,
                     int bs, int *stats);
double sd_debug_check(const Csr &A, const double *xyz, int gdim, const int *leaf_levels, int min_active, int bs, int *stats,
                      int *order_out)
{
  SdOrder o;
  std::vector<int> leaves;
  for (int k = 0; k < 3 && leaf_levels[k] > 0; ++k) leaves.push_back(leaf_levels[k]);
  subdomain_order(A.n_rows, A, A.n_rows, xyz, gdim, leaves, min_active, o);
  std::vector<int> rowptr, colind, src, diagpos;
  permute_pattern(A, o.order, A.n_rows, rowptr, colind, src, diagpos);
  // a part's rows may only couple with their own part and with rows of OTHER levels
  const int np = int(o.part_ptr.size()) - 1;
  for (size_t l = 0; l + 1 < o.level_part_ptr.size(); ++l) {
    const int lv0 = o.part_ptr[o.level_part_ptr[l]], lv1 = o.part_ptr[o.level_part_ptr[l + 1]];
    for (int p = o.level_part_ptr[l]; p < o.level_part_ptr[l + 1]; ++p)
      for (int r = o.part_ptr[p]; r < o.part_ptr[p + 1]; ++r)
        for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) {
          const int c = colind[e];
          if (!((c >= o.part_ptr[p] && c < o.part_ptr[p + 1]) || c < lv0 || c >= lv1)) return 3e30;
        }
  }
  (void)np;
  if (order_out) std::copy(o.order.begin(), o.order.end(), order_out);
  return sd_host_check(A.n_rows, rowptr, colind, diagpos, o.order, o.part_ptr, o.pcol_ptr, o.pcol, o.sep_colour_ptr,
                       o.level_part_ptr, bs, stats);
}

// CPU-only check of the block multicolour ordering and its packed storage (tests/test_host_cpu.py)
double bsell_host_check(const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
                        const std::vector<int> &blk_ptr, const std::vector<int> &colour_blk, int bs, int xcap, int *stats);
double bsell_debug_check(const Csr &A, int bs, int xcap, int *stats, int *order_out)
{
  std::vector<int> order, colour_ptr, blk_ptr, colour_blk;
  block_multicolour_order(A.n_rows, A, A.n_rows, order, colour_ptr, blk_ptr, colour_blk);
  std::vector<int> rowptr, colind, src, diagpos;
  permute_pattern(A, order, A.n_rows, rowptr, colind, src, diagpos);
  // blocks of one colour must not touch, and a row may only couple with rows of other colours or of its own block
  std::vector<int> blk_of(A.n_rows), col_of_blk(blk_ptr.size() - 1);
  for (size_t c = 0; c + 1 < colour_blk.size(); ++c)
    for (int b = colour_blk[c]; b < colour_blk[c + 1]; ++b) col_of_blk[b] = int(c);
  for (size_t b = 0; b + 1 < blk_ptr.size(); ++b)
    for (int r = blk_ptr[b]; r < blk_ptr[b + 1]; ++r) blk_of[r] = int(b);
  for (int r = 0; r < A.n_rows; ++r)
    for (int e = rowptr[r]; e < rowptr[r + 1]; ++e)
      if (blk_of[colind[e]] != blk_of[r] && col_of_blk[blk_of[colind[e]]] == col_of_blk[blk_of[r]]) return 1.5e30;
  if (order_out) std::copy(order.begin(), order.end(), order_out);
  return bsell_host_check(rowptr, colind, diagpos, blk_ptr, colour_blk, bs, xcap, stats);
}

// CPU-only check of the point multicolour ordering and the SELL-32 storage of its factors (tests/test_host_cpu.py)
double sell_host_check(const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
                       const std::vector<int> &colour_ptr, int bs, int lanes, int window, int *stats);
double sell_debug_check(const Csr &A, int bs, int lanes, int window, int *stats, int *order_out)
{
  std::vector<int> order, colour_ptr;
  multicolour_order(A.n_rows, A, A.n_rows, 0, order, colour_ptr);
  std::vector<int> rowptr, colind, src, diagpos;
  permute_pattern(A, order, A.n_rows, rowptr, colind, src, diagpos);
  // a colour is an independent set: a row only couples with itself inside its colour
  for (size_t c = 0; c + 1 < colour_ptr.size(); ++c)
    for (int r = colour_ptr[c]; r < colour_ptr[c + 1]; ++r)
      for (int e = rowptr[r]; e < rowptr[r + 1]; ++e)
        if (colind[e] != r && colind[e] >= colour_ptr[c] && colind[e] < colour_ptr[c + 1]) return 1.5e30;
  if (order_out) std::copy(order.begin(), order.end(), order_out);
  return sell_host_check(rowptr, colind, diagpos, colour_ptr, bs, lanes, window, stats);
}

static std::vector<int> sd_leaf_rows(int bs_rhs)
{ // rows per part and level: a part's rows + ring times bs_rhs doubles should leave room for two CTAs per SM.
  // NSB_SD_LEAF = "l1[,l2[,l3]]" overrides (one entry = one level).
  std::vector<int> v;
  if (const char *e = getenv("NSB_SD_LEAF")) {
    for (const char *p = e; *p;) {
      const int x = atoi(p);
      if (x > 0) v.push_back(x);
      while (*p && *p != ',') ++p;
      if (*p == ',') ++p;
    }
    if (!v.empty()) return v;
  }
  const int l1 = bs_rhs == 3 ? 3072 : bs_rhs == 2 ? 4096 : 8192;
  return {l1, l1 / 4, l1 / 4};
}
static int sd_min_active()
{ // fewer active rows than this: no further level, colour them
  const char *e = getenv("NSB_SD_MIN_ACTIVE");
  return e ? atoi(e) : 20000;
}

void ilu_build(Handle &H, DevIlu &ilu, const Csr &A, int n_owned_cols, int bs_rhs, int ordering, const double *xyz, int gdim)
{
  const int n = A.n_rows;
  const bool verbose = getenv("NSB_VERBOSE") && atoi(getenv("NSB_VERBOSE")) > 1;
  auto t_last = std::chrono::steady_clock::now();
  auto phase = [&](const char *what) {
    if (!verbose) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[nsb ilu_build] %-24s %8.2f s\n", what, std::chrono::duration<double>(now - t_last).count());
    t_last = now;
  };
  ilu.n = n;
  ilu.bs_rhs = bs_rhs;
  ilu.h_order.clear();
  std::vector<int> colour_ptr, blk_ptr, colour_blk;
  SdOrder sdo;
  if (ordering == 1) multicolour_order(n, A, n_owned_cols, ilu_chunk_rows(bs_rhs), ilu.h_order, colour_ptr);
  else if (ordering == 2) block_multicolour_order(n, A, n_owned_cols, ilu.h_order, colour_ptr, blk_ptr, colour_blk);
  else if (ordering == 3) {
    // parts as large as shared memory allows: shrink the leaves until every part (rows + ring) fits
    std::vector<int> leaves = sd_leaf_rows(bs_rhs);
    for (int attempt = 0;; ++attempt) {
      subdomain_order(n, A, n_owned_cols, xyz, gdim, leaves, sd_min_active(), sdo);
      std::vector<int> rp, ci, sr, dp;
      permute_pattern(A, sdo.order, n_owned_cols, rp, ci, sr, dp);
      if (sd_smem_needed(rp, ci, dp, sdo.part_ptr, bs_rhs) <= size_t(220) * 1024 || attempt == 6) break;
      for (int &l : leaves) l = std::max(32, l * 3 / 4);
    }
    ilu.h_order = sdo.order;
  }
  else { ilu.h_order.resize(n); for (int i = 0; i < n; ++i) ilu.h_order[i] = i; }
  const std::vector<int> &order = ilu.h_order;
  phase("ordering");
  std::vector<int> rowptr, colind, src, diagpos;
  permute_pattern(A, order, n_owned_cols, rowptr, colind, src, diagpos);
  phase("permuted pattern");
  ilu.nnz = int64_t(colind.size());
  ilu.rowptr.upload(rowptr);
  ilu.colind.upload(colind);
  ilu.src.upload(src);
  ilu.diagpos.upload(diagpos);
  ilu.order.upload(ilu.h_order);
  ilu.val.alloc(colind.size());
  ilu.dinv.alloc(n);
  std::vector<int> rows;
  ilu.stream = false;
  ilu.bsell = false;
  ilu.sdmode = false;
  if (ordering == 3) {
    // numeric factorisation: exact dependency levels of the permuted pattern (interior colours of all parts
    // in parallel, then the separator colours); triangular solves: kernels_sd.cu
    level_schedule(n, rowptr, colind, true, ilu.lvl_ptr_f, rows);
    ilu.lvl_rows_f.upload(rows);
    ilu.lvl_ptr_b.assign(1, 0);
    ilu.colour_ptr = sdo.sep_colour_ptr;
    sd_build(ilu, rowptr, colind, diagpos, sdo.part_ptr, sdo.pcol_ptr, sdo.pcol, sdo.sep_colour_ptr, sdo.level_part_ptr);
    return;
  }
  if (ordering == 2) {
    // numeric factorisation: exact dependency levels of the permuted pattern (about colours x the
    // longest chain inside a block, ~200 launches once per time step); triangular solves: one launch
    // per block colour (bsell_trsv)
    level_schedule(n, rowptr, colind, true, ilu.lvl_ptr_f, rows);
    ilu.lvl_rows_f.upload(rows);
    phase("level schedule");
    ilu.lvl_ptr_b.assign(1, 0);
    ilu.colour_ptr = colour_ptr;
    bsell_build(ilu, rowptr, colind, diagpos, blk_ptr, colour_blk);
    phase("block-SELL storage");
    ilu.bsell = true;
    return;
  }
  if (ordering == 1) {
    // colours are a valid (contiguous) schedule in both directions: a row of colour c only
    // couples with rows of other colours
    rows.resize(n);
    for (int k = 0; k < n; ++k) rows[k] = k;
    ilu.lvl_ptr_f = colour_ptr;
    ilu.lvl_rows_f.upload(rows);
    const int nc = int(colour_ptr.size()) - 1;
    ilu.lvl_ptr_b.assign(nc + 1, 0);
    std::vector<int> rb;
    rb.reserve(n);
    for (int c = nc - 1; c >= 0; --c) {
      for (int k = colour_ptr[c]; k < colour_ptr[c + 1]; ++k) rb.push_back(k);
      ilu.lvl_ptr_b[nc - c] = int(rb.size());
    }
    ilu.lvl_rows_b.upload(rb);
    phase("colour schedule");
    stream_build_ilu(H, ilu, rowptr, colind, diagpos, colour_ptr);
    phase("L / U streams, SELL");
    return;
  }
  level_schedule(n, rowptr, colind, true, ilu.lvl_ptr_f, rows);
  ilu.lvl_rows_f.upload(rows);
  level_schedule(n, rowptr, colind, false, ilu.lvl_ptr_b, rows);
  ilu.lvl_rows_b.upload(rows);
}

template <int BS>
__global__ void k_perm_gather(int n, const int *__restrict__ order, const double *__restrict__ x,
                              double *__restrict__ xp)
{ // xp[k] = x[order[k]]
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n * BS; t += gridDim.x * blockDim.x)
    xp[t] = x[int64_t(BS) * order[t / BS] + (t % BS)];
}
template <int BS>
__global__ void k_perm_scatter(int n, const int *__restrict__ order, const double *__restrict__ yp,
                               double *__restrict__ y)
{ // y[order[k]] = yp[k]
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n * BS; t += gridDim.x * blockDim.x)
    y[int64_t(BS) * order[t / BS] + (t % BS)] = yp[t];
}

__global__ void k_gather(int64_t n, const int *__restrict__ src, const double *__restrict__ a, double *__restrict__ out)
{
  for (int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; k < n; k += int64_t(gridDim.x) * blockDim.x)
    out[k] = a[src[k]];
}

// One warp per row of the current level.  Row i is updated in place in global memory: entries of
// rows of earlier levels are final (L scaled by dinv_j, U scaled by dinv_i as Ifpack stores them).
__global__ void __launch_bounds__(128) k_ilu_factor_level(int n_rows_lvl, const int *__restrict__ rows,
                                                          const int *__restrict__ rowptr,
                                                          const int *__restrict__ colind,
                                                          const int *__restrict__ diagpos, double *__restrict__ val,
                                                          double *__restrict__ dinv)
{
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_rows_lvl) return;
  const int i = rows[w];
  const int rs = rowptr[i], re = rowptr[i + 1], dp = diagpos[i];
  for (int jj = rs; jj < dp; ++jj) {
    const int j = colind[jj];
    const double multiplier = val[jj];
    const int js = diagpos[j] + 1, je = rowptr[j + 1];
    for (int k = js + lane; k < je; k += 32) {
      const int col = colind[k];
      // binary search for col in (jj, re) of row i
      int lo = jj + 1, hi = re - 1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1, c = colind[mid];
        if (c == col) { val[mid] -= multiplier * val[k]; break; }
        if (c < col) lo = mid + 1; else hi = mid - 1;
      }
    }
    __syncwarp();
    if (lane == 0) val[jj] = multiplier * dinv[j];
    __syncwarp();
  }
  double d = val[dp];
  const double MinDiag = 2.2250738585072014e-308, MaxDiag = 1.0 / MinDiag;
  if (fabs(d) > MaxDiag) d = (d < 0) ? -MinDiag : MinDiag; else d = 1.0 / d;
  __syncwarp();
  for (int k = dp + 1 + lane; k < re; k += 32) val[k] *= d;
  if (lane == 0) dinv[i] = d;
}

void ilu_factor(Handle &H, DevIlu &ilu, const double *A_val)
{
  if (ilu.n == 0) return;
  cudaStream_t s = H.stream;
  k_gather<<<unsigned(std::min<int64_t>((ilu.nnz + 255) / 256, kSM * 16)), 256, 0, s>>>(ilu.nnz, ilu.src.p, A_val,
                                                                                       ilu.val.p);
  H.launches++;
  const int nl = int(ilu.lvl_ptr_f.size()) - 1;
  for (int l = 0; l < nl; ++l) {
    const int cnt = ilu.lvl_ptr_f[l + 1] - ilu.lvl_ptr_f[l];
    k_ilu_factor_level<<<(cnt * 32 + 127) / 128, 128, 0, s>>>(cnt, ilu.lvl_rows_f.p + ilu.lvl_ptr_f[l], ilu.rowptr.p,
                                                              ilu.colind.p, ilu.diagpos.p, ilu.val.p, ilu.dinv.p);
    H.launches++;
  }
  NSB_CUDA(cudaGetLastError());
  if (ilu.sdmode) sd_fill(H, ilu);
  else if (ilu.bsell) bsell_fill(H, ilu);
  else if (ilu.sell) {
    sell_fill(H, ilu.sellL, ilu.val.p);
    sell_fill(H, ilu.sellU, ilu.val.p);
  } else if (ilu.stream) stream_split_factors(H, ilu);
}

// Triangular solves, BS right-hand sides interleaved per row (BS = dim for F_s, 1 for S).
template <int BS, int LPR>
__global__ void __launch_bounds__(128) k_trsv_fwd_level(int n_rows_lvl, const int *__restrict__ rows,
                                                        const int *__restrict__ rowptr,
                                                        const int *__restrict__ colind,
                                                        const int *__restrict__ diagpos,
                                                        const double *__restrict__ val, double *__restrict__ y)
{
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = tid / LPR, sl = tid % LPR;
  if (w >= n_rows_lvl) return;
  const int i = rows[w];
  double acc[BS];
#pragma unroll
  for (int d = 0; d < BS; ++d) acc[d] = 0.0;
  const int e = diagpos[i];
  for (int k = rowptr[i] + sl; k < e; k += LPR) {
    const double v = val[k];
    const double *yb = y + int64_t(BS) * colind[k];
#pragma unroll
    for (int d = 0; d < BS; ++d) acc[d] += v * yb[d];
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
    for (int d = 0; d < BS; ++d) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o, LPR);
  if (sl == 0) {
#pragma unroll
    for (int d = 0; d < BS; ++d) y[int64_t(BS) * i + d] -= acc[d];
  }
}

template <int BS, int LPR>
__global__ void __launch_bounds__(128) k_trsv_bwd_level(int n_rows_lvl, const int *__restrict__ rows,
                                                        const int *__restrict__ rowptr,
                                                        const int *__restrict__ colind,
                                                        const int *__restrict__ diagpos,
                                                        const double *__restrict__ val,
                                                        const double *__restrict__ dinv, double *__restrict__ y)
{
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = tid / LPR, sl = tid % LPR;
  if (w >= n_rows_lvl) return;
  const int i = rows[w];
  double acc[BS];
#pragma unroll
  for (int d = 0; d < BS; ++d) acc[d] = 0.0;
  const int e = rowptr[i + 1];
  for (int k = diagpos[i] + 1 + sl; k < e; k += LPR) {
    const double v = val[k];
    const double *yb = y + int64_t(BS) * colind[k];
#pragma unroll
    for (int d = 0; d < BS; ++d) acc[d] += v * yb[d];
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
    for (int d = 0; d < BS; ++d) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o, LPR);
  if (sl == 0) {
    const double di = dinv[i];
#pragma unroll
    for (int d = 0; d < BS; ++d) y[int64_t(BS) * i + d] = y[int64_t(BS) * i + d] * di - acc[d];
  }
}

// 3-RHS variants with the lane = d*8 + e mapping (see spmv_F3_kernel); one warp per row.
__global__ void __launch_bounds__(256) k_trsv_fwd_level3(int n_rows_lvl, const int *__restrict__ rows,
                                                         const int *__restrict__ rowptr,
                                                         const int *__restrict__ colind,
                                                         const int *__restrict__ diagpos,
                                                         const double *__restrict__ val, double *__restrict__ y)
{
  const int lane = threadIdx.x & 31;
  const int d = lane >> 3, e = lane & 7;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_rows_lvl; w += warps_per_grid) {
    const int i = rows[w];
    double acc = 0.0;
    if (d < 3) {
      const int re = diagpos[i];
      int k = rowptr[i] + e;
      for (; k + 8 < re; k += 16) {
        const int c0 = __ldcs(colind + k), c1 = __ldcs(colind + k + 8);
        const double v0 = __ldcs(val + k), v1 = __ldcs(val + k + 8);
        acc += v0 * y[int64_t(3) * c0 + d];
        acc += v1 * y[int64_t(3) * c1 + d];
      }
      if (k < re) acc += __ldcs(val + k) * y[int64_t(3) * __ldcs(colind + k) + d];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (e == 0 && d < 3) y[int64_t(3) * i + d] -= acc;
  }
}

__global__ void __launch_bounds__(256) k_trsv_bwd_level3(int n_rows_lvl, const int *__restrict__ rows,
                                                         const int *__restrict__ rowptr,
                                                         const int *__restrict__ colind,
                                                         const int *__restrict__ diagpos,
                                                         const double *__restrict__ val,
                                                         const double *__restrict__ dinv, double *__restrict__ y)
{
  const int lane = threadIdx.x & 31;
  const int d = lane >> 3, e = lane & 7;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_rows_lvl; w += warps_per_grid) {
    const int i = rows[w];
    double acc = 0.0;
    if (d < 3) {
      const int re = rowptr[i + 1];
      int k = diagpos[i] + 1 + e;
      for (; k + 8 < re; k += 16) {
        const int c0 = __ldcs(colind + k), c1 = __ldcs(colind + k + 8);
        const double v0 = __ldcs(val + k), v1 = __ldcs(val + k + 8);
        acc += v0 * y[int64_t(3) * c0 + d];
        acc += v1 * y[int64_t(3) * c1 + d];
      }
      if (k < re) acc += __ldcs(val + k) * y[int64_t(3) * __ldcs(colind + k) + d];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (e == 0 && d < 3) y[int64_t(3) * i + d] = y[int64_t(3) * i + d] * dinv[i] - acc;
  }
}

template <int BS>
static void trsv_levels(Handle &H, DevIlu &ilu, double *y, cudaStream_t s)
{
  constexpr int LPR = 8;
  const int nf = int(ilu.lvl_ptr_f.size()) - 1;
  const int nb = int(ilu.lvl_ptr_b.size()) - 1;
  if constexpr (BS == 3) {
    for (int l = 1; l < nf; ++l) {
      const int cnt = ilu.lvl_ptr_f[l + 1] - ilu.lvl_ptr_f[l];
      const unsigned g = unsigned(std::min<int64_t>((int64_t(cnt) + 7) / 8, int64_t(kSM) * 64));
      k_trsv_fwd_level3<<<g, 256, 0, s>>>(cnt, ilu.lvl_rows_f.p + ilu.lvl_ptr_f[l], ilu.rowptr.p, ilu.colind.p,
                                         ilu.diagpos.p, ilu.val.p, y);
    }
    for (int l = 0; l < nb; ++l) {
      const int cnt = ilu.lvl_ptr_b[l + 1] - ilu.lvl_ptr_b[l];
      const unsigned g = unsigned(std::min<int64_t>((int64_t(cnt) + 7) / 8, int64_t(kSM) * 64));
      k_trsv_bwd_level3<<<g, 256, 0, s>>>(cnt, ilu.lvl_rows_b.p + ilu.lvl_ptr_b[l], ilu.rowptr.p, ilu.colind.p,
                                         ilu.diagpos.p, ilu.val.p, ilu.dinv.p, y);
    }
    H.launches += (nf > 0 ? nf - 1 : 0) + nb;
    return;
  } else {
  for (int l = 1; l < nf; ++l) { // level 0 rows have an empty L part
    const int cnt = ilu.lvl_ptr_f[l + 1] - ilu.lvl_ptr_f[l];
    k_trsv_fwd_level<BS, LPR><<<(cnt * LPR + 127) / 128, 128, 0, s>>>(cnt, ilu.lvl_rows_f.p + ilu.lvl_ptr_f[l],
                                                                     ilu.rowptr.p, ilu.colind.p, ilu.diagpos.p,
                                                                     ilu.val.p, y);
  }
  for (int l = 0; l < nb; ++l) {
    const int cnt = ilu.lvl_ptr_b[l + 1] - ilu.lvl_ptr_b[l];
    k_trsv_bwd_level<BS, LPR><<<(cnt * LPR + 127) / 128, 128, 0, s>>>(cnt, ilu.lvl_rows_b.p + ilu.lvl_ptr_b[l],
                                                                     ilu.rowptr.p, ilu.colind.p, ilu.diagpos.p,
                                                                     ilu.val.p, ilu.dinv.p, y);
  }
  H.launches += (nf > 0 ? nf - 1 : 0) + nb;
  }
}

// NSB_PDL (default on): chain the colour sweeps of a triangular solve by programmatic dependent launch
// (nsb_internal.hpp).  Read when a solve is captured.  Measured (session M, profiles/README.md): ILU apply of the
// pressure matrix 0.285 -> 0.186 ms at 2 M DoF and 0.51 -> 0.41 ms at 19.9 M; of F_s 0.289 -> 0.258 / 1.98 -> 1.89 ms.
bool pdl_enabled()
{
  const char *e = getenv("NSB_PDL");
  return e ? atoi(e) != 0 : true;
}

// NSB_L2_PERSIST_MB (default 0 = off; measured SLOWER in session M: F_s apply 1.98 -> 2.06 ms with 48 MB, 2.53 ms with
// 79 MB set aside, and the SpMV loses the set-aside too -- kept as a knob for the record): set aside that much of the L2 for persisting lines and mark the staging
// vector of the captured solve as persisting (hit ratio = set-aside / window): a colour sweep gathers rows that
// OTHER colours wrote from the whole vector, which is larger than the L2 at 19.9 M DoF, so without a policy
// every gather is a DRAM miss; with it a fixed subset of the vector stays resident across the sweeps.  The window is
// attached to every kernel node of the captured graph (stream attributes are not inherited by captured nodes).
static void ilu_graph_l2_policy(cudaGraph_t g, void *base, size_t bytes)
{
  const char *e = getenv("NSB_L2_PERSIST_MB");
  const size_t want = e ? size_t(std::max(0, atoi(e))) << 20 : 0;
  if (!want || !bytes) return;
  int dev = 0;
  cudaDeviceProp prop;
  NSB_CUDA(cudaGetDevice(&dev));
  NSB_CUDA(cudaGetDeviceProperties(&prop, dev));
  const size_t set_aside = std::min<size_t>(want, size_t(prop.persistingL2CacheMaxSize));
  if (!set_aside || prop.accessPolicyMaxWindowSize <= 0) return;
  NSB_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside));
  cudaKernelNodeAttrValue v = {};
  v.accessPolicyWindow.base_ptr = base;
  v.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, size_t(prop.accessPolicyMaxWindowSize));
  v.accessPolicyWindow.hitRatio = float(std::min(1.0, double(set_aside) / double(v.accessPolicyWindow.num_bytes)));
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  size_t n = 0;
  NSB_CUDA(cudaGraphGetNodes(g, nullptr, &n));
  std::vector<cudaGraphNode_t> nodes(n);
  NSB_CUDA(cudaGraphGetNodes(g, nodes.data(), &n));
  for (cudaGraphNode_t node : nodes) {
    cudaGraphNodeType t;
    NSB_CUDA(cudaGraphNodeGetType(node, &t));
    if (t == cudaGraphNodeTypeKernel) NSB_CUDA(cudaGraphKernelNodeSetAttribute(node, cudaKernelNodeAttributeAccessPolicyWindow, &v));
  }
  if (getenv("NSB_VERBOSE"))
    fprintf(stderr, "[nsb] L2 policy: %zu MB set aside (max %d MB), window %zu MB (max %d MB), hit ratio %.2f, %zu nodes\n",
            set_aside >> 20, prop.persistingL2CacheMaxSize >> 20, v.accessPolicyWindow.num_bytes >> 20,
            prop.accessPolicyMaxWindowSize >> 20, v.accessPolicyWindow.hitRatio, n);
}

// drop the captured solves (they are re-captured, with the current NSB_PDL / NSB_L2_* settings, on the next use)
void ilu_reset_graphs(Handle &H)
{
  for (DevIlu *ilu : {&H.iluF, &H.iluS}) {
    if (ilu->graph_f) { cudaGraphExecDestroy(ilu->graph_f); ilu->graph_f = nullptr; }
    if (ilu->graph_x) { cudaFree(ilu->graph_x); ilu->graph_x = nullptr; }
  }
  // device-wide L2 knobs follow the environment of the next capture (profiling: scripts/prof_variants.py); a
  // process that never sets them never touches the device-wide limits
  static bool l2_knobs_used = false;
  if (!l2_knobs_used && !getenv("NSB_L2_FETCH") && !getenv("NSB_L2_PERSIST_MB")) return;
  l2_knobs_used = true;
  static size_t fetch_default = 0;
  if (!fetch_default && cudaDeviceGetLimit(&fetch_default, cudaLimitMaxL2FetchGranularity) != cudaSuccess) fetch_default = 0;
  const char *e = getenv("NSB_L2_FETCH"); // hint: DRAM -> L2 fetch granularity in bytes (32 / 64 / 128)
  const size_t gbytes = e && atoi(e) > 0 ? size_t(atoi(e)) : fetch_default;
  if (gbytes) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gbytes);
  if (!getenv("NSB_L2_PERSIST_MB")) {
    cudaCtxResetPersistingL2Cache();
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
  }
  cudaGetLastError(); // the knobs are hints: a refusal is not an error
}

// y = U^{-1} D^{-1} L^{-1} x.  The per-level launches are captured once into a CUDA graph that
// works in place on a fixed staging vector (graph nodes bake their pointers).
void ilu_solve(Handle &H, DevIlu &ilu, const double *x, double *y)
{
  if (ilu.n == 0) return;
  (&ilu == &H.iluF ? H.cnt_ilu_F : H.cnt_ilu_S)++;
  const int nvals = ilu.n * ilu.bs_rhs;
  cudaStream_t s = H.stream;
  if (!ilu.graph_f) {
    if (ilu.sell || ilu.bsell || ilu.sdmode) ilu.io.alloc(2);
    const size_t stage = ilu.sdmode ? size_t(sd_stride(ilu.bs_rhs)) * ilu.n
                         : ilu.sell ? size_t(4) * ilu.n : ilu.bsell ? size_t(bsell_stride(ilu.bs_rhs)) * ilu.n : size_t(nvals);
    NSB_CUDA(cudaMalloc((void **)&ilu.graph_x, sizeof(double) * stage));
    NSB_CUDA(cudaMemsetAsync(ilu.graph_x, 0, sizeof(double) * stage, s));
    cudaGraph_t g;
    const int64_t before = H.launches;
    NSB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    if (ilu.sdmode) sd_trsv(H, ilu, ilu.graph_x, s);
    else if (ilu.bsell) bsell_trsv(H, ilu, ilu.graph_x, s);
    else if (ilu.sell) sell_trsv(H, ilu, ilu.graph_x, s);
    else if (ilu.stream) stream_trsv(H, ilu, ilu.graph_x, s);
    else if (ilu.bs_rhs == 1) trsv_levels<1>(H, ilu, ilu.graph_x, s);
    else if (ilu.bs_rhs == 2) trsv_levels<2>(H, ilu, ilu.graph_x, s);
    else trsv_levels<3>(H, ilu, ilu.graph_x, s);
    NSB_CUDA(cudaStreamEndCapture(s, &g));
    ilu_graph_l2_policy(g, ilu.graph_x, sizeof(double) * stage);
    NSB_CUDA(cudaGraphInstantiate(&ilu.graph_f, g, 0));
    NSB_CUDA(cudaGraphDestroy(g));
    H.launches = before;
  }
  if (ilu.sell || ilu.bsell || ilu.sdmode) { // permutation in / out fused into the first forward and every backward launch
    sell_set_io(H, ilu, x, y);
    NSB_CUDA(cudaGraphLaunch(ilu.graph_f, s));
    H.launches += ilu.sdmode ? int64_t(sd_launches(ilu)) : 2 * (int64_t(ilu.colour_ptr.size()) - 1);
    return;
  }
  const unsigned pg = vgrid(nvals);
  if (ilu.bs_rhs == 1) k_perm_gather<1><<<pg, 256, 0, s>>>(ilu.n, ilu.order.p, x, ilu.graph_x);
  else if (ilu.bs_rhs == 2) k_perm_gather<2><<<pg, 256, 0, s>>>(ilu.n, ilu.order.p, x, ilu.graph_x);
  else k_perm_gather<3><<<pg, 256, 0, s>>>(ilu.n, ilu.order.p, x, ilu.graph_x);
  NSB_CUDA(cudaGraphLaunch(ilu.graph_f, s));
  if (ilu.bs_rhs == 1) k_perm_scatter<1><<<pg, 256, 0, s>>>(ilu.n, ilu.order.p, ilu.graph_x, y);
  else if (ilu.bs_rhs == 2) k_perm_scatter<2><<<pg, 256, 0, s>>>(ilu.n, ilu.order.p, ilu.graph_x, y);
  else k_perm_scatter<3><<<pg, 256, 0, s>>>(ilu.n, ilu.order.p, ilu.graph_x, y);
  H.launches += 2;
  H.launches += (int64_t(ilu.lvl_ptr_f.size()) - 2 > 0 ? int64_t(ilu.lvl_ptr_f.size()) - 2 : 0) +
                (int64_t(ilu.lvl_ptr_b.size()) - 1);
}

// --------------------------------------------------------------------------------------------
// S = B diag(V) Bt on the static symbolic pattern (numeric phase of SparseMatrix::mmult).
// One warp per pressure row; lanes own the Bt entries of the current node, so every S entry is
// accumulated in the reference's order (B row entries ascending, components innermost).
// --------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(128) k_spgemm_schur(int n_rows, const int *__restrict__ b_rowptr,
                                                      const int *__restrict__ b_colind,
                                                      const double *__restrict__ b_val,
                                                      const double *__restrict__ V, int n_nodes_owned, int goff_u,
                                                      const int *__restrict__ bt_rowptr,
                                                      const int *__restrict__ bt_colind,
                                                      const double *__restrict__ bt_val,
                                                      const int *__restrict__ s_rowptr,
                                                      const int *__restrict__ s_colind, double *__restrict__ s_val)
{
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_rows) return;
  const int ss = s_rowptr[w], se = s_rowptr[w + 1];
  for (int k = ss + lane; k < se; k += 32) s_val[k] = 0.0;
  __syncwarp();
  for (int e = b_rowptr[w]; e < b_rowptr[w + 1]; ++e) {
    const int node = b_colind[e];
    // Bt rows of ghost nodes are assembled redundantly (2-layer cell halo) and V = -1/D of ghost
    // nodes arrives by halo exchange, so ghost nodes need no special case here
    double m[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) m[d] = b_val[int64_t(e) * DIM + d] * V[int64_t(DIM) * node + d];
    for (int f = bt_rowptr[node] + lane; f < bt_rowptr[node + 1]; f += 32) {
      const int col = bt_colind[f];
      int lo = ss, hi = se - 1, pos = -1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1, c = s_colind[mid];
        if (c == col) { pos = mid; break; }
        if (c < col) lo = mid + 1; else hi = mid - 1;
      }
      double acc = s_val[pos];
#pragma unroll
      for (int d = 0; d < DIM; ++d) acc += m[d] * bt_val[int64_t(f) * DIM + d];
      s_val[pos] = acc;
    }
    __syncwarp();
  }
  (void)goff_u; (void)n_nodes_owned;
}

void spgemm_schur(Handle &H)
{
  const int n = H.n_p_owned;
  if (n == 0) return;
  const unsigned grid = grid_for(int64_t(n) * 32, 128);
  if (H.dim == 2)
    k_spgemm_schur<2><<<grid, 128, 0, H.stream>>>(n, H.B.rowptr.p, H.B.colind.p, H.B.val.p, H.d_negDinv.p,
                                                  H.n_nodes_owned, H.ghost_off_u(), H.Bt.rowptr.p, H.Bt.colind.p,
                                                  H.Bt.val.p, H.S.rowptr.p, H.S.colind.p, H.S.val.p);
  else
    k_spgemm_schur<3><<<grid, 128, 0, H.stream>>>(n, H.B.rowptr.p, H.B.colind.p, H.B.val.p, H.d_negDinv.p,
                                                  H.n_nodes_owned, H.ghost_off_u(), H.Bt.rowptr.p, H.Bt.colind.p,
                                                  H.Bt.val.p, H.S.rowptr.p, H.S.colind.p, H.S.val.p);
  NSB_CUDA(cudaGetLastError());
  H.launches++;
}

// D, 1/D, -1/D per velocity DoF (Preconditioners.hpp:135-140, 239-245, 350-355, 447-465)
template <int DIM>
__global__ void k_extract_diag(int n_nodes, int ptype, const int *__restrict__ diagpos, const double *__restrict__ F,
                               const double *__restrict__ massdiag, const double *__restrict__ masslump,
                               double *__restrict__ D, double *__restrict__ Dinv, double *__restrict__ negDinv)
{
  GRID_STRIDE(i, n_nodes) {
    double d, nd;
    if (ptype == NSB_PREC_YOSIDA) { d = massdiag[i]; nd = -1.0 / d; }
    else if (ptype == NSB_PREC_AYOSIDA) { d = F[diagpos[i]]; nd = -1.0 / masslump[i]; }
    else { d = F[diagpos[i]]; nd = -1.0 / d; }
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      D[int64_t(DIM) * i + c] = d;
      Dinv[int64_t(DIM) * i + c] = 1.0 / d;
      negDinv[int64_t(DIM) * i + c] = nd;
    }
  }
}

void extract_diag(Handle &H)
{
  const int n = H.n_nodes_owned;
  if (n == 0) return;
  if (H.dim == 2)
    k_extract_diag<2><<<vgrid(n), 256, 0, H.stream>>>(n, H.prm.precond_type, H.d_diagF.p, H.Fs.val.p, H.d_massdiag.p,
                                                      H.d_masslump.p, H.d_D.p, H.d_Dinv.p, H.d_negDinv.p);
  else
    k_extract_diag<3><<<vgrid(n), 256, 0, H.stream>>>(n, H.prm.precond_type, H.d_diagF.p, H.Fs.val.p, H.d_massdiag.p,
                                                      H.d_masslump.p, H.d_D.p, H.d_Dinv.p, H.d_negDinv.p);
  NSB_CUDA(cudaGetLastError());
  H.launches++;
}

// mass-matrix diagonal and |row| sums on the F_s pattern (computed once after the first assembly)
__global__ void k_mass_rows(int n_nodes, const int *__restrict__ rowptr, const int *__restrict__ diagpos,
                            const double *__restrict__ M, double *__restrict__ massdiag, double *__restrict__ masslump)
{
  GRID_STRIDE(i, n_nodes) {
    double s = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) s += fabs(M[k]);
    masslump[i] = s;
    massdiag[i] = M[diagpos[i]];
  }
}

void mass_rows(Handle &H)
{
  const int n = H.n_nodes_owned;
  if (n == 0) return;
  k_mass_rows<<<vgrid(n), 256, 0, H.stream>>>(n, H.Fs.rowptr.p, H.d_diagF.p, H.d_M.p, H.d_massdiag.p, H.d_masslump.p);
  NSB_CUDA(cudaGetLastError());
  H.launches++;
}

static DevBuf<double> g_flush;
void flush_l2(Handle &H)
{
  const size_t n = size_t(256) << 20 >> 3; // 256 MiB > 126 MB L2
  if (g_flush.n != n) g_flush.alloc(n);
  k_zero<<<kSM * 8, 256, 0, H.stream>>>(int(n), g_flush.p);
}

} // namespace nsb
