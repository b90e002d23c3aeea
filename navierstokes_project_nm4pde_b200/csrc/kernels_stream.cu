// kernels_stream.cu -- CSR-stream kernels: SpMV and colour-scheduled triangular solves whose
// global loads do not depend on the row structure.
//
// ncu (profiles/r01_*) showed that sub-warp-per-row CSR kernels on the P2 node graph (about 28
// entries per row, `dim` doubles gathered per entry) are bound by the L1/TEX pipe and by the
// dependent chain rowptr -> colind -> x, not by HBM.  Here a thread block owns a run of whole rows
// holding <= CH entries: phase 1 streams the entries with consecutive threads mapped to
// (entry, component) pairs, so val/colind are read fully coalesced, the `dim` doubles of a
// gathered node are read by adjacent lanes (one or two L1 sectors per entry) and every thread has
// several independent loads in flight; the products land in shared memory.  Phase 2 sums each
// row's products in entry order (the reference's summation order) and writes the row result.
//   MODE 0: y  = A x                (Epetra_CrsMatrix::Multiply)
//   MODE 1: y -= L y                (forward substitution of one colour, unit diagonal)
//   MODE 2: y  = y * dinv - U y     (backward substitution of one colour, Ifpack's scaled U)
#include <algorithm>

#include <cstdlib>

#include <memory>

#include "nsb_internal.hpp"

namespace nsb {

constexpr int kCH = 1024;     // entries per block
constexpr int kMaxRows = 128; // rows per block (phase 2 parallelism; also bounds blocks of empty rows)
constexpr int kSM = 148;
// rows of one colour are length-sorted only inside windows of this many rows, so that a slice
// gathers from a narrow band of the vector (L2 locality) at <2% padding
static int sell_window()
{
  const char *e = getenv("NSB_SELL_WINDOW");
  return e ? std::max(32, atoi(e)) : 4096;
}

template <int BS, int MODE>
__global__ void __launch_bounds__(256) k_stream(const int *__restrict__ blk, const int *__restrict__ rowptr,
                                                const int *__restrict__ colind, const double *__restrict__ val,
                                                const double *x, int n_owned, int goff, double *y,
                                                const double *__restrict__ dinv)
{
  __shared__ double prod[kCH * BS];
  if (MODE != 0) pdl_launch_dependents(); // the next colour's sweep may be scheduled (it blocks in pdl_wait)
  const int r0 = blk[blockIdx.x], r1 = blk[blockIdx.x + 1];
  const int k0 = rowptr[r0], k1 = rowptr[r1];
  const int n = (k1 - k0) * BS;
  int t = threadIdx.x;
  if constexpr (MODE != 0) {
    // triangular sweep: the first four (entry, component) pairs of every thread -- the whole block for one
    // right-hand side -- are loaded before pdl_wait(): they do not depend on the previous sweep and overlap its tail
    int c[4];
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int tt = t + u * 256;
      const bool ok = tt < n;
      const int e = tt / BS;
      c[u] = ok ? __ldcs(colind + k0 + e) : 0;
      v[u] = ok ? __ldcs(val + k0 + e) : 0.0;
    }
    pdl_wait(); // x (= y, the solution being swept) is read only from here on
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int tt = t + u * 256;
      if (tt < n) {
        const int d = tt - (tt / BS) * BS;
        prod[tt] = v[u] * x[int64_t(BS) * c[u] + (c[u] >= n_owned ? goff : 0) + d];
      }
    }
    t += 4 * 256;
  } else {
  for (; t + 3 * 256 < n; t += 4 * 256) {
    int c[4];
    double v[4], xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = (t + u * 256) / BS;
      c[u] = __ldcs(colind + k0 + e);
      v[u] = __ldcs(val + k0 + e);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int tt = t + u * 256, e = tt / BS, d = tt - e * BS;
      xv[u] = x[int64_t(BS) * c[u] + (c[u] >= n_owned ? goff : 0) + d];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) prod[t + u * 256] = v[u] * xv[u];
  }
  }
  for (; t < n; t += 256) {
    const int e = t / BS, d = t - e * BS;
    const int c = __ldcs(colind + k0 + e);
    prod[t] = __ldcs(val + k0 + e) * x[int64_t(BS) * c + (c >= n_owned ? goff : 0) + d];
  }
  __syncthreads();
  const int nr = (r1 - r0) * BS;
  for (int q = threadIdx.x; q < nr; q += 256) {
    const int rr = q / BS, d = q - rr * BS, r = r0 + rr;
    const int a = rowptr[r] - k0, b = rowptr[r + 1] - k0;
    double s = 0.0;
    for (int k = a; k < b; ++k) s += prod[k * BS + d];
    const int64_t o = int64_t(BS) * r + d;
    if (MODE == 0) y[o] = s;
    else if (MODE == 1) y[o] -= s;
    else y[o] = y[o] * dinv[r] - s;
  }
}


// ---------------------------------------------------------------------------------------------
// 3-component row kernels without shared memory.  The L1 data pipe issues one wavefront per cycle
// per SM (ncu: l1tex__data_pipe_lsu_wavefronts ~80% in both the sub-warp-per-row and the
// shared-memory CSR-stream kernels), so the design goal is the fewest wavefronts per entry with
// enough loads in flight: lane = d*8 + e (component d of entry e; lanes 24..31 idle), so one
// gather instruction covers the three doubles of 8 nodes in ~5 wavefronts and val / colind are
// one wavefront per 8 entries; a warp works on RW consecutive rows at once (independent
// accumulators, loads of all rows issued before any FMA); the per-row reduction is three
// shuffles (no L1 traffic).
// ---------------------------------------------------------------------------------------------
template <int MODE, int RW>
__global__ void __launch_bounds__(256) k_rows3(int row_begin, int row_end, const int *__restrict__ rowptr,
                                               const int *__restrict__ colind, const double *__restrict__ val,
                                               const double *x, int n_owned, int goff, double *y,
                                               const double *__restrict__ dinv)
{
  const int lane = threadIdx.x & 31;
  const int d = lane >> 3, e = lane & 7;
  const int dd = d < 3 ? d : 0;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r0 = row_begin + warp * RW; r0 < row_end; r0 += nwarps * RW) {
    int ks[RW], ke[RW];
    double acc[RW];
    int maxlen = 0;
#pragma unroll
    for (int j = 0; j < RW; ++j) {
      const int r = min(r0 + j, row_end - 1);
      ks[j] = rowptr[r];
      ke[j] = (r0 + j < row_end) ? rowptr[r + 1] : ks[j];
      maxlen = max(maxlen, ke[j] - ks[j]);
      acc[j] = 0.0;
    }
    for (int off = e; off < maxlen; off += 8) {
      int c[RW];
      double v[RW], xv[RW];
#pragma unroll
      for (int j = 0; j < RW; ++j) {
        const int k = ks[j] + off;
        const bool ok = k < ke[j];
        c[j] = ok ? __ldcs(colind + k) : -1;
        v[j] = ok ? __ldcs(val + k) : 0.0;
      }
#pragma unroll
      for (int j = 0; j < RW; ++j)
        xv[j] = (c[j] >= 0) ? x[int64_t(3) * c[j] + (c[j] >= n_owned ? goff : 0) + dd] : 0.0;
#pragma unroll
      for (int j = 0; j < RW; ++j) acc[j] += v[j] * xv[j];
    }
#pragma unroll
    for (int j = 0; j < RW; ++j) {
      double a = acc[j];
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      if (e == 0 && d < 3 && r0 + j < row_end) {
        const int64_t o = int64_t(3) * (r0 + j) + d;
        if (MODE == 0) y[o] = a;
        else if (MODE == 1) y[o] -= a;
        else y[o] = y[o] * dinv[r0 + j] - a;
      }
    }
  }
}

static inline unsigned rows3_grid(int nrows, int rw)
{
  const int64_t warps = (int64_t(nrows) + rw - 1) / rw;
  return unsigned(std::max<int64_t>(1, std::min<int64_t>((warps + 7) / 8, int64_t(kSM) * 8)));
}

// rows [r0, r1) -> blocks of whole rows with at most kCH entries; appends the block starts
static void make_row_blocks(const std::vector<int> &rowptr, int r0, int r1, std::vector<int> &blk)
{
  int r = r0;
  while (r < r1) {
    blk.push_back(r);
    const int base = rowptr[r];
    int e = r;
    while (e < r1 && rowptr[e + 1] - base <= kCH && e - r < kMaxRows) ++e;
    if (e == r) throw StateError("CSR-stream: a row has more than 1024 entries");
    r = e;
  }
}

void stream_build_spmv(Handle &H)
{
  std::vector<int> b;
  make_row_blocks(H.hFs.rowptr, 0, H.hFs.n_rows, b);
  b.push_back(H.hFs.n_rows);
  H.n_blk_Fs = int(b.size()) - 1;
  H.d_blk_Fs.upload(b);
  b.clear();
  make_row_blocks(H.hS.rowptr, 0, H.hS.n_rows, b);
  b.push_back(H.hS.n_rows);
  H.n_blk_S = int(b.size()) - 1;
  H.d_blk_S.upload(b);
}

void stream_spmv_F(Handle &H, const double *x_u, int goff_u, double *y_u)
{
  if (H.n_blk_Fs == 0) return;
  if (H.dim == 2)
    k_stream<2, 0><<<H.n_blk_Fs, 256, 0, H.stream>>>(H.d_blk_Fs.p, H.Fs.rowptr.p, H.Fs.colind.p, H.Fs.val.p, x_u,
                                                     H.n_nodes_owned, goff_u, y_u, nullptr);
  else {
    sell_spmv_F(H, x_u, goff_u, y_u);
    return;
  }
  NSB_CUDA(cudaGetLastError());
  H.launches++;
}

void stream_spmv_S(Handle &H, const double *x_p, int goff_p, double *y_p)
{
  if (H.n_blk_S == 0) return;
  k_stream<1, 0><<<H.n_blk_S, 256, 0, H.stream>>>(H.d_blk_S.p, H.S.rowptr.p, H.S.colind.p, H.S.val.p, x_p,
                                                  H.n_p_owned, goff_p, y_p, nullptr);
  NSB_CUDA(cudaGetLastError());
  H.launches++;
}

// strictly lower / strictly upper parts of a factor pattern (diagonal inside) as two CSR structures; map*: position of
// every entry in the combined pattern
void split_lu(const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
              std::vector<int> &Lp, std::unique_ptr<int[]> &Lc, std::unique_ptr<int[]> &mapL, std::vector<int> &Up,
              std::unique_ptr<int[]> &Uc, std::unique_ptr<int[]> &mapU)
{
  const int n = int(rowptr.size()) - 1;
  Lp.assign(n + 1, 0);
  Up.assign(n + 1, 0);
  for (int k = 0; k < n; ++k) {
    Lp[k + 1] = Lp[k] + (diagpos[k] - rowptr[k]);
    Up[k + 1] = Up[k] + (rowptr[k + 1] - diagpos[k] - 1);
  }
  // uninitialised storage, first touched by the filling threads
  const size_t nL = size_t(Lp[n]), nU = size_t(Up[n]);
  Lc.reset(new int[nL + 1]); mapL.reset(new int[nL + 1]); Uc.reset(new int[nU + 1]); mapU.reset(new int[nU + 1]);
#pragma omp parallel for schedule(static)
  for (int k = 0; k < n; ++k) {
    int o = Lp[k];
    for (int e = rowptr[k]; e < diagpos[k]; ++e, ++o) { Lc[o] = colind[e]; mapL[o] = e; }
    o = Up[k];
    for (int e = diagpos[k] + 1; e < rowptr[k + 1]; ++e, ++o) { Uc[o] = colind[e]; mapU[o] = e; }
  }
}

// ---- colour-scheduled triangular solves on split L / U factors --------------------------------
// h_rowptr / h_colind: the combined factor pattern in the permuted index space (diag inside).
void stream_build_ilu(Handle &H, DevIlu &ilu, const std::vector<int> &rowptr, const std::vector<int> &colind,
                      const std::vector<int> &diagpos, const std::vector<int> &colour_ptr)
{
  const int n = ilu.n;
  std::vector<int> Lp, Up;
  std::unique_ptr<int[]> Lc, mapL, Uc, mapU;
  split_lu(rowptr, colind, diagpos, Lp, Lc, mapL, Up, Uc, mapU);
  const size_t nL = size_t(Lp[n]), nU = size_t(Up[n]);
  ilu.colour_ptr = colour_ptr;
  const int nc = int(colour_ptr.size()) - 1;
  std::vector<int> bL, bU;
  ilu.cblkL.assign(nc + 1, 0);
  ilu.cblkU.assign(nc + 1, 0);
  for (int c = 0; c < nc; ++c) {
    ilu.cblkL[c] = int(bL.size());
    make_row_blocks(Lp, colour_ptr[c], colour_ptr[c + 1], bL);
    bL.push_back(colour_ptr[c + 1]); // terminator: blocks of a colour never read past it
    ilu.cblkU[c] = int(bU.size());
    make_row_blocks(Up, colour_ptr[c], colour_ptr[c + 1], bU);
    bU.push_back(colour_ptr[c + 1]);
  }
  ilu.cblkL[nc] = int(bL.size());
  ilu.cblkU[nc] = int(bU.size());
  ilu.Lp.upload(Lp); ilu.Lc.upload(Lc.get(), nL); ilu.mapL.upload(mapL.get(), nL);
  ilu.Up.upload(Up); ilu.Uc.upload(Uc.get(), nU); ilu.mapU.upload(mapU.get(), nU);
  ilu.Lv.alloc(nL); ilu.Uv.alloc(nU);
  ilu.blkL.upload(bL); ilu.blkU.upload(bU);
  ilu.stream = true;
  ilu.sell = false;
  if (ilu.bs_rhs == 3) { // SELL-32 copies of the factors for the 3-component solves
    sell_build(Lp, Lc.get(), mapL.get(), colour_ptr, sell_window(), sell_lanes_for(n), ilu.sellL);
    sell_build(Up, Uc.get(), mapU.get(), colour_ptr, sell_window(), sell_lanes_for(n), ilu.sellU);
    ilu.sell = true;
  }
  (void)H;
}

__global__ void k_gather_map(int64_t n, const int *__restrict__ map, const double *__restrict__ a,
                             double *__restrict__ out)
{
  for (int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; k < n; k += int64_t(gridDim.x) * blockDim.x)
    out[k] = a[map[k]];
}

// after the numeric factorisation: copy the L and U parts into their streaming arrays
void stream_split_factors(Handle &H, DevIlu &ilu)
{
  const int64_t nl = int64_t(ilu.Lv.n), nu = int64_t(ilu.Uv.n);
  if (nl) k_gather_map<<<unsigned(std::min<int64_t>((nl + 255) / 256, kSM * 16)), 256, 0, H.stream>>>(nl, ilu.mapL.p, ilu.val.p, ilu.Lv.p);
  if (nu) k_gather_map<<<unsigned(std::min<int64_t>((nu + 255) / 256, kSM * 16)), 256, 0, H.stream>>>(nu, ilu.mapU.p, ilu.val.p, ilu.Uv.p);
  H.launches += 2;
}

static void rows3_trsv(Handle &H, DevIlu &ilu, double *y, cudaStream_t s)
{
  const int nc = int(ilu.colour_ptr.size()) - 1;
  for (int c = 1; c < nc; ++c) {
    const int a = ilu.colour_ptr[c], b = ilu.colour_ptr[c + 1];
    if (b <= a) continue;
    k_rows3<1, 4><<<rows3_grid(b - a, 4), 256, 0, s>>>(a, b, ilu.Lp.p, ilu.Lc.p, ilu.Lv.p, y, ilu.n, 0, y, nullptr);
    H.launches++;
  }
  for (int c = nc - 1; c >= 0; --c) {
    const int a = ilu.colour_ptr[c], b = ilu.colour_ptr[c + 1];
    if (b <= a) continue;
    k_rows3<2, 4><<<rows3_grid(b - a, 4), 256, 0, s>>>(a, b, ilu.Up.p, ilu.Uc.p, ilu.Uv.p, y, ilu.n, 0, y, ilu.dinv.p);
    H.launches++;
  }
}

template <int BS>
static void stream_trsv_t(Handle &H, DevIlu &ilu, double *y, cudaStream_t s)
{
  const int nc = int(ilu.colour_ptr.size()) - 1;
  const bool pdl = pdl_enabled();
  bool chained = false; // the first sweep of the chain is an ordinary launch
  const double *yc = y;
  const double *none = nullptr;
  for (int c = 1; c < nc; ++c) { // colour 0 has no lower part
    const int nb = ilu.cblkL[c + 1] - ilu.cblkL[c] - 1;
    if (nb <= 0) continue;
    launch_k(k_stream<BS, 1>, nb, 256, 0, s, pdl && chained, ilu.blkL.p + ilu.cblkL[c], ilu.Lp.p, ilu.Lc.p, ilu.Lv.p, yc, ilu.n, 0, y, none);
    chained = true;
    H.launches++;
  }
  for (int c = nc - 1; c >= 0; --c) {
    const int nb = ilu.cblkU[c + 1] - ilu.cblkU[c] - 1;
    if (nb <= 0) continue;
    launch_k(k_stream<BS, 2>, nb, 256, 0, s, pdl && chained, ilu.blkU.p + ilu.cblkU[c], ilu.Up.p, ilu.Uc.p, ilu.Uv.p, yc, ilu.n, 0, y, ilu.dinv.p);
    chained = true;
    H.launches++;
  }
}

void stream_trsv(Handle &H, DevIlu &ilu, double *y, cudaStream_t s)
{
  if (ilu.bs_rhs == 1) stream_trsv_t<1>(H, ilu, y, s);
  else if (ilu.bs_rhs == 2) stream_trsv_t<2>(H, ilu, y, s);
  else rows3_trsv(H, ilu, y, s);
  NSB_CUDA(cudaGetLastError());
}

} // namespace nsb
