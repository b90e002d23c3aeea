// kernels_sell.cu -- sliced-ELL (SELL-32) kernels for the 3-component node graph: SpMV with F_s and
// the colour-scheduled ILU(0) triangular solves.
//
// Why this format (ncu evidence in profiles/r01_*): every CSR variant tried on this matrix
// (sub-warp per row, warp per row, shared-memory CSR-stream) ended either bound by the L1 data
// pipe (l1tex__data_pipe_lsu_wavefronts ~80%: one wavefront per cycle per SM, >1.5 wavefronts
// per stored entry) or by the dependent chain rowptr -> colind -> x with too few rows in flight.
// SELL-32 fixes both:
//   * one thread per row, 32 rows of similar length per slice stored column-major, so colind/val
//     are read as perfectly coalesced 128 B / 256 B wavefronts (0.09 wavefronts per entry) and
//     every lane has `len` independent gathers (2048 rows in flight per SM);
//   * the gathered vector is kept in a padded 4-doubles-per-node staging array, so the three
//     components of a node come with ONE 256-bit load (LDG.E.ENL2.256) instead of three 8-byte
//     accesses;
//   * no reduction at all (a row lives in one thread), entries are summed in column order.
// Rows of one colour are mutually independent, so inside a colour they are sorted by length
// (no padding waste) without changing the ILU(0) factors.
#include <algorithm>
#include <cstdlib>
#include <numeric>

#include <cmath>
#include <memory>

#include "nsb_internal.hpp"

namespace nsb {

constexpr int kSM = 148;

__device__ __forceinline__ void ld256(const double *p, double &a, double &b, double &c)
{
  // the fourth double of the padded node is loaded into a scratch PTX register and dropped
  // (volatile: must not be scheduled above a preceding pdl_wait())
  asm volatile("{ .reg .f64 pad; ld.global.v4.f64 {%0,%1,%2,pad}, [%3]; }" : "=d"(a), "=d"(b), "=d"(c) : "l"(p));
}

// line towards L2 without binding a value (L2 is the point of coherence: safe before pdl_wait())
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// in / out vectors of a captured triangular solve: the graph bakes kernel arguments, so the
// caller's pointers are passed through this device-resident slot (set by k_set_io before launch)
struct TrsvIo { const double *x; double *y; };
__global__ void k_set_io(TrsvIo *io, const double *x, double *y) { io->x = x; io->y = y; }

// MODE 0: y[3r+d]  = sum                          (SpMV, unpadded output)
// MODE 1: y[4r+d]  = x[3 order[r]+d] - sum        (forward substitution of one colour; y is the padded
//                                                  staging in factor order, x the caller's vector:
//                                                  the permutation is fused into the sweep)
// MODE 2: y[4r+d]  = y*dinv - sum, also stored to out[3 order[r]+d]   (backward substitution)
// U: depth of the (col, val) software pipeline.  PAD (MODE 0 only): gather from the padded copy of x
// (one 256-bit load per entry) or straight from the caller's vector (three 64-bit loads, no copy pass).
// LPR: lanes per row.  1 = one thread per row (32 rows per slice; throughput layout for large
// matrices).  4 = four adjacent lanes share a row (8 rows per slice, entry e of a row sits in lane
// e % 4 of group e / 4) and combine their partial sums with two shuffles: a quarter of the dependent
// load chain per row and four times the warps, for colour sweeps too small to fill the machine.
template <int MODE, int U = 4, bool PAD = true, int LPR = 1>
__global__ void __launch_bounds__(256, U == 4 ? (MODE == 0 ? 6 : 5) : 3) k_sell3(int s0, int s1, const int *__restrict__ slice_ptr,
                                               const int *__restrict__ rowid, const int *__restrict__ col,
                                               const double *__restrict__ val, const double *xp, double *y,
                                               const double *__restrict__ dinv, const int *__restrict__ order = nullptr,
                                               const TrsvIo *__restrict__ io = nullptr)
{
  // one warp per slice, no grid-stride loop: slices are length-sorted inside windows, so a fixed
  // stride would hand the same warps the long slices of every window (ncu: 52% achieved occupancy)
  const int lane = threadIdx.x & 31;
  const int s = s0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (MODE != 0) pdl_launch_dependents(); // the next colour's sweep may be scheduled (it blocks in pdl_wait)
  if (s < s1) {
    const int base = slice_ptr[s];
    const int len = (slice_ptr[s + 1] - base) >> 5;
    const int r = rowid[(int64_t(s) << 5) + lane];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    // row-local operands of the triangular sweeps are requested before the entry loop so that their
    // latency (order -> x) overlaps it.  Everything up to pdl_wait() is independent of the previous sweep
    // (the caller's x is not written inside the chain) and overlaps its tail.
    double b0 = 0.0, b1 = 0.0, b2 = 0.0, di = 1.0;
    int ro = 0;
    if (MODE != 0 && r >= 0) {
      ro = order[r];
      if (MODE == 1) {
        const double *xi = io->x + 3 * int64_t(ro);
        b0 = xi[0]; b1 = xi[1]; b2 = xi[2];
      } else
        di = dinv[r];
    }
    const int *cp = col + base + lane;
    const double *vp = val + base + lane;
    // software pipeline: the colind/val loads of the next group of U entries are issued before the
    // gathers of the current group, so one memory round trip covers U entries
    int c[U], nc[U];
    double v[U], nv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = u < len;
      c[u] = ok ? __ldcs(cp + u * 32) : 0;
      v[u] = ok ? __ldcs(vp + u * 32) : 0.0;
    }
    if (MODE != 0) pdl_wait(); // the staging vector y is read (gathers, own row) only from here on
    for (int k = 0; k < len; k += U) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = k + U + u < len;
        nc[u] = ok ? __ldcs(cp + (k + U + u) * 32) : 0;
        nv[u] = ok ? __ldcs(vp + (k + U + u) * 32) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        double x0, x1, x2;
        if (PAD) ld256(xp + 4 * int64_t(c[u]), x0, x1, x2);
        else { const double *xb = xp + 3 * int64_t(c[u]); x0 = xb[0]; x1 = xb[1]; x2 = xb[2]; }
        a0 += v[u] * x0;
        a1 += v[u] * x1;
        a2 += v[u] * x2;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { c[u] = nc[u]; v[u] = nv[u]; }
    }
    if (LPR == 4) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, 1); a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
      a0 += __shfl_xor_sync(0xffffffffu, a0, 2); a1 += __shfl_xor_sync(0xffffffffu, a1, 2); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
    }
    if (r >= 0 && (LPR == 1 || (lane & 3) == 0)) {
      if (MODE == 0) {
        double *o = y + 3 * int64_t(r);
        o[0] = a0; o[1] = a1; o[2] = a2;
      } else {
        double *o = y + 4 * int64_t(r);
        if (MODE == 1) {
          o[0] = b0 - a0; o[1] = b1 - a1; o[2] = b2 - a2; o[3] = 0.0;
        } else {
          const double r0 = o[0] * di - a0, r1 = o[1] * di - a1, r2 = o[2] * di - a2;
          o[0] = r0; o[1] = r1; o[2] = r2;
          double *yo = io->y + 3 * int64_t(ro);
          yo[0] = r0; yo[1] = r1; yo[2] = r2;
        }
      }
    }
  }
}

__global__ void k_sell_fill(int64_t n, const int *__restrict__ map, const double *__restrict__ src,
                            double *__restrict__ val)
{
  for (int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; k < n; k += int64_t(gridDim.x) * blockDim.x) {
    const int m = map[k];
    val[k] = m >= 0 ? src[m] : 0.0;
  }
}

// x (stride 3, ghosts at +goff) -> padded copy (stride 4)
__global__ void k_pad3(int n_nodes, int n_owned, int goff, const double *__restrict__ x, double *__restrict__ xp)
{
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_nodes * 4; t += gridDim.x * blockDim.x) {
    const int i = t >> 2, d = t & 3;
    xp[t] = d < 3 ? x[int64_t(3) * i + (i >= n_owned ? goff : 0) + d] : 0.0;
  }
}
// Build SELL-32 for the rows of `rowptr/colind`.  ranges: row ranges that must not share a slice
// (colours); inside a range rows are sorted by length (descending) within windows of `window`
// rows.  src[e]: index of CSR entry e in the value array the SELL values are filled from.
// lanes: 1 or 4 lanes per row (see k_sell3).
void sell_build(const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &src,
                const std::vector<int> &ranges, int window, int lanes, DevSell &out)
{
  sell_build(rowptr, colind.data(), src.empty() ? nullptr : src.data(), ranges, window, lanes, out);
}

// colind / src as plain arrays (src may be null: the entries are numbered as in rowptr / colind)
void sell_build(const std::vector<int> &rowptr, const int *colind, const int *src, const std::vector<int> &ranges, int window,
                int lanes, DevSell &out)
{
  sell_build_keep(rowptr, colind, src, ranges, window, lanes, out, nullptr);
}

// keep (test only, sell_host_check): host copies of the packed arrays
void sell_build_keep(const std::vector<int> &rowptr, const int *colind, const int *src, const std::vector<int> &ranges,
                     int window, int lanes, DevSell &out, SellHost *keep)
{
  out.range_slice.assign(ranges.size(), 0);
  out.lanes = lanes;
  const int R = 32 / lanes; // rows per slice
  // 1. rows of every range, length-sorted inside windows (the windows are independent: all host threads)
  size_t n_rows = 0;
  for (size_t g = 0; g + 1 < ranges.size(); ++g) n_rows += size_t(ranges[g + 1] - ranges[g]);
  std::vector<int> rows(n_rows);
  struct Slice { size_t first; int ns; }; // position of the slice's first row in `rows`, rows in the slice
  std::vector<Slice> slices;
  std::vector<std::pair<size_t, size_t>> windows;
  size_t base_r = 0;
  for (size_t g = 0; g + 1 < ranges.size(); ++g) {
    out.range_slice[g] = int(slices.size());
    const int a = ranges[g], b = ranges[g + 1];
    std::iota(rows.begin() + base_r, rows.begin() + base_r + (b - a), a);
    for (int w0 = 0; w0 < b - a; w0 += window) windows.emplace_back(base_r + w0, base_r + std::min(b - a, w0 + window));
    for (int s = 0; s < b - a; s += R) slices.push_back(Slice{base_r + s, std::min(R, b - a - s)});
    base_r += size_t(b - a);
  }
  const int64_t nw = int64_t(windows.size());
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t w = 0; w < nw; ++w)
    std::stable_sort(rows.begin() + windows[w].first, rows.begin() + windows[w].second, [&](int x, int y) {
      return rowptr[x + 1] - rowptr[x] > rowptr[y + 1] - rowptr[y];
    });
  // 2. slice lengths (groups of 32 slots), prefix sum
  const int64_t nsl = int64_t(slices.size());
  std::vector<int> slice_ptr(size_t(nsl) + 1, 0);
  std::vector<int64_t> off(size_t(nsl) + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nsl; ++i) {
    int len = 0;
    for (int l = 0; l < slices[i].ns; ++l) {
      const int r = rows[slices[i].first + l];
      len = std::max(len, (rowptr[r + 1] - rowptr[r] + lanes - 1) / lanes);
    }
    off[i + 1] = int64_t(len) * 32;
  }
  for (int64_t i = 0; i < nsl; ++i) off[i + 1] += off[i];
  if (off[nsl] > int64_t(0x7fffffff)) throw StateError("SELL: more than 2^31 slots");
  for (int64_t i = 0; i <= nsl; ++i) slice_ptr[i] = int(off[i]);
  // 3. fill the slices
  // (uninitialised storage: every element is written below, by the thread that touches its pages first --
  // zero-filling 1.5 GB of std::vector on one thread cost more than the fill itself)
  const size_t n_slots = size_t(off[nsl]);
  std::unique_ptr<int[]> rowid(new int[size_t(nsl) * 32 + 1]), col(new int[n_slots + 1]), map(new int[n_slots + 1]);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nsl; ++i) {
    const int ns = slices[i].ns;
    const int *rs = &rows[slices[i].first];
    const size_t base = size_t(off[i]);
    const int len = int((off[i + 1] - off[i]) / 32);
    int rl_of[32], pad_of[32], rp_of[32];
    for (int l = 0; l < 32; ++l) {
      const int r = l / lanes < ns ? rs[l / lanes] : -1;
      rowid[size_t(i) * 32 + l] = r;
      rp_of[l] = r >= 0 ? rowptr[r] : 0;
      rl_of[l] = r >= 0 ? rowptr[r + 1] - rowptr[r] : 0;
      pad_of[l] = (r >= 0 && rl_of[l] > 0) ? colind[rowptr[r]] : 0;
    }
    for (int k = 0; k < len; ++k) // slot order = memory order
      for (int l = 0; l < 32; ++l) {
        const size_t o = base + size_t(k) * 32 + l;
        const int e = k * lanes + l % lanes;
        if (e < rl_of[l]) { col[o] = colind[rp_of[l] + e]; map[o] = src ? src[rp_of[l] + e] : rp_of[l] + e; }
        else { col[o] = pad_of[l]; map[o] = -1; }
      }
  }
  out.range_slice.back() = int(nsl);
  out.n_slices = int(nsl);
  out.n_slots = int64_t(n_slots);
  out.slice_ptr.upload(slice_ptr);
  out.rowid.upload(rowid.get(), size_t(nsl) * 32);
  out.col.upload(col.get(), n_slots);
  out.map.upload(map.get(), n_slots);
  out.val.alloc(n_slots);
  if (keep) {
    keep->slice_ptr = slice_ptr;
    keep->rowid.assign(rowid.get(), rowid.get() + size_t(nsl) * 32);
    keep->col.assign(col.get(), col.get() + n_slots);
    keep->map.assign(map.get(), map.get() + n_slots);
  }
}

// ---- CPU emulation of k_sell3 on the packed storage (nsb_debug_sell_check, tests/test_host_cpu.py) ----------------
// SELL-32 copies of the split L / U factors with the colours as row ranges, and of the whole matrix for the SpMV,
// built exactly as for the device (uploads skipped: setup dry run), filled with a pseudo-random factor and walked the
// way the kernel does: one warp per slice, lane l owns row rowid[32 s + l] (LPR lanes per row, partial sums added),
// `len` steps of 32 slots; colours forward for L, backward for U, the slices of a colour in REVERSE order.  Returns
// the largest difference to plain substitution / CSR SpMV relative to the largest entry (>= 1e30: structural
// violation).  stats[3]: colours, slices of L, padded slots of L per stored entry in per mille.
double sell_host_check(const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
                       const std::vector<int> &colour_ptr, int bs, int lanes, int window, int *stats)
{
  const int n = int(rowptr.size()) - 1, ncol = int(colour_ptr.size()) - 1;
  std::vector<int> Lp, Up;
  std::unique_ptr<int[]> Lc, mapL, Uc, mapU;
  split_lu(rowptr, colind, diagpos, Lp, Lc, mapL, Up, Uc, mapU);
  DevSell dL, dU, dA;
  SellHost hL, hU, hA;
  const bool was_on = g_dry.on;
  g_dry.on = true;
  try {
    sell_build_keep(Lp, Lc.get(), mapL.get(), colour_ptr, window, lanes, dL, &hL);
    sell_build_keep(Up, Uc.get(), mapU.get(), colour_ptr, window, lanes, dU, &hU);
    sell_build_keep(rowptr, colind.data(), nullptr, {0, n}, 2048, lanes, dA, &hA);
  } catch (...) { g_dry.on = was_on; throw; }
  g_dry.on = was_on;
  if (!was_on) g_dry.log.clear();
  if (stats) { stats[0] = ncol; stats[1] = dL.n_slices; stats[2] = Lp[n] ? int(1000.0 * double(dL.n_slots) / Lp[n]) : 0; }
  auto rnd = [](uint64_t k) { k = (k + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull; k ^= k >> 31; k *= 0x94D049BB133111EBull; k ^= k >> 29;
                              return double(k >> 11) / double(1ull << 53); };
  std::vector<double> val(colind.size()), dinv(n), x(size_t(n) * bs);
  for (int k = 0; k < n; ++k) {
    const int len = std::max(1, rowptr[k + 1] - rowptr[k]);
    for (int e = rowptr[k]; e < rowptr[k + 1]; ++e) val[e] = (rnd(uint64_t(e)) - 0.5) / len;
    dinv[k] = 0.5 + rnd(uint64_t(k) + (1ull << 40));
    for (int d = 0; d < bs; ++d) x[size_t(k) * bs + d] = rnd(uint64_t(k) * 3 + d + (1ull << 41)) - 0.5;
  }
  std::vector<double> ref(x), ref_mv(size_t(n) * bs, 0.0);
  for (int k = 0; k < n; ++k)
    for (int e = rowptr[k]; e < rowptr[k + 1]; ++e)
      for (int d = 0; d < bs; ++d) ref_mv[size_t(k) * bs + d] += val[e] * x[size_t(colind[e]) * bs + d];
  for (int k = 0; k < n; ++k)
    for (int e = rowptr[k]; e < diagpos[k]; ++e)
      for (int d = 0; d < bs; ++d) ref[size_t(k) * bs + d] -= val[e] * ref[size_t(colind[e]) * bs + d];
  for (int k = n - 1; k >= 0; --k) {
    for (int d = 0; d < bs; ++d) ref[size_t(k) * bs + d] *= dinv[k];
    for (int e = diagpos[k] + 1; e < rowptr[k + 1]; ++e)
      for (int d = 0; d < bs; ++d) ref[size_t(k) * bs + d] -= val[e] * ref[size_t(colind[e]) * bs + d];
  }
  // one slice as the kernel walks it: sums[l] = partial sum of lane l, combined over the `lanes` lanes of a row
  auto slice_sums = [&](const SellHost &S, int s, const std::vector<double> &v, const std::vector<double> &src_vec,
                        double (&sum)[32][3], std::vector<char> &seen) -> double {
    const int base = S.slice_ptr[s], len = (S.slice_ptr[s + 1] - base) >> 5;
    if (base % 32 || (S.slice_ptr[s + 1] - base) % 32) return 1e30;
    for (int l = 0; l < 32; ++l) {
      const int r = S.rowid[size_t(s) * 32 + l];
      if (r >= n || (l % lanes && r != S.rowid[size_t(s) * 32 + l - 1])) return 2e30; // the lanes of a row are adjacent
      if (r >= 0 && l % lanes == 0) { if (seen[r]) return 3e30; seen[r] = 1; }         // every row in exactly one slice
      for (int d = 0; d < bs; ++d) sum[l][d] = 0.0;
      for (int k = 0; k < len; ++k) {
        const size_t o = size_t(base) + size_t(k) * 32 + l;
        const double a = S.map[o] >= 0 ? v[S.map[o]] : 0.0; // k_sell_fill
        if (S.col[o] < 0 || S.col[o] >= n) return 4e30;      // padding names a valid row (value 0)
        for (int d = 0; d < bs; ++d) sum[l][d] += a * src_vec[size_t(S.col[o]) * bs + d];
      }
    }
    for (int l = 0; l < 32; l += lanes)
      for (int j = 1; j < lanes; ++j)
        for (int d = 0; d < bs; ++d) sum[l][d] += sum[l + j][d];
    return 0.0;
  };
  double sum[32][3];
  std::vector<char> seen(n, 0);
  // SpMV
  std::vector<double> y_mv(size_t(n) * bs, 0.0);
  for (int s = 0; s < dA.n_slices; ++s) {
    const double bad = slice_sums(hA, s, val, x, sum, seen);
    if (bad > 0) return bad;
    for (int l = 0; l < 32; l += lanes) {
      const int r = hA.rowid[size_t(s) * 32 + l];
      if (r >= 0) for (int d = 0; d < bs; ++d) y_mv[size_t(r) * bs + d] = sum[l][d];
    }
  }
  for (int k = 0; k < n; ++k) if (!seen[k]) return 5e30;
  // triangular sweeps
  std::vector<double> y(x);
  for (int dir = 0; dir < 2; ++dir) {
    const SellHost &S = dir == 0 ? hL : hU;
    const DevSell &D = dir == 0 ? dL : dU;
    std::fill(seen.begin(), seen.end(), 0);
    for (int cc = 0; cc < ncol; ++cc) {
      const int c = dir == 0 ? cc : ncol - 1 - cc;
      for (int s = D.range_slice[c + 1] - 1; s >= D.range_slice[c]; --s) {
        const double bad = slice_sums(S, s, val, y, sum, seen);
        if (bad > 0) return bad + 10e30;
        for (int l = 0; l < 32; l += lanes) {
          const int r = S.rowid[size_t(s) * 32 + l];
          if (r < 0) continue;
          if (r < colour_ptr[c] || r >= colour_ptr[c + 1]) return 6e30; // a slice never mixes colours
          for (int d = 0; d < bs; ++d)
            y[size_t(r) * bs + d] = dir == 0 ? y[size_t(r) * bs + d] - sum[l][d] : y[size_t(r) * bs + d] * dinv[r] - sum[l][d];
        }
      }
    }
    for (int k = 0; k < n; ++k) if (!seen[k]) return 7e30;
  }
  double err = 0, scale = 0;
  for (size_t k = 0; k < y.size(); ++k) { err = std::max(err, std::fabs(y[k] - ref[k])); scale = std::max(scale, std::fabs(ref[k])); }
  for (size_t k = 0; k < y.size(); ++k) { err = std::max(err, std::fabs(y_mv[k] - ref_mv[k])); scale = std::max(scale, std::fabs(ref_mv[k])); }
  return err / std::max(scale, 1e-300);
}

void sell_fill(Handle &H, DevSell &S, const double *src)
{
  if (S.n_slots == 0) return;
  k_sell_fill<<<unsigned(std::min<int64_t>((S.n_slots + 255) / 256, kSM * 16)), 256, 0, H.stream>>>(S.n_slots, S.map.p,
                                                                                                 src, S.val.p);
  H.launches++;
}

static inline unsigned sell_grid(int n_slices) { return unsigned(std::max(1, (n_slices + 7) / 8)); }

static int env_int(const char *name, int def)
{
  const char *e = getenv(name);
  return e ? atoi(e) : def;
}

// lanes per row for a matrix with n_rows rows: the 4-lane layout below ~1.5 M rows, where a colour
// sweep (or the whole SpMV) cannot fill 148 SMs with one row per thread.  NSB_SELL_LANES overrides.
int sell_lanes_for(int n_rows)
{
  const int e = env_int("NSB_SELL_LANES", 0);
  if (e == 1 || e == 4) return e;
  return n_rows < 1500000 ? 4 : 1;
}

template <int MODE, bool PAD>
static void launch_sell(cudaStream_t s, const DevSell &S, int a, int b, int depth, const double *xp, double *y,
                        const double *dinv, const int *order, const TrsvIo *io, bool pdl = false)
{
  const unsigned g = sell_grid(b - a);
  const int *sp = S.slice_ptr.p, *ri = S.rowid.p, *cl = S.col.p;
  const double *vl = S.val.p;
  if (S.lanes == 4) {
    if (depth == 8) launch_k(k_sell3<MODE, 8, PAD, 4>, g, 256, 0, s, pdl, a, b, sp, ri, cl, vl, xp, y, dinv, order, io);
    else launch_k(k_sell3<MODE, 4, PAD, 4>, g, 256, 0, s, pdl, a, b, sp, ri, cl, vl, xp, y, dinv, order, io);
  } else {
    if (depth == 8) launch_k(k_sell3<MODE, 8, PAD, 1>, g, 256, 0, s, pdl, a, b, sp, ri, cl, vl, xp, y, dinv, order, io);
    else launch_k(k_sell3<MODE, 4, PAD, 1>, g, 256, 0, s, pdl, a, b, sp, ri, cl, vl, xp, y, dinv, order, io);
  }
}

// y_u = F_s x_u (3D).  Single rank (no ghost offset): the three components are gathered straight from
// x (measured at 19.9 M DoF: 0.68 ms against 0.75 ms with the padded copy pass + 256-bit gathers).
// With ghosts (multi-rank) x is first copied into the padded staging vector (k_pad3), which also
// folds the ghost offset.  NSB_SPMV_PAD=1 forces the padded path, NSB_SELL_U the pipeline depth
// (4 | 8; 8 measured 5-7% faster on both the SpMV and the ILU sweeps with one lane per row).
void sell_spmv_F(Handle &H, const double *x_u, int goff_u, double *y_u)
{
  if (H.sellF_dirty) {
    sell_fill(H, H.sellF, H.Fs.val.p);
    H.sellF_dirty = false;
  }
  static const int depth_env = env_int("NSB_SELL_U", 0), pad = env_int("NSB_SPMV_PAD", 0);
  const DevSell &S = H.sellF;
  const int depth = depth_env ? depth_env : (S.lanes == 4 ? 4 : 8);
  const int nn = H.n_nodes;
  if (!pad && H.n_nodes == H.n_nodes_owned) {
    launch_sell<0, false>(H.stream, S, 0, S.n_slices, depth, x_u, y_u, nullptr, nullptr, nullptr);
    NSB_CUDA(cudaGetLastError());
    H.launches += 1;
    return;
  }
  k_pad3<<<unsigned(std::min((nn * 4 + 255) / 256, kSM * 16)), 256, 0, H.stream>>>(nn, H.n_nodes_owned, goff_u, x_u,
                                                                                  H.d_xpad.p);
  launch_sell<0, true>(H.stream, S, 0, S.n_slices, depth, H.d_xpad.p, y_u, nullptr, nullptr, nullptr);
  NSB_CUDA(cudaGetLastError());
  H.launches += 2;
}

// yp: padded staging (factor order).  The caller's in / out vectors come through ilu.io (sell_set_io).
// Colour 0 has no lower part (zero-length slices): its forward launch is the fused permutation.
void sell_trsv(Handle &H, DevIlu &ilu, double *yp, cudaStream_t s)
{
  const int nc = int(ilu.colour_ptr.size()) - 1;
  const TrsvIo *io = reinterpret_cast<const TrsvIo *>(ilu.io.p);
  static const int depth_env = env_int("NSB_SELL_U", 0);
  const int depth = depth_env ? depth_env : (ilu.sellL.lanes == 4 ? 4 : 8);
  const bool pdl = pdl_enabled();
  bool chained = false; // the first sweep of the chain is an ordinary launch
  for (int c = 0; c < nc; ++c) {
    const int a = ilu.sellL.range_slice[c], b = ilu.sellL.range_slice[c + 1];
    if (b <= a) continue;
    launch_sell<1, true>(s, ilu.sellL, a, b, depth, yp, yp, nullptr, ilu.order.p, io, pdl && chained);
    chained = true;
    H.launches++;
  }
  for (int c = nc - 1; c >= 0; --c) {
    const int a = ilu.sellU.range_slice[c], b = ilu.sellU.range_slice[c + 1];
    if (b <= a) continue;
    launch_sell<2, true>(s, ilu.sellU, a, b, depth, yp, yp, ilu.dinv.p, ilu.order.p, io, pdl && chained);
    chained = true;
    H.launches++;
  }
  NSB_CUDA(cudaGetLastError());
}


// ------------------------------------------------------------------------------------------------
// Block multicolour triangular solves (ilu_ordering = 2; the ordering is built in kernels_linalg.cu:
// block_multicolour_order).  One warp solves one block of <= 32 consecutive factor rows, one launch
// sweeps all blocks of one block colour.  Per block:
//   1. the entries coupling with OTHER blocks (final when this colour is swept): the block's rows are
//      sorted by their number of such entries and handled in four passes of eight rows, four adjacent
//      lanes per row (entry e of a row sits in lane e % 4 of step e / 4): perfectly coalesced (col, val)
//      streams, one 256-bit gather per entry for the 3 components, no per-entry row tag and no
//      segmented scan -- the partial sums of a row are combined by two shuffles (fixed order:
//      deterministic).  (The first version packed the entries densely and reduced them with a
//      five-round segmented warp scan: ncu, profiles/r02: 2 440 instructions per block, 96 registers,
//      31 % occupancy, instruction- and latency-bound at 1.9 TB/s.)
//   2. the entries INSIDE the block are eliminated sequentially: lane l owns row l; at step r the
//      finished value of row r is broadcast by shuffle and the lanes whose next entry sits in column r
//      consume it (entries staged in shared memory, sorted by column, so a lane only advances a cursor).
// The permutation into factor order is fused into the forward sweep (reads the caller's x through
// `order`) and the inverse permutation into the backward sweep, as in k_sell3.
// ------------------------------------------------------------------------------------------------
constexpr int kBW = 4; // warps (= blocks of rows) per CTA
// outside rows of a block staged in shared memory; the (few) others are gathered from global.  NSB_BSELL_XCAP overrides
// (0: no staging at all).
static int bsell_xcap()
{
  const char *e = getenv("NSB_BSELL_XCAP");
  return e ? std::max(0, atoi(e)) : 1024;
}

// NSB_BSELL_PREFETCH (default on; read when a solve is captured): software prefetch of a block's gathers into L2.
// Session O: S-matrix apply 0.815 -> 0.672 ms at 19.9 M DoF and 0.686 -> 0.550 ms at 2 M; F_s apply 1.883 -> 1.854 ms
// and 0.458 -> 0.388 ms.
static int bsell_prefetch()
{
  const char *e = getenv("NSB_BSELL_PREFETCH");
  return e ? atoi(e) : 1;
}

// NSB_BSELL_PIPE (read when a solve is captured): the software-pipelined walk over the four passes (k_bsell<.., PIPE>).
// Default: on for one right-hand side (the pressure matrix: 0.673 -> 0.509 ms per apply at 19.9 M DoF, 0.551 -> 0.382 ms
// at 2 M, session Q), off for the 3-component velocity block, where it measured slower both at 64 registers with
// spills (1.867 -> 2.152 ms) and at 80 registers / 6 CTAs per SM without (1.855 -> 2.082 ms, session T).
static int bsell_pipe(int bs)
{
  const char *e = getenv("NSB_BSELL_PIPE");
  return e ? atoi(e) : (bs == 1 ? 1 : 0);
}

int bsell_stride(int bs_rhs) { return bs_rhs == 3 ? 4 : bs_rhs; }

static size_t bsell_warp_bytes(int bs, int max_int, int max_nx)
{ // acc[32*bs] doubles | xs[max_nx*bs] doubles | ival[max_int] doubles | ioff[33] ushort (72 B) | icol[max_int] bytes
  const size_t b = size_t(32) * bs * 8 + size_t(max_nx) * bs * 8 + size_t(max_int) * 8 + 72 + size_t(max_int);
  return (b + 15) & ~size_t(15);
}

template <int BS>
__device__ __forceinline__ void bsell_gather(const double *yp, int c, double (&x)[BS])
{
  if constexpr (BS == 3) ld256(yp + 4 * int64_t(c), x[0], x[1], x[2]);
  else if constexpr (BS == 2) {
    double a, b;
    asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(yp + 2 * int64_t(c)));
    x[0] = a; x[1] = b;
  } else {
    double a;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(a) : "l"(yp + c));
    x[0] = a;
  }
}

// DIR 0: forward substitution  y = x - L y          (unit diagonal, Ifpack: L scaled by dinv_j)
// DIR 1: backward substitution y = y * dinv - U y   (Ifpack: U scaled by dinv_i), also stored to io->y
// (8 CTAs per SM = 64 registers; a 48-register build for 10 CTAs per SM was measured slower: 2.13 ms against 1.98 ms
// per apply at 19.9 M DoF, session M)
// PIPE: the entries coupling with other blocks are walked as ONE software-pipelined stream over the four passes
// (the (col, val) loads of the next four steps are in flight while the current four gathers are, and the first four
// are issued before pdl_wait()) instead of pass by pass: ncu (session P, profiles/r02) shows the pass-by-pass version
// waiting on a dependent (col, val) -> gather round trip per pass -- ~12 dependent round trips and 22 us per block,
// long-scoreboard stalls 12.8 per issued instruction at 43 % occupancy.
template <int BS>
__device__ __forceinline__ void bsell_flush(double (&a)[BS], double *acc, int prow, int q, int lane)
{
  constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
  for (int o = 1; o < 4; o <<= 1)
#pragma unroll
    for (int d = 0; d < BS; ++d) a[d] += __shfl_xor_sync(FULL, a[d], o);
  const int lr = __shfl_sync(FULL, prow, q * 8 + (lane >> 2)); // local row of (pass q, slot lane / 4)
  if ((lane & 3) == 0) { // every local row sits in exactly one (pass, slot): acc needs no clearing
#pragma unroll
    for (int d = 0; d < BS; ++d) acc[lr * BS + d] = a[d];
  }
#pragma unroll
  for (int d = 0; d < BS; ++d) a[d] = 0.0;
}

template <int BS, int DIR, bool STAGE, bool PIPE = false>
__global__ void __launch_bounds__(kBW * 32, 8) k_bsell(int b0, int b1, const int *__restrict__ blk_row,
                                                       const int *__restrict__ e_ptr, const unsigned *__restrict__ e_len,
                                                       const unsigned char *__restrict__ e_prow,
                                                       const unsigned short *__restrict__ e_lix,
                                                       const int *__restrict__ e_col,
                                                       const double *__restrict__ e_val, const int *__restrict__ x_ptr,
                                                       const int *__restrict__ x_ids, const int *__restrict__ i_ptr,
                                                       const unsigned short *__restrict__ i_off,
                                                       const unsigned char *__restrict__ i_col,
                                                       const double *__restrict__ i_val,
                                                       const unsigned *__restrict__ i_mask, double *yp,
                                                       const double *__restrict__ dinv, const int *__restrict__ order,
                                                       const TrsvIo *__restrict__ io, int max_int, int max_nx, int warp_bytes,
                                                       int prefetch)
{
  constexpr int PS = BS == 3 ? 4 : BS;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char bsell_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = b0 + blockIdx.x * kBW + warp;
  pdl_launch_dependents(); // the next colour's sweep may be scheduled (it blocks in pdl_wait)
  if (b >= b1) return; // the whole warp leaves; there is no block-wide barrier below
  unsigned char *base = bsell_smem + size_t(warp) * warp_bytes;
  double *acc = reinterpret_cast<double *>(base);
  double *xs = acc + 32 * BS;          // the block's distinct outside rows, staged once
  double *sval = xs + size_t(max_nx) * BS;
  unsigned short *soff = reinterpret_cast<unsigned short *>(sval + max_int);
  unsigned char *scol = reinterpret_cast<unsigned char *>(soff) + 72;
  const int r0 = blk_row[b], nr = blk_row[b + 1] - r0;
  const int row = r0 + lane;
  const bool valid = lane < nr;
  double res[BS];
#pragma unroll
  for (int d = 0; d < BS; ++d) res[d] = 0.0;
  // ---- everything that does not depend on the previous sweep: row ids, the caller's x (not written inside the
  // chain), the inverse diagonal, the in-block entries (into shared memory) -- overlaps the previous sweep's tail
  int ro = 0;
  double di = 1.0;
  if (valid) {
    ro = order[row];
    if (DIR == 0) {
      const double *xi = io->x + int64_t(BS) * ro;
#pragma unroll
      for (int d = 0; d < BS; ++d) res[d] = xi[d];
    } else
      di = dinv[row];
  }
  const int ib = i_ptr[b], ni = i_ptr[b + 1] - ib;
  for (int k = lane; k < ni; k += 32) {
    sval[k] = __ldcs(i_val + ib + k);
    scol[k] = i_col[ib + k];
  }
  for (int k = lane; k < 33; k += 32) soff[k] = i_off[size_t(b) * 33 + k];
  const unsigned lens = e_len[b];
  const int eb = e_ptr[b];
  const int xb = STAGE ? x_ptr[b] : 0, nx = STAGE ? min(x_ptr[b + 1] - xb, max_nx) : 0;
  // ---- software prefetch: a warp walks its passes serially, each pass a dependent (col, val) -> gather round trip
  // to DRAM.  All the block's column indices are read once here and the rows they name -- and the value stream --
  // are requested into L2, so the passes below find their operands there.  L2 is the point of coherence, so this
  // is legal before pdl_wait(): rows the previous colour is still writing are simply updated in place.
  if (prefetch) {
    const int tot = int(lens & 255u) + int((lens >> 8) & 255u) + int((lens >> 16) & 255u) + int(lens >> 24);
    if (STAGE) {
      for (int k = lane; k < nx; k += 32) prefetch_l2(yp + int64_t(PS) * x_ids[xb + k]);
    } else {
      const int *gp0 = e_col + eb + lane;
#pragma unroll 4
      for (int k = 0; k < tot; ++k) prefetch_l2(yp + int64_t(PS) * __ldg(gp0 + k * 32));
    }
    if ((lane & 15) == 0) { // the value stream: 256 B per step, one request per 128-B line
      const double *vp0 = e_val + eb + lane;
      for (int k = 0; k < tot; ++k) prefetch_l2(vp0 + k * 32);
    }
  }
  // PIPE: first stage of the (col, val) pipeline and the block's (pass, slot) -> local row table, before the wait
  constexpr int U = 4;
  const int tot_ext = int(lens & 255u) + int((lens >> 8) & 255u) + int((lens >> 16) & 255u) + int(lens >> 24);
  [[maybe_unused]] int pc[U];
  [[maybe_unused]] double pv[U];
  [[maybe_unused]] int prow = 0;
  if constexpr (PIPE) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = u < tot_ext;
      pc[u] = ok ? (STAGE ? int(__ldcs(e_lix + eb + lane + u * 32)) : __ldcs(e_col + eb + lane + u * 32)) : 0;
      pv[u] = ok ? __ldcs(e_val + eb + lane + u * 32) : 0.0;
    }
    prow = int(e_prow[size_t(b) * 32 + lane]);
  }
  pdl_wait(); // the staging vector yp is read only from here on
  if (DIR == 1 && valid) {
    const double *yi = yp + int64_t(PS) * row;
#pragma unroll
    for (int d = 0; d < BS; ++d) res[d] = yi[d] * di;
  }
  // ---- stage the distinct rows of other blocks this block couples with (each is used ~3 times): ONE
  // round trip to HBM / L2 for all of them instead of one per step of the passes below
  if (STAGE) {
    for (int k = lane; k < nx; k += 32) {
      double x[BS];
      bsell_gather<BS>(yp, x_ids[xb + k], x);
#pragma unroll
      for (int d = 0; d < BS; ++d) xs[k * BS + d] = x[d];
    }
  }
  __syncwarp();
  // ---- entries coupling with other blocks: four passes of eight rows, four lanes per row
  if constexpr (PIPE) {
    const unsigned short *cp = e_lix + eb + lane;
    const int *gp = e_col + eb + lane;
    const double *vp = e_val + eb + lane;
    double a[BS];
#pragma unroll
    for (int d = 0; d < BS; ++d) a[d] = 0.0;
    int q = 0, bound = int(lens & 255u); // bound: step after the last one of pass q (cumulative)
    while (q < 4 && bound == 0) { // leading empty passes
      bsell_flush<BS>(a, acc, prow, q, lane);
      ++q;
      bound += q < 4 ? int((lens >> (8 * q)) & 255u) : 0;
    }
    for (int k0 = 0; k0 < tot_ext; k0 += U) {
      int nc[U];
      double nv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { // next stage of the pipeline
        const int kk = k0 + U + u;
        const bool ok = kk < tot_ext;
        nc[u] = ok ? (STAGE ? int(__ldcs(cp + kk * 32)) : __ldcs(gp + kk * 32)) : 0;
        nv[u] = ok ? __ldcs(vp + kk * 32) : 0.0;
      }
      double x[U][BS];
#pragma unroll
      for (int u = 0; u < U; ++u) { // the gathers of this stage, all in flight together
        if (k0 + u < tot_ext) {
          if (STAGE && pc[u] < max_nx) {
#pragma unroll
            for (int d = 0; d < BS; ++d) x[u][d] = xs[pc[u] * BS + d];
          } else
            bsell_gather<BS>(yp, STAGE ? __ldcs(gp + (k0 + u) * 32) : pc[u], x[u]);
        } else {
#pragma unroll
          for (int d = 0; d < BS; ++d) x[u][d] = 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k0 + u < tot_ext) { // warp-uniform
#pragma unroll
          for (int d = 0; d < BS; ++d) a[d] += pv[u] * x[u][d];
          while (q < 4 && k0 + u + 1 == bound) { // end of pass q (and of the empty passes that follow it)
            bsell_flush<BS>(a, acc, prow, q, lane);
            ++q;
            bound += q < 4 ? int((lens >> (8 * q)) & 255u) : 0;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { pc[u] = nc[u]; pv[u] = nv[u]; }
    }
    while (q < 4) { // trailing empty passes (or a block without outside entries)
      bsell_flush<BS>(a, acc, prow, q, lane);
      ++q;
    }
  } else {
    const unsigned short *cp = e_lix + eb + lane;
    const int *gp = e_col + eb + lane;
    const double *vp = e_val + eb + lane;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int len = int((lens >> (8 * q)) & 255u);
      double a[BS];
#pragma unroll
      for (int d = 0; d < BS; ++d) a[d] = 0.0;
#pragma unroll 4
      for (int k = 0; k < len; ++k) {
        const double v = __ldcs(vp + k * 32);
        const int li = STAGE ? int(__ldcs(cp + k * 32)) : 0;
        if (STAGE && li < max_nx) { // staged
          const double *x = xs + li * BS;
#pragma unroll
          for (int d = 0; d < BS; ++d) a[d] += v * x[d];
        } else { // not staged (or beyond the staging capacity): straight from the staging vector
          double x[BS];
          bsell_gather<BS>(yp, __ldcs(gp + k * 32), x);
#pragma unroll
          for (int d = 0; d < BS; ++d) a[d] += v * x[d];
        }
      }
      cp += len * 32;
      gp += len * 32;
      vp += len * 32;
#pragma unroll
      for (int o = 1; o < 4; o <<= 1)
#pragma unroll
        for (int d = 0; d < BS; ++d) a[d] += __shfl_xor_sync(FULL, a[d], o);
      if ((lane & 3) == 0) { // every local row sits in exactly one (pass, slot): acc needs no clearing
        const int lr = int(e_prow[size_t(b) * 32 + q * 8 + (lane >> 2)]);
#pragma unroll
        for (int d = 0; d < BS; ++d) acc[lr * BS + d] = a[d];
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int d = 0; d < BS; ++d) res[d] -= acc[lane * BS + d];
  // ---- entries inside the block: sequential elimination, finished rows broadcast by shuffle
  if (ni > 0) {
    int p = valid ? int(soff[lane]) : 0;
    const int pe = valid ? int(soff[lane + 1]) : 0;
    int cc = p < pe ? int(scol[p]) : 255;
    double cv = p < pe ? sval[p] : 0.0;
    // (visiting only the rows with in-block dependants through a bit mask was measured slower than this
    // fixed, unrolled loop: session J, profiles/r02)
#pragma unroll 4
    for (int step = 0; step < 32; ++step) {
      const int r = DIR == 0 ? step : 31 - step;
      if (r >= nr) continue; // warp-uniform
      double y[BS];
#pragma unroll
      for (int d = 0; d < BS; ++d) y[d] = __shfl_sync(FULL, res[d], r);
      if (cc == r) {
#pragma unroll
        for (int d = 0; d < BS; ++d) res[d] -= cv * y[d];
        ++p;
        if (p < pe) { cc = int(scol[p]); cv = sval[p]; }
        else cc = 255;
      }
    }
  }
  if (valid) {
    double *o = yp + int64_t(PS) * row;
#pragma unroll
    for (int d = 0; d < BS; ++d) o[d] = res[d];
    if (BS == 3 && DIR == 0) o[3] = 0.0;
    if (DIR == 1) {
      double *yo = io->y + int64_t(BS) * ro;
#pragma unroll
      for (int d = 0; d < BS; ++d) yo[d] = res[d];
    }
  }
}

// The blocks are independent, so runs of kBsellChunk consecutive blocks are packed by different host threads into
// chunk-local arrays (positions relative to the chunk) and the chunks are concatenated in order afterwards: the result
// does not depend on the number of threads.
namespace {
constexpr int kBsellChunk = 2048;
struct BsellChunk {
  std::vector<int> e_ptr, i_ptr, x_ptr;    // [blocks of the chunk + 1], relative to the chunk
  std::vector<int> e_map, e_gcol, i_map, x_ids;
  std::vector<unsigned short> e_col;
  std::vector<unsigned char> i_col;
  int max_nx = 0, max_int = 0;
  const char *error = nullptr;
};
template <typename T>
struct RawArray { // uninitialised storage, filled by append_chunks
  std::unique_ptr<T[]> p;
  size_t n = 0;
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
};
template <typename T>
void append_chunks(RawArray<T> &dst, const std::vector<BsellChunk> &chunks, std::vector<T> BsellChunk::*member)
{
  std::vector<size_t> at(chunks.size() + 1, 0);
  for (size_t c = 0; c < chunks.size(); ++c) at[c + 1] = at[c] + (chunks[c].*member).size();
  dst.n = at[chunks.size()];
  dst.p.reset(new T[dst.n + 1]); // pages first touched by the copying threads
  const int64_t nch = int64_t(chunks.size());
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t c = 0; c < nch; ++c) {
    const std::vector<T> &v = chunks[c].*member;
    if (!v.empty()) std::memcpy(dst.p.get() + at[c], v.data(), v.size() * sizeof(T));
  }
}
} // namespace

// host copies of one factor's block storage for the CPU emulation of the sweeps (bsell_host_check, test only)
struct BsellHost {
  std::vector<int> e_ptr, i_ptr, x_ptr, e_map, e_gcol, i_map, x_ids;
  std::vector<unsigned short> e_lix, i_off;
  std::vector<unsigned char> e_prow, i_col;
  std::vector<unsigned> e_len;
};

static void bsell_build_one(const std::vector<int> &rowptr, const std::vector<int> &colind,
                            const std::vector<int> &diagpos, const std::vector<int> &blk_ptr,
                            const std::vector<int> &colour_blk, bool lower, DevBsell &out, BsellHost *keep = nullptr)
{
  const int nb = int(blk_ptr.size()) - 1;
  std::vector<unsigned> e_len(nb, 0u), i_mask(nb, 0u); // i_mask: local rows that occur as an intra-block column
  std::vector<unsigned char> e_prow(size_t(nb) * 32, 0);
  std::vector<unsigned short> i_off(size_t(nb) * 33, 0);
  const int nchunks = (nb + kBsellChunk - 1) / kBsellChunk;
  std::vector<BsellChunk> chunks(nchunks);
#pragma omp parallel
  {
    // per-thread scratch
    std::vector<int> xloc(rowptr.size() - 1, -1); // factor row -> position in the current block's list
    std::vector<std::pair<int, int>> tmp;
    std::vector<std::vector<int>> ext(32); // CSR positions of the entries of each local row that leave the block
    std::vector<int> srt(32);
#pragma omp for schedule(dynamic, 1)
    for (int ch = 0; ch < nchunks; ++ch) {
      BsellChunk &C = chunks[ch];
      const int b_lo = ch * kBsellChunk, b_hi = std::min(nb, b_lo + kBsellChunk);
      C.e_ptr.assign(b_hi - b_lo + 1, 0); C.i_ptr.assign(b_hi - b_lo + 1, 0); C.x_ptr.assign(b_hi - b_lo + 1, 0);
      { // one allocation per array instead of repeated growth (large reallocations serialise on the page tables)
        const size_t nnz_chunk = size_t(rowptr[blk_ptr[b_hi]] - rowptr[blk_ptr[b_lo]]);
        C.e_col.reserve(nnz_chunk); C.e_gcol.reserve(nnz_chunk); C.e_map.reserve(nnz_chunk);
        C.i_col.reserve(nnz_chunk / 2); C.i_map.reserve(nnz_chunk / 2); C.x_ids.reserve(nnz_chunk / 2);
      }
      for (int b = b_lo; b < b_hi && !C.error; ++b) {
        const int lb = b - b_lo;
        const int r0 = blk_ptr[b], r1 = blk_ptr[b + 1];
        if (r1 - r0 > 32) { C.error = "bsell: block with more than 32 rows"; break; }
        for (int lr = 0; lr < 32; ++lr) ext[lr].clear();
        for (int r = r0; r < r1; ++r) {
          const int a = lower ? rowptr[r] : diagpos[r] + 1, z = lower ? diagpos[r] : rowptr[r + 1];
          for (int e = a; e < z; ++e) {
            const int c = colind[e];
            const bool intra = lower ? (c >= r0) : (c < r1);
            if (!intra) ext[r - r0].push_back(e);
          }
        }
        // distinct outside rows in first-use order
        for (int lr = 0; lr < 32; ++lr)
          for (int e : ext[lr]) {
            const int c = colind[e];
            if (xloc[c] < 0) { xloc[c] = int(C.x_ids.size()) - C.x_ptr[lb]; C.x_ids.push_back(c); }
          }
        C.x_ptr[lb + 1] = int(C.x_ids.size());
        const int nx = C.x_ptr[lb + 1] - C.x_ptr[lb];
        if (nx > 65535) C.error = "bsell: more than 65535 outside rows in a block";
        C.max_nx = std::max(C.max_nx, nx);
        // passes of eight rows with similar numbers of outside entries; local rows >= r1 - r0 are empty fillers
        for (int lr = 0; lr < 32; ++lr) srt[lr] = lr;
        std::stable_sort(srt.begin(), srt.end(), [&](int x, int y) { return ext[x].size() > ext[y].size(); });
        unsigned lens = 0;
        for (int q = 0; q < 4 && !C.error; ++q) {
          int len = 0;
          for (int j = 0; j < 8; ++j) len = std::max(len, (int(ext[srt[q * 8 + j]].size()) + 3) / 4);
          if (len > 255) { C.error = "bsell: more than 1020 outside entries in a row"; break; }
          lens |= unsigned(len) << (8 * q);
          const size_t base = C.e_col.size();
          C.e_col.resize(base + size_t(len) * 32, 0);
          C.e_gcol.resize(base + size_t(len) * 32, 0);
          C.e_map.resize(base + size_t(len) * 32, -1);
          for (int l = 0; l < 32; ++l) {
            const std::vector<int> &ex = ext[srt[q * 8 + l / 4]];
            for (int k = 0; k < len; ++k) {
              const size_t o = base + size_t(k) * 32 + l;
              const size_t e = size_t(k) * 4 + (l & 3);
              if (e < ex.size()) { C.e_col[o] = (unsigned short)xloc[colind[ex[e]]]; C.e_gcol[o] = colind[ex[e]]; C.e_map[o] = ex[e]; }
              else { C.e_col[o] = 0; C.e_gcol[o] = C.x_ids[C.x_ptr[lb]]; } // padding: value 0 times the block's first outside row
            }
          }
        }
        for (int k = C.x_ptr[lb]; k < C.x_ptr[lb + 1]; ++k) xloc[C.x_ids[k]] = -1;
        for (int t = 0; t < 32; ++t) e_prow[size_t(b) * 32 + t] = (unsigned char)srt[t];
        e_len[b] = lens;
        C.e_ptr[lb + 1] = int(C.e_col.size());
        const size_t ibase = C.i_col.size();
        for (int lr = 0; lr < 32; ++lr) {
          i_off[size_t(b) * 33 + lr] = (unsigned short)(C.i_col.size() - ibase);
          const int r = r0 + lr;
          if (r >= r1) continue;
          const int a = lower ? rowptr[r] : diagpos[r] + 1, z = lower ? diagpos[r] : rowptr[r + 1];
          tmp.clear();
          for (int e = a; e < z; ++e) {
            const int c = colind[e];
            const bool intra = lower ? (c >= r0) : (c < r1);
            if (intra) tmp.emplace_back(c - r0, e);
          }
          if (!lower) std::reverse(tmp.begin(), tmp.end()); // descending columns for the backward sweep
          for (auto &ce : tmp) { C.i_col.push_back((unsigned char)ce.first); C.i_map.push_back(ce.second); i_mask[b] |= 1u << ce.first; }
        }
        i_off[size_t(b) * 33 + 32] = (unsigned short)(C.i_col.size() - ibase);
        C.max_int = std::max(C.max_int, int(C.i_col.size() - ibase));
        C.i_ptr[lb + 1] = int(C.i_col.size());
      }
      if (C.error) // leave the scratch clean for the next chunk of this thread
        std::fill(xloc.begin(), xloc.end(), -1);
    }
  }
  for (const BsellChunk &C : chunks)
    if (C.error) throw StateError(C.error);
  // concatenate: per-block pointers shifted by the chunk's base
  std::vector<int> e_ptr(nb + 1, 0), i_ptr(nb + 1, 0), x_ptr(nb + 1, 0);
  int max_nx = 0, max_int = 0;
  {
    int64_t eb = 0, ib = 0, xb = 0;
    for (int ch = 0; ch < nchunks; ++ch) {
      const BsellChunk &C = chunks[ch];
      const int b_lo = ch * kBsellChunk, cnt = int(C.e_ptr.size()) - 1;
      if (eb + int64_t(C.e_col.size()) > int64_t(0x7fffffff)) throw StateError("bsell: more than 2^31 slots");
      for (int lb = 0; lb < cnt; ++lb) {
        e_ptr[b_lo + lb + 1] = int(eb) + C.e_ptr[lb + 1];
        i_ptr[b_lo + lb + 1] = int(ib) + C.i_ptr[lb + 1];
        x_ptr[b_lo + lb + 1] = int(xb) + C.x_ptr[lb + 1];
      }
      eb += int64_t(C.e_col.size()); ib += int64_t(C.i_col.size()); xb += int64_t(C.x_ids.size());
      max_nx = std::max(max_nx, C.max_nx);
      max_int = std::max(max_int, C.max_int);
    }
  }
  RawArray<int> e_map, e_gcol, i_map, x_ids;
  RawArray<unsigned short> e_col; // index into the block's list of distinct outside rows
  RawArray<unsigned char> i_col;
  append_chunks(e_map, chunks, &BsellChunk::e_map);
  append_chunks(e_gcol, chunks, &BsellChunk::e_gcol); // the same entries as factor rows (kernels without staging)
  append_chunks(i_map, chunks, &BsellChunk::i_map);
  append_chunks(x_ids, chunks, &BsellChunk::x_ids);
  append_chunks(e_col, chunks, &BsellChunk::e_col);
  append_chunks(i_col, chunks, &BsellChunk::i_col);
  chunks.clear();
  out.n_blocks = nb;
  out.max_int = max_int;
  out.max_nx = max_nx;
  // shared memory is sized per launch (= per block colour) by the largest list of that colour
  out.col_max_nx.assign(colour_blk.size() > 0 ? colour_blk.size() - 1 : 0, 0);
  for (size_t c = 0; c + 1 < colour_blk.size(); ++c)
    for (int b = colour_blk[c]; b < colour_blk[c + 1]; ++b)
      out.col_max_nx[c] = std::min(bsell_xcap(), std::max(out.col_max_nx[c], x_ptr[b + 1] - x_ptr[b]));
  if (getenv("NSB_VERBOSE") && atoi(getenv("NSB_VERBOSE")) > 0)
    std::fprintf(stderr, "[nsb bsell %s] blocks %d, ext slots %zu (%.1f per row), intra %zu, outside rows %zu (max %d per block), max intra %d\n",
                 lower ? "L" : "U", nb, e_col.size(), double(e_col.size()) / std::max(1, blk_ptr[nb]), i_col.size(), x_ids.size(),
                 max_nx, max_int);
  out.x_ptr.upload(x_ptr);
  if (x_ids.empty()) out.x_ids.upload(std::vector<int>(1, 0));
  else out.x_ids.upload(x_ids.p.get(), x_ids.n);
  out.n_ext = int64_t(e_col.size());
  out.n_int = int64_t(i_col.size());
  out.e_ptr.upload(e_ptr); out.e_lix.upload(e_col.p.get(), e_col.n); out.e_col.upload(e_gcol.p.get(), e_gcol.n);
  out.e_map.upload(e_map.p.get(), e_map.n);
  out.e_len.upload(e_len); out.e_prow.upload(e_prow);
  out.e_val.alloc(e_col.size());
  out.i_ptr.upload(i_ptr); out.i_map.upload(i_map.p.get(), i_map.n); out.i_off.upload(i_off); out.i_col.upload(i_col.p.get(), i_col.n);
  out.i_mask.upload(i_mask);
  out.i_val.alloc(i_col.size());
  if (keep) {
    keep->e_ptr = e_ptr; keep->i_ptr = i_ptr; keep->x_ptr = x_ptr; keep->e_len = e_len; keep->e_prow = e_prow; keep->i_off = i_off;
    keep->e_map.assign(e_map.p.get(), e_map.p.get() + e_map.n);
    keep->e_gcol.assign(e_gcol.p.get(), e_gcol.p.get() + e_gcol.n);
    keep->i_map.assign(i_map.p.get(), i_map.p.get() + i_map.n);
    keep->x_ids.assign(x_ids.p.get(), x_ids.p.get() + x_ids.n);
    keep->e_lix.assign(e_col.p.get(), e_col.p.get() + e_col.n);
    keep->i_col.assign(i_col.p.get(), i_col.p.get() + i_col.n);
  }
}

// ---- CPU emulation of k_bsell on the packed storage (nsb_debug_bsell_check, tests/test_host_cpu.py) ---------------
// Builds the block storage of both factors exactly as bsell_build does (the uploads are skipped: setup dry run), fills
// it with a pseudo-random factor and walks it the way the kernel does -- per block: the four passes of eight rows with
// four lanes per row over the entries that leave the block (staged list index below `xcap`, factor row otherwise),
// (pass, slot) -> local row through e_prow, then the sequential elimination over the in-block entries -- colour by
// colour, the blocks of a colour in REVERSE order (they must be independent).  Returns the largest difference to plain
// forward / backward substitution on the permuted CSR pattern, relative to the largest entry; 1e30 and above name a
// structural violation.  stats[4]: blocks, colours, max outside rows per block, max in-block entries per block.
double bsell_host_check(const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
                        const std::vector<int> &blk_ptr, const std::vector<int> &colour_blk, int bs, int xcap, int *stats)
{
  const int n = int(rowptr.size()) - 1, nb = int(blk_ptr.size()) - 1, ncol = int(colour_blk.size()) - 1;
  DevBsell dL, dU;
  BsellHost hL, hU;
  const bool was_on = g_dry.on;
  g_dry.on = true; // no device: the uploads inside bsell_build_one are recorded, not copied
  try {
    bsell_build_one(rowptr, colind, diagpos, blk_ptr, colour_blk, true, dL, &hL);
    bsell_build_one(rowptr, colind, diagpos, blk_ptr, colour_blk, false, dU, &hU);
  } catch (...) { g_dry.on = was_on; throw; }
  g_dry.on = was_on;
  if (!was_on) g_dry.log.clear();
  if (stats) { stats[0] = nb; stats[1] = ncol; stats[2] = std::max(dL.max_nx, dU.max_nx); stats[3] = std::max(dL.max_int, dU.max_int); }
  // pseudo-random factor: small off-diagonal entries (a well-conditioned solve), inverse diagonal in [0.5, 1.5]
  auto rnd = [](uint64_t k) { k = (k + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull; k ^= k >> 31; k *= 0x94D049BB133111EBull; k ^= k >> 29;
                              return double(k >> 11) / double(1ull << 53); };
  std::vector<double> val(colind.size()), dinv(n), x(size_t(n) * bs);
  for (int k = 0; k < n; ++k) {
    const int len = std::max(1, rowptr[k + 1] - rowptr[k]);
    for (int e = rowptr[k]; e < rowptr[k + 1]; ++e) val[e] = (rnd(uint64_t(e)) - 0.5) / len;
    dinv[k] = 0.5 + rnd(uint64_t(k) + (1ull << 40));
    for (int d = 0; d < bs; ++d) x[size_t(k) * bs + d] = rnd(uint64_t(k) * 3 + d + (1ull << 41)) - 0.5;
  }
  // reference: y = U^-1 D^-1 L^-1 x with Ifpack's scaling (kernels_linalg.cu, k_ilu_factor_level)
  std::vector<double> ref(x);
  for (int k = 0; k < n; ++k)
    for (int e = rowptr[k]; e < diagpos[k]; ++e)
      for (int d = 0; d < bs; ++d) ref[size_t(k) * bs + d] -= val[e] * ref[size_t(colind[e]) * bs + d];
  for (int k = n - 1; k >= 0; --k) {
    for (int d = 0; d < bs; ++d) ref[size_t(k) * bs + d] *= dinv[k];
    for (int e = diagpos[k] + 1; e < rowptr[k + 1]; ++e)
      for (int d = 0; d < bs; ++d) ref[size_t(k) * bs + d] -= val[e] * ref[size_t(colind[e]) * bs + d];
  }
  // emulation
  std::vector<double> y(x);
  const double nan = std::nan("");
  for (int dir = 0; dir < 2; ++dir) {
    const BsellHost &B = dir == 0 ? hL : hU;
    const DevBsell &D = dir == 0 ? dL : dU;
    std::vector<double> e_val(B.e_map.size()), i_val(B.i_map.size());
    for (size_t o = 0; o < e_val.size(); ++o) e_val[o] = B.e_map[o] >= 0 ? val[B.e_map[o]] : 0.0; // k_sell_fill
    for (size_t o = 0; o < i_val.size(); ++o) i_val[o] = val[B.i_map[o]];
    for (int cc = 0; cc < ncol; ++cc) {
      const int c = dir == 0 ? cc : ncol - 1 - cc;
      const int cap = std::min(xcap, D.col_max_nx[c]); // rows of the staged list beyond the capacity come from e_col
      for (int b = colour_blk[c + 1] - 1; b >= colour_blk[c]; --b) {
        const int r0 = blk_ptr[b], nr = blk_ptr[b + 1] - r0;
        if (nr < 1 || nr > 32) return 1e30;
        double res[32][3] = {}, acc[32][3];
        for (int l = 0; l < 32; ++l)
          for (int d = 0; d < bs; ++d) {
            acc[l][d] = nan;
            if (l < nr) res[l][d] = dir == 0 ? y[size_t(r0 + l) * bs + d] : y[size_t(r0 + l) * bs + d] * dinv[r0 + l];
          }
        size_t pos = size_t(B.e_ptr[b]);
        if (pos % 32) return 2e30;
        const unsigned lens = B.e_len[b];
        for (int q = 0; q < 4; ++q) {
          const int len = int((lens >> (8 * q)) & 255u);
          double a[8][3] = {};
          for (int k = 0; k < len; ++k)
            for (int l = 0; l < 32; ++l) {
              const size_t o = pos + size_t(k) * 32 + l;
              const int li = int(B.e_lix[o]);
              if (B.e_map[o] >= 0 && (li >= B.x_ptr[b + 1] - B.x_ptr[b] || B.x_ids[B.x_ptr[b] + li] != B.e_gcol[o])) return 3e30;
              const int col = li < cap ? B.x_ids[B.x_ptr[b] + li] : B.e_gcol[o];
              if (col >= r0 && col < r0 + nr) return 4e30; // an outside entry must leave the block
              for (int d = 0; d < bs; ++d) a[l / 4][d] += e_val[o] * y[size_t(col) * bs + d];
            }
          pos += size_t(len) * 32;
          for (int slot = 0; slot < 8; ++slot) {
            const int lr = int(B.e_prow[size_t(b) * 32 + q * 8 + slot]);
            if (lr > 31 || !std::isnan(acc[lr][0])) return 5e30; // every local row sits in exactly one (pass, slot)
            for (int d = 0; d < bs; ++d) acc[lr][d] = a[slot][d];
          }
        }
        if (pos != size_t(B.e_ptr[b + 1])) return 6e30;
        for (int l = 0; l < 32; ++l)
          for (int d = 0; d < bs; ++d) res[l][d] -= acc[l][d];
        // in-block entries: sequential elimination, one cursor per local row
        const int ib = B.i_ptr[b];
        int p[32], pe[32];
        for (int l = 0; l < 32; ++l) { p[l] = l < nr ? int(B.i_off[size_t(b) * 33 + l]) : 0; pe[l] = l < nr ? int(B.i_off[size_t(b) * 33 + l + 1]) : 0; }
        for (int step = 0; step < 32; ++step) {
          const int r = dir == 0 ? step : 31 - step;
          if (r >= nr) continue;
          for (int l = 0; l < nr; ++l)
            if (p[l] < pe[l] && int(B.i_col[ib + p[l]]) == r) {
              if ((dir == 0 && r >= l) || (dir == 1 && r <= l)) return 7e30; // L: earlier rows only, U: later rows only
              for (int d = 0; d < bs; ++d) res[l][d] -= i_val[ib + p[l]] * res[r][d];
              ++p[l];
            }
        }
        for (int l = 0; l < nr; ++l)
          if (p[l] != pe[l]) return 8e30; // an in-block entry the sweep never reaches
        if (ib + int(B.i_off[size_t(b) * 33 + 32]) != B.i_ptr[b + 1]) return 9e30;
        for (int l = 0; l < nr; ++l)
          for (int d = 0; d < bs; ++d) y[size_t(r0 + l) * bs + d] = res[l][d];
      }
    }
  }
  double err = 0, scale = 0;
  for (size_t k = 0; k < y.size(); ++k) { err = std::max(err, std::fabs(y[k] - ref[k])); scale = std::max(scale, std::fabs(ref[k])); }
  return err / std::max(scale, 1e-300);
}


void bsell_build(DevIlu &ilu, const std::vector<int> &rowptr, const std::vector<int> &colind,
                 const std::vector<int> &diagpos, const std::vector<int> &blk_ptr, const std::vector<int> &colour_blk)
{
  ilu.blk_row.upload(blk_ptr);
  ilu.colour_blk = colour_blk;
  bsell_build_one(rowptr, colind, diagpos, blk_ptr, colour_blk, true, ilu.bL);
  bsell_build_one(rowptr, colind, diagpos, blk_ptr, colour_blk, false, ilu.bU);
  // a colour whose blocks stage long lists may need more than the default 48 KB of dynamic shared memory
  // (set here: the launches happen inside a stream capture)
  const size_t need = kBW * std::max(bsell_warp_bytes(ilu.bs_rhs, ilu.bL.max_int, std::min(bsell_xcap(), ilu.bL.max_nx)),
                                     bsell_warp_bytes(ilu.bs_rhs, ilu.bU.max_int, std::min(bsell_xcap(), ilu.bU.max_nx)));
  if (need > size_t(227) * 1024) throw StateError("bsell: a block does not fit shared memory");
  if (need > size_t(48) * 1024) {
    const int lim = int(need);
    auto raise = [&](auto kernel) { NSB_CUDA_SETUP(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim)); };
    if (ilu.bs_rhs == 3) { raise(k_bsell<3, 0, false, false>); raise(k_bsell<3, 1, false, false>); raise(k_bsell<3, 0, false, true>); raise(k_bsell<3, 1, false, true>); }
    else if (ilu.bs_rhs == 2) { raise(k_bsell<2, 0, false, false>); raise(k_bsell<2, 1, false, false>); raise(k_bsell<2, 0, false, true>); raise(k_bsell<2, 1, false, true>); }
    else { raise(k_bsell<1, 0, true, false>); raise(k_bsell<1, 1, true, false>); raise(k_bsell<1, 0, true, true>); raise(k_bsell<1, 1, true, true>); }
  }
}

void bsell_fill(Handle &H, DevIlu &ilu)
{
  auto fill = [&](int64_t n, const int *map, double *val) {
    if (n == 0) return;
    k_sell_fill<<<unsigned(std::min<int64_t>((n + 255) / 256, kSM * 16)), 256, 0, H.stream>>>(n, map, ilu.val.p, val);
    H.launches++;
  };
  fill(ilu.bL.n_ext, ilu.bL.e_map.p, ilu.bL.e_val.p);
  fill(ilu.bL.n_int, ilu.bL.i_map.p, ilu.bL.i_val.p);
  fill(ilu.bU.n_ext, ilu.bU.e_map.p, ilu.bU.e_val.p);
  fill(ilu.bU.n_int, ilu.bU.i_map.p, ilu.bU.i_val.p);
}

template <int BS, int DIR>
static void launch_bsell(cudaStream_t s, const DevIlu &ilu, const DevBsell &B, int colour, int b0, int b1, double *yp,
                         const TrsvIo *io, bool pdl)
{
  // staging pays for the pressure matrix (long rows, one right-hand side: the sweeps are latency-bound) and
  // costs occupancy for the 3-component velocity block (measured, profiles/r02 sessions H-J)
  constexpr bool STAGE = BS == 1;
  const int max_nx = STAGE ? B.col_max_nx[colour] : 0;
  const size_t wb = bsell_warp_bytes(BS, B.max_int, max_nx);
  const unsigned grid = unsigned((b1 - b0 + kBW - 1) / kBW);
  auto go = [&](auto kernel) {
    launch_k(kernel, grid, kBW * 32, wb * kBW, s, pdl, b0, b1, ilu.blk_row.p, B.e_ptr.p, B.e_len.p, B.e_prow.p, B.e_lix.p, B.e_col.p,
             B.e_val.p, B.x_ptr.p, B.x_ids.p, B.i_ptr.p, B.i_off.p, B.i_col.p, B.i_val.p, B.i_mask.p, yp, ilu.dinv.p, ilu.order.p, io,
             B.max_int, max_nx, int(wb), bsell_prefetch());
  };
  if (bsell_pipe(BS)) go(k_bsell<BS, DIR, STAGE, true>);
  else go(k_bsell<BS, DIR, STAGE, false>);
}

template <int BS>
static void bsell_trsv_t(Handle &H, DevIlu &ilu, double *yp, cudaStream_t s)
{
  const int nc = int(ilu.colour_blk.size()) - 1;
  const TrsvIo *io = reinterpret_cast<const TrsvIo *>(ilu.io.p);
  const bool pdl = pdl_enabled();
  bool chained = false; // the first sweep of the chain is an ordinary launch
  for (int c = 0; c < nc; ++c) {
    const int a = ilu.colour_blk[c], b = ilu.colour_blk[c + 1];
    if (b <= a) continue;
    launch_bsell<BS, 0>(s, ilu, ilu.bL, c, a, b, yp, io, pdl && chained);
    chained = true;
    H.launches++;
  }
  for (int c = nc - 1; c >= 0; --c) {
    const int a = ilu.colour_blk[c], b = ilu.colour_blk[c + 1];
    if (b <= a) continue;
    launch_bsell<BS, 1>(s, ilu, ilu.bU, c, a, b, yp, io, pdl && chained);
    chained = true;
    H.launches++;
  }
}

// yp: staging vector in factor order (bsell_stride doubles per row); in / out through ilu.io
void bsell_trsv(Handle &H, DevIlu &ilu, double *yp, cudaStream_t s)
{
  if (ilu.bs_rhs == 3) bsell_trsv_t<3>(H, ilu, yp, s);
  else if (ilu.bs_rhs == 2) bsell_trsv_t<2>(H, ilu, yp, s);
  else bsell_trsv_t<1>(H, ilu, yp, s);
  NSB_CUDA(cudaGetLastError());
}

void sell_set_io(Handle &H, DevIlu &ilu, const double *x, double *y)
{
  k_set_io<<<1, 1, 0, H.stream>>>(reinterpret_cast<TrsvIo *>(ilu.io.p), x, y);
  H.launches++;
}

} // namespace nsb
