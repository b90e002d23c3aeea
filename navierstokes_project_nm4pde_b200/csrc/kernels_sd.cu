// kernels_sd.cu -- subdomain-resident ILU(0) triangular solves (ilu_ordering = 3).
//
// Replaces TrilinosWrappers::PreconditionILU::vmult = U^-1 D^-1 L^-1 (Ifpack_ILU::ApplyInverse,
// include/Preconditioners.hpp:319-320) for the factors in the two-level "subdomain" ordering built by
// subdomain_order (kernels_linalg.cu): [interior rows of part 0 | part 1 | ... | separator rows].
//
// Interior rows (k_sd_trsv): ONE CTA solves ONE part start to finish.  The part's slice of the vector
// (and, in the backward solve, the separator values it couples with: its "ring") is staged into shared
// memory once; the part's factor entries -- packed in processing order, 16-bit part-local column
// indices, 10 bytes per entry -- are streamed from HBM exactly once; colours inside the part are
// separated by __syncthreads() instead of kernel boundaries.  Every vector entry is therefore read
// from and written to HBM once per solve, against ~2.6 times the algorithmic bytes for global colour
// sweeps, and ~85% of the rows need one launch instead of one per colour.
// Separator rows (k_sd_sep): SELL-32 colour sweeps over the global staging vector, as before but on
// ~15% of the rows.
//
// Slice layout (interior): rows of one colour of one part are cut into slices of 32 / LPR rows; LPR
// adjacent lanes share a row (entry e of a row sits in lane e % LPR of step e / LPR), partial sums
// are combined by shuffles.  A slice occupies len * 40 doubles of the part's stream:
//   [ len x 32 doubles: values, step-major ][ len x 32 uint16: part-local columns ].
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>

#include "nsb_internal.hpp"

namespace nsb {

constexpr int kSdThreads = 256;
constexpr int kSdLpr = 4;     // lanes per row in the interior slices
constexpr int kSM_sd = 148;

struct TrsvIoSd { const double *x; double *y; }; // same slot as TrsvIo in kernels_sell.cu (set by k_set_io)

template <int BS>
__device__ __forceinline__ void sd_gather_global(const double *yp, int c, double (&x)[BS])
{
  if constexpr (BS == 3) {
    asm("{ .reg .f64 pad; ld.global.v4.f64 {%0,%1,%2,pad}, [%3]; }" : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]) : "l"(yp + 4 * int64_t(c)));
  } else if constexpr (BS == 2) {
    asm("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(x[0]), "=d"(x[1]) : "l"(yp + 2 * int64_t(c)));
  } else {
    asm("ld.global.f64 %0, [%1];" : "=d"(x[0]) : "l"(yp + c));
  }
}

// DIR 0: forward substitution  y = x - L y          (unit diagonal; Ifpack stores L scaled by dinv_j)
// DIR 1: backward substitution z = y * dinv - U z   (Ifpack stores U scaled by dinv_i)
// yp: staging vector in factor order, PS doubles per row.  The permutation into factor order is fused
// into the forward kernel (reads the caller's x through `order`), the inverse permutation into the
// backward kernel (writes the caller's y).
template <int BS, int DIR>
__global__ void __launch_bounds__(kSdThreads) k_sd_trsv(const SdPart *__restrict__ parts, const int4 *__restrict__ slices,
                                                        const int *__restrict__ cslice, const double *__restrict__ stream,
                                                        const int *__restrict__ ring_rows, double *yp,
                                                        const double *__restrict__ dinv, const int *__restrict__ order,
                                                        const TrsvIoSd *__restrict__ io)
{
  constexpr int PS = BS == 3 ? 4 : BS;
  constexpr int LPR = kSdLpr, NW = kSdThreads / 32, U = 4;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) double sd_ys[]; // [(ni + nring)][BS]
  const SdPart P = parts[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ni = P.ni;
  // ---- stage in
  if (DIR == 0) {
    const double *x = io->x;
    for (int l = tid; l < ni; l += kSdThreads) {
      const double *xi = x + int64_t(BS) * order[P.row0 + l];
#pragma unroll
      for (int d = 0; d < BS; ++d) sd_ys[l * BS + d] = xi[d];
    }
  } else {
    for (int t = tid; t < ni * BS; t += kSdThreads) {
      const int l = t / BS, d = t - l * BS;
      sd_ys[t] = yp[int64_t(PS) * (P.row0 + l) + d];
    }
    for (int j = tid; j < P.nring; j += kSdThreads) {
      double v[BS];
      sd_gather_global<BS>(yp, ring_rows[P.ring0 + j], v);
#pragma unroll
      for (int d = 0; d < BS; ++d) sd_ys[(ni + j) * BS + d] = v[d];
    }
  }
  __syncthreads();
  // ---- colours of the part, in processing order
  for (int c = 0; c < P.ncol; ++c) {
    const int s0 = cslice[P.cs0 + c], s1 = cslice[P.cs0 + c + 1];
    for (int s = s0 + warp; s < s1; s += NW) {
      const int4 S = slices[s]; // x: offset in the stream / 8 doubles, y: len, z: first local row, w: rows
      const int len = S.y;
      const double *vp = stream + int64_t(unsigned(S.x)) * 8 + lane;
      const unsigned short *cp = reinterpret_cast<const unsigned short *>(stream + int64_t(unsigned(S.x)) * 8 + int64_t(len) * 32) + lane;
      double acc[BS];
#pragma unroll
      for (int d = 0; d < BS; ++d) acc[d] = 0.0;
      double v[U], nv[U];
      int ix[U], nix[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = u < len;
        v[u] = ok ? __ldcs(vp + u * 32) : 0.0;
        ix[u] = ok ? int(__ldcs(cp + u * 32)) : 0;
      }
      for (int k = 0; k < len; k += U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool ok = k + U + u < len;
          nv[u] = ok ? __ldcs(vp + (k + U + u) * 32) : 0.0;
          nix[u] = ok ? int(__ldcs(cp + (k + U + u) * 32)) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const double *yb = sd_ys + ix[u] * BS;
#pragma unroll
          for (int d = 0; d < BS; ++d) acc[d] += v[u] * yb[d];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { v[u] = nv[u]; ix[u] = nix[u]; }
      }
#pragma unroll
      for (int o = 1; o < LPR; o <<= 1)
#pragma unroll
        for (int d = 0; d < BS; ++d) acc[d] += __shfl_xor_sync(FULL, acc[d], o);
      const int lr = lane / LPR;
      if ((lane % LPR) == 0 && lr < S.w) {
        const int row = S.z + lr;
        double *yr = sd_ys + row * BS;
        if (DIR == 0) {
#pragma unroll
          for (int d = 0; d < BS; ++d) yr[d] -= acc[d];
        } else {
          const double di = dinv[P.row0 + row];
#pragma unroll
          for (int d = 0; d < BS; ++d) yr[d] = yr[d] * di - acc[d];
        }
      }
    }
    __syncthreads();
  }
  // ---- stage out
  if (DIR == 0) {
    for (int t = tid; t < ni * PS; t += kSdThreads) {
      const int l = t / PS, d = t - l * PS;
      yp[int64_t(PS) * P.row0 + t] = d < BS ? sd_ys[l * BS + d] : 0.0;
    }
  } else {
    double *y = io->y;
    for (int l = tid; l < ni; l += kSdThreads) {
      double *yo = y + int64_t(BS) * order[P.row0 + l];
#pragma unroll
      for (int d = 0; d < BS; ++d) yo[d] = sd_ys[l * BS + d];
    }
  }
}

// Separator rows: one colour per launch, SELL-32 (one thread per row) over the global staging vector.
// DIR 0: yp[r] = x[order[r]] - sum ; DIR 1: yp[r] = yp[r] * dinv[r] - sum, also stored to the caller's y.
template <int BS, int DIR>
__global__ void __launch_bounds__(256) k_sd_sep(int s0, int s1, const int *__restrict__ slice_ptr, const int *__restrict__ rowid,
                                                const int *__restrict__ col, const double *__restrict__ val, double *yp,
                                                const double *__restrict__ dinv, const int *__restrict__ order,
                                                const TrsvIoSd *__restrict__ io)
{
  constexpr int PS = BS == 3 ? 4 : BS, U = 4;
  const int lane = threadIdx.x & 31;
  const int s = s0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (s >= s1) return;
  const int base = slice_ptr[s];
  const int len = (slice_ptr[s + 1] - base) >> 5;
  const int r = rowid[(int64_t(s) << 5) + lane];
  double b[BS], acc[BS];
#pragma unroll
  for (int d = 0; d < BS; ++d) { b[d] = 0.0; acc[d] = 0.0; }
  double di = 1.0;
  int ro = 0;
  if (r >= 0) {
    ro = order[r];
    if (DIR == 0) {
      const double *xi = io->x + int64_t(BS) * ro;
#pragma unroll
      for (int d = 0; d < BS; ++d) b[d] = xi[d];
    } else {
      di = dinv[r];
      const double *yi = yp + int64_t(PS) * r;
#pragma unroll
      for (int d = 0; d < BS; ++d) b[d] = yi[d];
    }
  }
  const int *cp = col + base + lane;
  const double *vp = val + base + lane;
  int c[U], nc[U];
  double v[U], nv[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const bool ok = u < len;
    c[u] = ok ? __ldcs(cp + u * 32) : 0;
    v[u] = ok ? __ldcs(vp + u * 32) : 0.0;
  }
  for (int k = 0; k < len; k += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = k + U + u < len;
      nc[u] = ok ? __ldcs(cp + (k + U + u) * 32) : 0;
      nv[u] = ok ? __ldcs(vp + (k + U + u) * 32) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double x[BS];
      sd_gather_global<BS>(yp, c[u], x);
#pragma unroll
      for (int d = 0; d < BS; ++d) acc[d] += v[u] * x[d];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) { c[u] = nc[u]; v[u] = nv[u]; }
  }
  if (r >= 0) {
    double *o = yp + int64_t(PS) * r;
    if (DIR == 0) {
#pragma unroll
      for (int d = 0; d < BS; ++d) o[d] = b[d] - acc[d];
      if (BS == 3) o[3] = 0.0;
    } else {
      double *yo = io->y + int64_t(BS) * ro;
#pragma unroll
      for (int d = 0; d < BS; ++d) {
        const double z = b[d] * di - acc[d];
        o[d] = z;
        yo[d] = z;
      }
    }
  }
}

// one warp per slice: values of the packed stream from the factor values (padding slots = 0)
__global__ void k_sd_fill(int n_slices, const int4 *__restrict__ slices, const int64_t *__restrict__ map_off,
                          const int *__restrict__ map, const double *__restrict__ src, double *__restrict__ stream)
{
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slices; s += warps) {
    const int4 S = slices[s];
    double *vp = stream + int64_t(unsigned(S.x)) * 8;
    const int *mp = map + map_off[s];
    for (int k = 0; k < S.y; ++k) {
      const int m = mp[k * 32 + lane];
      vp[k * 32 + lane] = m >= 0 ? src[m] : 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side: pack the interior rows of one triangular factor
// rowptr / colind / diagpos: the factor pattern in factor order (diagonal inside).
struct HostSdTri {
  std::vector<int4> slices;
  std::vector<int64_t> map_off;
  std::vector<int> cslice, map, ring_rows;
  std::vector<double> stream;
  int max_local = 0;
  int64_t n_doubles = 0;
};

static void sd_pack(const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
                    const std::vector<int> &part_ptr, const std::vector<int> &pcol_ptr, const std::vector<int> &pcol,
                    bool lower, int n_interior, HostSdTri &out, std::vector<SdPart> &parts, bool fill_parts)
{
  constexpr int LPR = kSdLpr, RPS = 32 / LPR;
  const int np = int(part_ptr.size()) - 1;
  std::vector<int4> &slices = out.slices;
  std::vector<int64_t> &map_off = out.map_off;
  std::vector<int> &cslice = out.cslice, &ring_rows = out.ring_rows;
  slices.clear(); map_off.clear(); cslice.clear(); ring_rows.clear();
  // pass 1: slices, sizes, rings
  std::vector<int> ring_local; // factor row of a separator -> position in the current part's ring
  if (!lower) ring_local.assign(rowptr.size() - 1, -1);
  int64_t n_doubles = 0, n_slots = 0;
  struct PartTmp { int slice0, cs0, ncol, ring0, nring; };
  std::vector<PartTmp> pt(np);
  for (int p = 0; p < np; ++p) {
    const int r0 = part_ptr[p], r1 = part_ptr[p + 1];
    const int ncol = pcol_ptr[p + 1] - pcol_ptr[p] - 1;
    pt[p].slice0 = int(slices.size());
    pt[p].cs0 = int(cslice.size());
    pt[p].ncol = ncol;
    pt[p].ring0 = int(ring_rows.size());
    if (!lower) // ring: separator rows the part's U entries couple with, in first-use order
      for (int r = r0; r < r1; ++r)
        for (int e = diagpos[r] + 1; e < rowptr[r + 1]; ++e) {
          const int cc = colind[e];
          if (cc >= r1) {
            if (cc < n_interior) throw StateError("subdomain ILU: an interior row couples with another part");
            if (ring_local[cc] < 0) { ring_local[cc] = int(ring_rows.size()) - pt[p].ring0; ring_rows.push_back(cc); }
          }
        }
    pt[p].nring = int(ring_rows.size()) - pt[p].ring0;
    if ((r1 - r0) + pt[p].nring > 65535) throw StateError("subdomain ILU: part too large for 16-bit local indices");
    for (int ci = 0; ci < ncol; ++ci) {
      const int c = lower ? ci : ncol - 1 - ci; // backward solve: colours in reverse
      const int a = pcol[pcol_ptr[p] + c], b = pcol[pcol_ptr[p] + c + 1];
      cslice.push_back(int(slices.size()));
      for (int s = a; s < b; s += RPS) {
        const int nr = std::min(RPS, b - s);
        int len = 0;
        for (int r = s; r < s + nr; ++r) {
          const int rl = lower ? diagpos[r] - rowptr[r] : rowptr[r + 1] - diagpos[r] - 1;
          len = std::max(len, (rl + LPR - 1) / LPR);
        }
        if (len == 0 && lower) continue; // forward: nothing to subtract; backward: the row is still scaled by dinv
        if (n_doubles / 8 > int64_t(0xffffffffu)) throw StateError("subdomain ILU: stream too large");
        slices.push_back(make_int4(int(unsigned(n_doubles / 8)), len, s - r0, nr));
        map_off.push_back(n_slots);
        n_doubles += int64_t(len) * 40;
        n_slots += int64_t(len) * 32;
      }
    }
    cslice.push_back(int(slices.size()));
    if (!lower)
      for (int j = pt[p].ring0; j < int(ring_rows.size()); ++j) ring_local[ring_rows[j]] = -1;
  }
  // pass 2: columns and the fill map
  std::vector<double> &stream = out.stream;
  std::vector<int> &map = out.map;
  stream.assign(size_t(n_doubles), 0.0);
  map.assign(size_t(n_slots), -1);
#pragma omp parallel
  {
    std::vector<int> rl; // per-thread ring lookup (factor row -> ring position), sparse reset
    if (!lower) rl.assign(rowptr.size() - 1, -1);
#pragma omp for schedule(dynamic, 8)
    for (int p = 0; p < np; ++p) {
      const int r0 = part_ptr[p], r1 = part_ptr[p + 1], ni = r1 - r0;
      if (!lower)
        for (int j = 0; j < pt[p].nring; ++j) rl[ring_rows[pt[p].ring0 + j]] = j;
      const int sl0 = pt[p].slice0, sl1 = p + 1 < np ? pt[p + 1].slice0 : int(slices.size());
      for (int s = sl0; s < sl1; ++s) {
        const int4 S = slices[s];
        unsigned short *ip = reinterpret_cast<unsigned short *>(stream.data() + int64_t(unsigned(S.x)) * 8 + int64_t(S.y) * 32);
        int *mp = map.data() + map_off[s];
        for (int l = 0; l < 32; ++l) {
          const int lr = l / LPR, q = l % LPR;
          const int r = lr < S.w ? r0 + S.z + lr : -1;
          const int ea = r < 0 ? 0 : (lower ? rowptr[r] : diagpos[r] + 1), ez = r < 0 ? 0 : (lower ? diagpos[r] : rowptr[r + 1]);
          for (int k = 0; k < S.y; ++k) {
            const int e = ea + k * LPR + q;
            unsigned short li = 0;
            int m = -1;
            if (e < ez) {
              const int cc = colind[e];
              int loc;
              if (cc >= r0 && cc < r1) loc = cc - r0;
              else if (!lower && cc >= n_interior && rl[cc] >= 0) loc = ni + rl[cc];
              else loc = -1;
              if (loc < 0) { loc = 0; m = -2; } // flagged below (cannot throw inside the parallel region)
              else m = e;
              li = (unsigned short)loc;
            }
            ip[k * 32 + l] = li;
            mp[k * 32 + l] = m;
          }
        }
      }
      if (!lower)
        for (int j = 0; j < pt[p].nring; ++j) rl[ring_rows[pt[p].ring0 + j]] = -1;
    }
  }
  for (int m : map)
    if (m == -2) throw StateError("subdomain ILU: an interior row couples outside its part and ring");
  out.n_doubles = n_doubles;
  out.max_local = 0;
  for (int p = 0; p < np; ++p) {
    out.max_local = std::max(out.max_local, part_ptr[p + 1] - part_ptr[p] + pt[p].nring);
    if (fill_parts) { parts[p].row0 = part_ptr[p]; parts[p].ni = part_ptr[p + 1] - part_ptr[p]; }
    if (lower) { parts[p].l_cs0 = pt[p].cs0; parts[p].l_ncol = pt[p].ncol; }
    else { parts[p].u_cs0 = pt[p].cs0; parts[p].u_ncol = pt[p].ncol; parts[p].ring0 = pt[p].ring0; parts[p].nring = pt[p].nring; }
  }
}

static void sd_upload(const HostSdTri &h, bool lower, DevSdTri &out)
{
  out.n_slices = int(h.slices.size());
  out.n_doubles = h.n_doubles;
  out.max_local = h.max_local;
  out.slices.upload(h.slices);
  out.map_off.upload(h.map_off);
  out.cslice.upload(h.cslice);
  out.map.upload(h.map);
  out.stream.upload(h.stream);
  if (!lower) out.ring_rows.upload(h.ring_rows.empty() ? std::vector<int>(1, 0) : h.ring_rows);
}

// Host emulation of k_sd_fill + k_sd_trsv on the packed format (same slices, lanes and local indices as
// the kernels; CPU test of the build without a GPU).  ys: the part's shared-memory vector.
static void sd_emulate_fill(HostSdTri &T, const std::vector<double> &val)
{
  for (size_t s = 0; s < T.slices.size(); ++s) {
    const int4 S = T.slices[s];
    double *vp = T.stream.data() + int64_t(unsigned(S.x)) * 8;
    const int *mp = T.map.data() + T.map_off[s];
    for (int k = 0; k < S.y * 32; ++k) vp[k] = mp[k] >= 0 ? val[mp[k]] : 0.0;
  }
}
static void sd_emulate_parts(const HostSdTri &T, const std::vector<SdPart> &parts, int dir, int bs, int ps,
                             const std::vector<int> &order, const std::vector<double> &dinv, const double *x,
                             std::vector<double> &yp, double *y)
{
  constexpr int LPR = kSdLpr;
  std::vector<double> ys;
  for (const SdPart &P : parts) {
    ys.assign(size_t(P.ni + P.nring) * bs, 0.0);
    for (int l = 0; l < P.ni; ++l)
      for (int d = 0; d < bs; ++d)
        ys[l * bs + d] = dir == 0 ? x[int64_t(bs) * order[P.row0 + l] + d] : yp[int64_t(ps) * (P.row0 + l) + d];
    for (int j = 0; j < P.nring; ++j)
      for (int d = 0; d < bs; ++d) ys[(P.ni + j) * bs + d] = yp[int64_t(ps) * T.ring_rows[P.ring0 + j] + d];
    for (int c = 0; c < P.ncol; ++c)
      for (int s = T.cslice[P.cs0 + c]; s < T.cslice[P.cs0 + c + 1]; ++s) {
        const int4 S = T.slices[s];
        const double *vp = T.stream.data() + int64_t(unsigned(S.x)) * 8;
        const unsigned short *cp = reinterpret_cast<const unsigned short *>(vp + int64_t(S.y) * 32);
        for (int lr = 0; lr < S.w; ++lr) {
          double acc[3] = {0, 0, 0};
          for (int q = 0; q < LPR; ++q)
            for (int k = 0; k < S.y; ++k) {
              const int l = lr * LPR + q;
              for (int d = 0; d < bs; ++d) acc[d] += vp[k * 32 + l] * ys[int(cp[k * 32 + l]) * bs + d];
            }
          double *yr = ys.data() + size_t(S.z + lr) * bs;
          for (int d = 0; d < bs; ++d) yr[d] = dir == 0 ? yr[d] - acc[d] : yr[d] * dinv[P.row0 + S.z + lr] - acc[d];
        }
      }
    for (int l = 0; l < P.ni; ++l)
      for (int d = 0; d < bs; ++d) {
        if (dir == 0) yp[int64_t(ps) * (P.row0 + l) + d] = ys[l * bs + d];
        else y[int64_t(bs) * order[P.row0 + l] + d] = ys[l * bs + d];
      }
  }
}

// CPU check of the whole subdomain-ILU build (ordering -> packed format -> emulated solve) against plain
// substitution on the permuted CSR; returns the largest entry-wise difference relative to max |y|.
double sd_host_check(int n, const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
                     const std::vector<int> &order, const std::vector<int> &part_ptr, const std::vector<int> &pcol_ptr,
                     const std::vector<int> &pcol, const std::vector<int> &sep_colour_ptr, int bs, int *stats)
{
  const int np = int(part_ptr.size()) - 1, n_interior = part_ptr[np], ps = bs == 3 ? 4 : bs;
  std::vector<double> val(colind.size()), dinv(n), x(size_t(n) * bs), y_ref(size_t(n) * bs), y(size_t(n) * bs, 0.0);
  unsigned long long st = 88172645463325252ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return double(st >> 11) / 9007199254740992.0 - 0.5; };
  for (int r = 0; r < n; ++r) {
    const int rl = rowptr[r + 1] - rowptr[r];
    for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) val[e] = rnd() / rl; // diagonally dominant-ish: stable substitution
    dinv[r] = 1.0 + 0.5 * rnd();
  }
  for (auto &v : x) v = rnd();
  // reference: yp = x[order] ; forward ; backward ; y[order] = yp
  std::vector<double> t(size_t(n) * bs);
  for (int r = 0; r < n; ++r)
    for (int d = 0; d < bs; ++d) {
      double a = x[int64_t(bs) * order[r] + d];
      for (int e = rowptr[r]; e < diagpos[r]; ++e) a -= val[e] * t[int64_t(bs) * colind[e] + d];
      t[int64_t(bs) * r + d] = a;
    }
  for (int r = n - 1; r >= 0; --r)
    for (int d = 0; d < bs; ++d) {
      double a = t[int64_t(bs) * r + d] * dinv[r];
      for (int e = diagpos[r] + 1; e < rowptr[r + 1]; ++e) a -= val[e] * t[int64_t(bs) * colind[e] + d];
      t[int64_t(bs) * r + d] = a;
      y_ref[int64_t(bs) * order[r] + d] = a;
    }
  // packed path
  HostSdTri L, U;
  std::vector<SdPart> pl(np), pu;
  sd_pack(rowptr, colind, diagpos, part_ptr, pcol_ptr, pcol, true, n_interior, L, pl, true);
  pu = pl;
  sd_pack(rowptr, colind, diagpos, part_ptr, pcol_ptr, pcol, false, n_interior, U, pu, false);
  for (int p = 0; p < np; ++p) {
    pl[p].cs0 = pl[p].l_cs0; pl[p].ncol = pl[p].l_ncol; pl[p].ring0 = 0; pl[p].nring = 0;
    pu[p].cs0 = pu[p].u_cs0; pu[p].ncol = pu[p].u_ncol;
  }
  sd_emulate_fill(L, val);
  sd_emulate_fill(U, val);
  std::vector<double> yp(size_t(n) * ps, 0.0);
  sd_emulate_parts(L, pl, 0, bs, ps, order, dinv, x.data(), yp, nullptr);
  for (int r = n_interior; r < n; ++r) // separator colours forward (plain CSR: the SELL kernels are not emulated)
    for (int d = 0; d < bs; ++d) {
      double a = x[int64_t(bs) * order[r] + d];
      for (int e = rowptr[r]; e < diagpos[r]; ++e) a -= val[e] * yp[int64_t(ps) * colind[e] + d];
      yp[int64_t(ps) * r + d] = a;
    }
  for (int r = n - 1; r >= n_interior; --r)
    for (int d = 0; d < bs; ++d) {
      double a = yp[int64_t(ps) * r + d] * dinv[r];
      for (int e = diagpos[r] + 1; e < rowptr[r + 1]; ++e) a -= val[e] * yp[int64_t(ps) * colind[e] + d];
      yp[int64_t(ps) * r + d] = a;
      y[int64_t(bs) * order[r] + d] = a;
    }
  sd_emulate_parts(U, pu, 1, bs, ps, order, dinv, x.data(), yp, y.data());
  double err = 0.0, scale = 0.0;
  for (size_t i = 0; i < y.size(); ++i) { err = std::max(err, std::fabs(y[i] - y_ref[i])); scale = std::max(scale, std::fabs(y_ref[i])); }
  // colours must be independent sets (checked on the permuted pattern)
  for (int p = 0; p < np; ++p)
    for (int c = pcol_ptr[p]; c + 1 < pcol_ptr[p + 1]; ++c)
      for (int r = pcol[c]; r < pcol[c + 1]; ++r)
        for (int e = rowptr[r]; e < rowptr[r + 1]; ++e)
          if (colind[e] != r && colind[e] >= pcol[c] && colind[e] < pcol[c + 1]) return 1e30;
  for (size_t c = 0; c + 1 < sep_colour_ptr.size(); ++c)
    for (int r = sep_colour_ptr[c]; r < sep_colour_ptr[c + 1]; ++r)
      for (int e = rowptr[r]; e < rowptr[r + 1]; ++e)
        if (colind[e] != r && colind[e] >= sep_colour_ptr[c] && colind[e] < sep_colour_ptr[c + 1]) return 2e30;
  if (stats) {
    stats[0] = np; stats[1] = n_interior; stats[2] = int(sep_colour_ptr.size()) - 1;
    stats[3] = std::max(L.max_local, U.max_local);
    int mc = 0;
    for (int p = 0; p < np; ++p) mc = std::max(mc, pcol_ptr[p + 1] - pcol_ptr[p] - 1);
    stats[4] = mc;
    int64_t slots = 0, used = 0;
    for (HostSdTri *T : {&L, &U}) {
      slots += int64_t(T->map.size());
      for (int m : T->map) used += m >= 0;
    }
    stats[5] = int(1000.0 * double(used) / double(std::max<int64_t>(slots, 1))); // slot efficiency, per mille
  }
  return scale > 0 ? err / scale : err;
}

static void sd_set_smem_attr(int bs)
{
  const int lim = 227 * 1024;
  if (bs == 3) {
    NSB_CUDA(cudaFuncSetAttribute(k_sd_trsv<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    NSB_CUDA(cudaFuncSetAttribute(k_sd_trsv<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  } else if (bs == 2) {
    NSB_CUDA(cudaFuncSetAttribute(k_sd_trsv<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    NSB_CUDA(cudaFuncSetAttribute(k_sd_trsv<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  } else {
    NSB_CUDA(cudaFuncSetAttribute(k_sd_trsv<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    NSB_CUDA(cudaFuncSetAttribute(k_sd_trsv<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  }
}

void sd_build(DevIlu &ilu, const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
              const std::vector<int> &part_ptr, const std::vector<int> &pcol_ptr, const std::vector<int> &pcol,
              const std::vector<int> &sep_colour_ptr)
{
  const int n = ilu.n, np = int(part_ptr.size()) - 1;
  const int n_interior = part_ptr[np];
  DevSd &sd = ilu.sd;
  sd.n_parts = np;
  sd.n_interior = n_interior;
  std::vector<SdPart> pl(np), pu;
  {
    HostSdTri hl;
    sd_pack(rowptr, colind, diagpos, part_ptr, pcol_ptr, pcol, true, n_interior, hl, pl, true);
    sd_upload(hl, true, sd.L);
  }
  pu = pl;
  {
    HostSdTri hu;
    sd_pack(rowptr, colind, diagpos, part_ptr, pcol_ptr, pcol, false, n_interior, hu, pu, false);
    sd_upload(hu, false, sd.U);
  }
  // one SdPart table per direction (cs0 / ncol are what k_sd_trsv reads)
  std::vector<SdPart> fwd(np), bwd(np);
  for (int p = 0; p < np; ++p) {
    fwd[p] = pl[p]; fwd[p].cs0 = pl[p].l_cs0; fwd[p].ncol = pl[p].l_ncol; fwd[p].ring0 = 0; fwd[p].nring = 0;
    bwd[p] = pu[p]; bwd[p].cs0 = pu[p].u_cs0; bwd[p].ncol = pu[p].u_ncol;
  }
  sd.parts_f.upload(fwd);
  sd.parts_b.upload(bwd);
  // separator rows: split L / U CSR restricted to them, SELL-32 per colour
  std::vector<int> Lp(n + 1, 0), Up(n + 1, 0), Lc, Uc, mapL, mapU;
  for (int k = 0; k < n; ++k) {
    if (k >= n_interior) {
      for (int e = rowptr[k]; e < diagpos[k]; ++e) { Lc.push_back(colind[e]); mapL.push_back(e); }
      for (int e = diagpos[k] + 1; e < rowptr[k + 1]; ++e) { Uc.push_back(colind[e]); mapU.push_back(e); }
    }
    Lp[k + 1] = int(Lc.size());
    Up[k + 1] = int(Uc.size());
  }
  sd.sep_colour_ptr = sep_colour_ptr;
  sell_build(Lp, Lc, mapL, sep_colour_ptr, 4096, 1, ilu.sellL);
  sell_build(Up, Uc, mapU, sep_colour_ptr, 4096, 1, ilu.sellU);
  // dynamic shared memory of the part kernels (set here: ilu_solve launches them inside a stream capture)
  const size_t need = size_t(std::max(sd.L.max_local, sd.U.max_local)) * ilu.bs_rhs * sizeof(double);
  if (need > size_t(227) * 1024) throw StateError("subdomain ILU: a part does not fit shared memory (lower NSB_SD_LEAF)");
  sd_set_smem_attr(ilu.bs_rhs);
  ilu.sdmode = true;
}

void sd_fill(Handle &H, DevIlu &ilu)
{
  DevSd &sd = ilu.sd;
  for (DevSdTri *T : {&sd.L, &sd.U}) {
    if (T->n_slices == 0) continue;
    k_sd_fill<<<unsigned(std::min((T->n_slices * 32 + 255) / 256, kSM_sd * 16)), 256, 0, H.stream>>>(
        T->n_slices, T->slices.p, T->map_off.p, T->map.p, ilu.val.p, T->stream.p);
    H.launches++;
  }
  sell_fill(H, ilu.sellL, ilu.val.p);
  sell_fill(H, ilu.sellU, ilu.val.p);
  NSB_CUDA(cudaGetLastError());
}

int sd_stride(int bs_rhs) { return bs_rhs == 3 ? 4 : bs_rhs; }

template <int BS>
static void sd_trsv_t(Handle &H, DevIlu &ilu, double *yp, cudaStream_t s)
{
  DevSd &sd = ilu.sd;
  const TrsvIoSd *io = reinterpret_cast<const TrsvIoSd *>(ilu.io.p);
  const size_t smem_f = size_t(sd.L.max_local) * BS * sizeof(double), smem_b = size_t(sd.U.max_local) * BS * sizeof(double);
  const int nsc = int(sd.sep_colour_ptr.size()) - 1;
  if (sd.n_parts > 0) {
    k_sd_trsv<BS, 0><<<sd.n_parts, kSdThreads, smem_f, s>>>(sd.parts_f.p, sd.L.slices.p, sd.L.cslice.p, sd.L.stream.p, nullptr, yp,
                                                          ilu.dinv.p, ilu.order.p, io);
    H.launches++;
  }
  for (int c = 0; c < nsc; ++c) {
    const int a = ilu.sellL.range_slice[c], b = ilu.sellL.range_slice[c + 1];
    if (b <= a) continue;
    k_sd_sep<BS, 0><<<unsigned((b - a + 7) / 8), 256, 0, s>>>(a, b, ilu.sellL.slice_ptr.p, ilu.sellL.rowid.p, ilu.sellL.col.p,
                                                             ilu.sellL.val.p, yp, ilu.dinv.p, ilu.order.p, io);
    H.launches++;
  }
  for (int c = nsc - 1; c >= 0; --c) {
    const int a = ilu.sellU.range_slice[c], b = ilu.sellU.range_slice[c + 1];
    if (b <= a) continue;
    k_sd_sep<BS, 1><<<unsigned((b - a + 7) / 8), 256, 0, s>>>(a, b, ilu.sellU.slice_ptr.p, ilu.sellU.rowid.p, ilu.sellU.col.p,
                                                             ilu.sellU.val.p, yp, ilu.dinv.p, ilu.order.p, io);
    H.launches++;
  }
  if (sd.n_parts > 0) {
    k_sd_trsv<BS, 1><<<sd.n_parts, kSdThreads, smem_b, s>>>(sd.parts_b.p, sd.U.slices.p, sd.U.cslice.p, sd.U.stream.p,
                                                          sd.U.ring_rows.p, yp, ilu.dinv.p, ilu.order.p, io);
    H.launches++;
  }
}

// yp: staging vector in factor order (sd_stride doubles per row); in / out through ilu.io
void sd_trsv(Handle &H, DevIlu &ilu, double *yp, cudaStream_t s)
{
  if (ilu.bs_rhs == 3) sd_trsv_t<3>(H, ilu, yp, s);
  else if (ilu.bs_rhs == 2) sd_trsv_t<2>(H, ilu, yp, s);
  else sd_trsv_t<1>(H, ilu, yp, s);
  NSB_CUDA(cudaGetLastError());
}

} // namespace nsb
