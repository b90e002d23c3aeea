// kernels_sd.cu -- subdomain-resident ILU(0) triangular solves (ilu_ordering = 3).
//
// Replaces TrilinosWrappers::PreconditionILU::vmult = U^-1 D^-1 L^-1 (Ifpack_ILU::ApplyInverse,
// include/Preconditioners.hpp:319-320) for the factors in the multi-level "subdomain" ordering built by
// subdomain_order (kernels_linalg.cu): [parts of level 1 | parts of level 2 | ... | remaining rows].
//
// Part rows (k_sd_trsv): ONE CTA solves ONE part start to finish.
//   * The part's slice of the vector and the rows of other levels it couples with (its "ring") are
//     staged into shared memory once; every gather of the sweep is a shared-memory access.
//   * The part's factor entries are packed in processing order -- FP64 value + 16-bit part-local column,
//     10 bytes per entry -- as a sequence of ROUNDS of <= 24 KB.  One elected thread streams the rounds
//     into a 3-stage shared-memory ring with bulk asynchronous copies (cp.async.bulk, the TMA engine)
//     that complete on mbarriers; the round table itself arrives the same way.  The sweep therefore
//     never waits on an HBM round trip between colours: up to 72 KB per SM are in flight while the
//     warps work out of shared memory, and every factor entry is read from HBM exactly once.
//   * Colours inside a part are separated by __syncthreads() instead of kernel boundaries.
// Against global colour sweeps (k_sell3) this removes the re-gathering of the vector from HBM (2.6 times
// the algorithmic bytes, profiles/r01_traffic.json) and turns ~30 dependent launches into one per level.
// Remaining rows (k_sd_sep): SELL-32 colour sweeps over the global staging vector, on a few % of the rows.
//
// Round block: [ n_slices x int4 header ][ slice blocks ].   header: x = offset of the slice block in the
// round block (bytes), y = steps, z = rows, w = offset of the values inside the slice block.
// Slice block (<= 8 rows, 4 lanes per row: entry e of a row sits in lane e % 4 of step e / 4):
//   [ 8 x uint16 part-local row ][ backward only: 8 x FP64 inverse diagonal ][ steps x 32 FP64 values ]
//   [ steps x 32 uint16 part-local columns ]
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "nsb_internal.hpp"

namespace nsb {

constexpr int kSdThreads = 512;            // 16 warps, one CTA per SM
constexpr int kSdLpr = 4;                  // lanes per row
constexpr int kSdRps = 32 / kSdLpr;        // rows per slice
constexpr int kSdStages = 3;               // ring stages
constexpr int kSdRoundBytes = 24 * 1024;   // capacity of a stage
constexpr int kSdMaxSlices = 32;           // slices per round: two per warp
constexpr int kSdBarBytes = 64;            // mbarriers at the start of shared memory
constexpr int kSM_sd = 148;

struct TrsvIoSd { const double *x; double *y; }; // same slot as TrsvIo in kernels_sell.cu (set by k_set_io)

__host__ __device__ inline size_t sd_smem_bytes(int max_rounds, int max_local, int bs)
{
  return size_t(kSdBarBytes) + size_t((max_rounds + 3) / 4 * 4) * 16 + size_t(kSdStages) * kSdRoundBytes +
         size_t(max_local) * bs * sizeof(double);
}

// ---- mbarrier / bulk-copy primitives (PTX; sm_90+)
__device__ __forceinline__ unsigned sd_smem_u32(const void *p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void sd_mbar_init(unsigned long long *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sd_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sd_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sd_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool sd_mbar_try_wait(unsigned long long *bar, unsigned parity)
{
  unsigned ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
               : "=r"(ok)
               : "r"(sd_smem_u32(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}
// a bulk copy that never lands (bad descriptor) must not hang the GPU: trap after ~seconds
__device__ __forceinline__ void sd_mbar_wait(unsigned long long *bar, unsigned parity)
{
  if (sd_mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!sd_mbar_try_wait(bar, parity))
    if (clock64() - t0 > 8000000000LL) __trap();
}
__device__ __forceinline__ void sd_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sd_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(sd_smem_u32(bar))
               : "memory");
}

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
// backward kernel (writes the caller's y; the staging vector is kept current for the rings of the
// levels that follow).
template <int BS, int DIR>
__global__ void __launch_bounds__(kSdThreads, 1) k_sd_trsv(const SdPart *__restrict__ parts, const int4 *__restrict__ rounds,
                                                           const double *__restrict__ stream, const int *__restrict__ ring_rows,
                                                           double *yp, const int *__restrict__ order,
                                                           const TrsvIoSd *__restrict__ io, int max_rounds)
{
  constexpr int PS = BS == 3 ? 4 : BS;
  constexpr int LPR = kSdLpr, NW = kSdThreads / 32, NST = kSdStages;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char sd_smem[];
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(sd_smem); // [0, NST): stage full; [NST]: round table
  int4 *rtab = reinterpret_cast<int4 *>(sd_smem + kSdBarBytes);
  unsigned char *ring = sd_smem + kSdBarBytes + size_t((max_rounds + 3) / 4 * 4) * 16;
  double *ys = reinterpret_cast<double *>(ring + size_t(NST) * kSdRoundBytes); // [(ni + nring)][BS]
  const SdPart P = parts[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ni = P.ni, nr = P.nrounds;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s <= NST; ++s) sd_mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0 && nr > 0) { // the part's round table, by bulk copy
    sd_mbar_expect_tx(&bars[NST], unsigned(nr) * 16u);
    sd_bulk_g2s(rtab, rounds + P.round0, unsigned(nr) * 16u, &bars[NST]);
  }
  // ---- stage the vector in
  if (DIR == 0) {
    const double *x = io->x;
    for (int l = tid; l < ni; l += kSdThreads) {
      const double *xi = x + int64_t(BS) * order[P.row0 + l];
#pragma unroll
      for (int d = 0; d < BS; ++d) ys[l * BS + d] = xi[d];
    }
  } else {
    for (int t = tid; t < ni * BS; t += kSdThreads) {
      const int l = t / BS, d = t - l * BS;
      ys[t] = yp[int64_t(PS) * (P.row0 + l) + d];
    }
  }
  // ring: rows of other levels this part couples with (final in the staging vector when this level runs)
  for (int j = tid; j < P.nring; j += kSdThreads) {
    double v[BS];
    sd_gather_global<BS>(yp, ring_rows[P.ring0 + j], v);
#pragma unroll
    for (int d = 0; d < BS; ++d) ys[(ni + j) * BS + d] = v[d];
  }
  auto issue = [&](int q) { // thread 0: start the bulk copy of round q into its stage
    const int4 R = rtab[q];
    unsigned long long *bar = &bars[q % NST];
    sd_mbar_expect_tx(bar, unsigned(R.y));
    sd_bulk_g2s(ring + size_t(q % NST) * kSdRoundBytes, reinterpret_cast<const unsigned char *>(stream) + size_t(unsigned(R.x)) * 16,
                unsigned(R.y), bar);
  };
  if (nr > 0) sd_mbar_wait(&bars[NST], 0); // every thread reads the table below
  if (tid == 0)
    for (int q = 0; q < NST && q < nr; ++q) issue(q);
  __syncthreads(); // vector staged
  // ---- rounds, in processing order (a round never mixes colours)
  for (int r = 0; r < nr; ++r) {
    const int st = r % NST;
    sd_mbar_wait(&bars[st], unsigned(r / NST) & 1u);
    const unsigned char *blk = ring + size_t(st) * kSdRoundBytes;
    const int nsl = rtab[r].z;
    for (int s = warp; s < nsl; s += NW) {
      const int4 S = reinterpret_cast<const int4 *>(blk)[s];
      const unsigned char *sb = blk + S.x;
      const int len = S.y;
      const double *vp = reinterpret_cast<const double *>(sb + S.w) + lane;
      const unsigned short *cp = reinterpret_cast<const unsigned short *>(sb + S.w + size_t(len) * 256) + lane;
      double acc[BS];
#pragma unroll
      for (int d = 0; d < BS; ++d) acc[d] = 0.0;
#pragma unroll 4
      for (int k = 0; k < len; ++k) {
        const double v = vp[k * 32];
        const double *yb = ys + int(cp[k * 32]) * BS;
#pragma unroll
        for (int d = 0; d < BS; ++d) acc[d] += v * yb[d];
      }
#pragma unroll
      for (int o = 1; o < LPR; o <<= 1)
#pragma unroll
        for (int d = 0; d < BS; ++d) acc[d] += __shfl_xor_sync(FULL, acc[d], o);
      const int lr = lane / LPR;
      if ((lane % LPR) == 0 && lr < S.z) {
        const int row = int(reinterpret_cast<const unsigned short *>(sb)[lr]);
        double *yr = ys + row * BS;
        if (DIR == 0) {
#pragma unroll
          for (int d = 0; d < BS; ++d) yr[d] -= acc[d];
        } else {
          const double di = reinterpret_cast<const double *>(sb + 16)[lr];
#pragma unroll
          for (int d = 0; d < BS; ++d) yr[d] = yr[d] * di - acc[d];
        }
      }
    }
    __syncthreads(); // the round's rows are final; its stage is free
    if (tid == 0 && r + NST < nr) issue(r + NST);
  }
  // ---- stage the vector out
  for (int t = tid; t < ni * PS; t += kSdThreads) {
    const int l = t / PS, d = t - l * PS;
    yp[int64_t(PS) * P.row0 + t] = d < BS ? ys[l * BS + d] : 0.0;
  }
  if (DIR == 1) {
    double *y = io->y;
    for (int l = tid; l < ni; l += kSdThreads) {
      double *yo = y + int64_t(BS) * order[P.row0 + l];
#pragma unroll
      for (int d = 0; d < BS; ++d) yo[d] = ys[l * BS + d];
    }
  }
}

// Remaining rows: one colour per launch, SELL-32 (one thread per row) over the global staging vector.
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
__global__ void k_sd_fill(int n_slices, const int2 *__restrict__ slices, const int64_t *__restrict__ map_off,
                          const int *__restrict__ map, const double *__restrict__ src, double *__restrict__ stream)
{
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slices; s += warps) {
    const int2 S = slices[s];
    double *vp = stream + int64_t(unsigned(S.x));
    const int *mp = map + map_off[s];
    for (int k = 0; k < S.y; ++k) {
      const int m = mp[k * 32 + lane];
      vp[k * 32 + lane] = m >= 0 ? src[m] : 0.0;
    }
  }
}
__global__ void k_sd_fill_dinv(int64_t n, const int2 *__restrict__ dfill, const double *__restrict__ dinv, double *__restrict__ stream)
{
  for (int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; k < n; k += int64_t(gridDim.x) * blockDim.x)
    stream[unsigned(dfill[k].x)] = dinv[dfill[k].y];
}

// ------------------------------------------------------------------------------------------------
// host side: pack the part rows of one triangular factor
// rowptr / colind / diagpos: the factor pattern in factor order (diagonal inside).
struct HostSdTri {
  std::vector<int4> rounds;
  std::vector<int2> fill_slices, dfill;
  std::vector<int64_t> map_off;
  std::vector<int> map, ring_rows;
  std::vector<double> stream;
  int64_t n_doubles = 0;
};

static void sd_pack(const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
                    const std::vector<int> &part_ptr, const std::vector<int> &pcol_ptr, const std::vector<int> &pcol,
                    const std::vector<int> &level_part_ptr, bool lower, HostSdTri &out, std::vector<SdPart> &parts)
{
  constexpr int LPR = kSdLpr, RPS = kSdRps;
  const int np = int(part_ptr.size()) - 1;
  const int head = lower ? 16 : 16 + 8 * RPS; // rows (+ inverse diagonals) before the values of a slice block
  auto row_len = [&](int r) { return lower ? diagpos[r] - rowptr[r] : rowptr[r + 1] - diagpos[r] - 1; };
  std::vector<int> lvl_lo(np, 0), lvl_hi(np, 0); // factor rows of the level each part belongs to
  for (size_t l = 0; l + 1 < level_part_ptr.size(); ++l)
    for (int p = level_part_ptr[l]; p < level_part_ptr[l + 1]; ++p) {
      lvl_lo[p] = part_ptr[level_part_ptr[l]];
      lvl_hi[p] = part_ptr[level_part_ptr[l + 1]];
    }
  // ---- pass 1 (sequential): rings, slices, rounds, sizes
  struct SliceTmp { int rows[RPS]; int nr, len; int64_t block_off; int64_t map_off; }; // block_off: bytes in the stream
  std::vector<SliceTmp> slices;
  std::vector<int> part_slice0(np + 1, 0);
  out.rounds.clear(); out.ring_rows.clear();
  std::vector<int> ring_local(rowptr.size() - 1, -1), srt;
  int64_t n_bytes = 0, n_slots = 0;
  for (int p = 0; p < np; ++p) {
    const int r0 = part_ptr[p], r1 = part_ptr[p + 1];
    const int ncol = pcol_ptr[p + 1] - pcol_ptr[p] - 1;
    parts[p].row0 = r0; parts[p].ni = r1 - r0;
    parts[p].round0 = int(out.rounds.size());
    parts[p].ring0 = int(out.ring_rows.size());
    part_slice0[p] = int(slices.size());
    // ring: rows of other levels the part's entries couple with (L: earlier levels, U: later levels and the
    // remaining rows), in first-use order
    for (int r = r0; r < r1; ++r) {
      const int ea = lower ? rowptr[r] : diagpos[r] + 1, ez = lower ? diagpos[r] : rowptr[r + 1];
      for (int e = ea; e < ez; ++e) {
        const int cc = colind[e];
        if (cc >= r0 && cc < r1) continue;
        if (cc >= lvl_lo[p] && cc < lvl_hi[p]) throw StateError("subdomain ILU: a row couples with another part of its level");
        if (ring_local[cc] < 0) { ring_local[cc] = int(out.ring_rows.size()) - parts[p].ring0; out.ring_rows.push_back(cc); }
      }
    }
    parts[p].nring = int(out.ring_rows.size()) - parts[p].ring0;
    for (int j = parts[p].ring0; j < int(out.ring_rows.size()); ++j) ring_local[out.ring_rows[j]] = -1;
    if ((r1 - r0) + parts[p].nring > 65535) throw StateError("subdomain ILU: part too large for 16-bit local indices");
    for (int ci = 0; ci < ncol; ++ci) {
      const int c = lower ? ci : ncol - 1 - ci; // backward solve: colours in reverse
      const int a = pcol[pcol_ptr[p] + c], b = pcol[pcol_ptr[p] + c + 1];
      srt.resize(b - a);
      std::iota(srt.begin(), srt.end(), a);
      std::stable_sort(srt.begin(), srt.end(), [&](int x, int y) { return row_len(x) > row_len(y); });
      // rounds of this colour: <= kSdMaxSlices slices and <= kSdRoundBytes bytes each
      int round_nsl = 0;
      int64_t round_bytes = 0, round_start = n_bytes;
      std::vector<size_t> round_slices;
      auto close_round = [&]() {
        if (round_nsl == 0) return;
        // the headers come first: shift the slice blocks behind them
        const int64_t hb = int64_t(round_nsl) * 16;
        for (size_t si : round_slices) slices[si].block_off += hb;
        if (round_start / 16 > int64_t(0xffffffffu)) throw StateError("subdomain ILU: stream too large");
        out.rounds.push_back(make_int4(int(unsigned(round_start / 16)), int(hb + round_bytes), round_nsl, c));
        n_bytes = round_start + hb + round_bytes;
        round_nsl = 0; round_bytes = 0; round_start = n_bytes; round_slices.clear();
      };
      for (int s = 0; s < b - a; s += RPS) {
        SliceTmp T;
        T.nr = std::min(RPS, b - a - s);
        T.len = 0;
        for (int k = 0; k < RPS; ++k) T.rows[k] = k < T.nr ? srt[s + k] : -1;
        for (int k = 0; k < T.nr; ++k) T.len = std::max(T.len, (row_len(T.rows[k]) + LPR - 1) / LPR);
        if (T.len == 0 && lower) break; // sorted by length: nothing left to subtract in this colour (backward rows are still scaled)
        const int64_t bytes = head + int64_t(T.len) * 320;
        if (16 + bytes > kSdRoundBytes) throw StateError("subdomain ILU: a slice does not fit a round");
        if (round_nsl == kSdMaxSlices || int64_t(round_nsl + 1) * 16 + round_bytes + bytes > kSdRoundBytes) close_round();
        T.block_off = round_start + round_bytes; // headers added in close_round
        T.map_off = n_slots;
        round_bytes += bytes;
        n_slots += int64_t(T.len) * 32;
        round_slices.push_back(slices.size());
        slices.push_back(T);
        ++round_nsl;
      }
      close_round();
    }
    parts[p].nrounds = int(out.rounds.size()) - parts[p].round0;
  }
  part_slice0[np] = int(slices.size());
  if (n_bytes % 16) throw StateError("subdomain ILU: misaligned stream");
  // ---- pass 2 (parallel over parts): headers, rows, columns, fill maps
  out.n_doubles = n_bytes / 8;
  out.stream.assign(size_t(out.n_doubles) + 2, 0.0);
  out.map.assign(size_t(n_slots), -1);
  out.fill_slices.resize(slices.size());
  out.map_off.resize(slices.size());
  unsigned char *sbytes = reinterpret_cast<unsigned char *>(out.stream.data());
  // headers: walk the rounds (slices are stored in round order)
  {
    size_t si = 0;
    for (const int4 &R : out.rounds) {
      const int64_t rb = int64_t(unsigned(R.x)) * 16;
      int4 *hd = reinterpret_cast<int4 *>(sbytes + rb);
      for (int k = 0; k < R.z; ++k, ++si) hd[k] = make_int4(int(slices[si].block_off - rb), slices[si].len, slices[si].nr, head);
    }
    if (si != slices.size()) throw StateError("subdomain ILU: round / slice bookkeeping");
  }
  bool bad = false;
#pragma omp parallel
  {
    std::vector<int> rl(rowptr.size() - 1, -1); // per-thread ring lookup (factor row -> ring position), sparse reset
#pragma omp for schedule(dynamic, 8)
    for (int p = 0; p < np; ++p) {
      const int r0 = part_ptr[p], r1 = part_ptr[p + 1], ni = r1 - r0;
      for (int j = 0; j < parts[p].nring; ++j) rl[out.ring_rows[parts[p].ring0 + j]] = j;
      for (int s = part_slice0[p]; s < part_slice0[p + 1]; ++s) {
        const SliceTmp &T = slices[s];
        unsigned char *sb = sbytes + T.block_off;
        unsigned short *rows = reinterpret_cast<unsigned short *>(sb);
        for (int k = 0; k < RPS; ++k) rows[k] = (unsigned short)(T.rows[k] >= 0 ? T.rows[k] - r0 : 0);
        unsigned short *ip = reinterpret_cast<unsigned short *>(sb + head + size_t(T.len) * 256);
        int *mp = out.map.data() + T.map_off;
        out.fill_slices[s] = make_int2(int(unsigned((T.block_off + head) / 8)), T.len);
        out.map_off[s] = T.map_off;
        for (int l = 0; l < 32; ++l) {
          const int lr = l / LPR, q = l % LPR;
          const int r = T.rows[lr];
          const int ea = r < 0 ? 0 : (lower ? rowptr[r] : diagpos[r] + 1), ez = r < 0 ? 0 : (lower ? diagpos[r] : rowptr[r + 1]);
          for (int k = 0; k < T.len; ++k) {
            const int e = ea + k * LPR + q;
            unsigned short li = 0;
            int m = -1;
            if (e < ez) {
              const int cc = colind[e];
              int loc = -1;
              if (cc >= r0 && cc < r1) loc = cc - r0;
              else if (rl[cc] >= 0) loc = ni + rl[cc];
              if (loc < 0) { loc = 0; bad = true; } // cannot throw inside the parallel region
              else m = e;
              li = (unsigned short)loc;
            }
            ip[k * 32 + l] = li;
            mp[k * 32 + l] = m;
          }
        }
      }
      for (int j = 0; j < parts[p].nring; ++j) rl[out.ring_rows[parts[p].ring0 + j]] = -1;
    }
  }
  if (bad) throw StateError("subdomain ILU: a row couples outside its part and ring");
  out.dfill.clear();
  if (!lower)
    for (const SliceTmp &T : slices)
      for (int k = 0; k < T.nr; ++k) out.dfill.push_back(make_int2(int(unsigned((T.block_off + 16) / 8 + k)), T.rows[k]));
}

static void sd_upload(const HostSdTri &h, DevSdTri &out)
{
  out.n_rounds = int(h.rounds.size());
  out.n_slices = int(h.fill_slices.size());
  out.n_doubles = h.n_doubles;
  out.rounds.upload(h.rounds.empty() ? std::vector<int4>(1, make_int4(0, 0, 0, 0)) : h.rounds);
  out.fill_slices.upload(h.fill_slices);
  out.map_off.upload(h.map_off);
  out.map.upload(h.map);
  out.stream.upload(h.stream);
  out.dfill.upload(h.dfill);
  out.ring_rows.upload(h.ring_rows.empty() ? std::vector<int>(1, 0) : h.ring_rows);
}

// Host emulation of k_sd_fill + k_sd_trsv on the packed format (same rounds, slices, lanes and local
// indices as the kernels; CPU test of the build without a GPU).  ys: the part's shared-memory vector.
static void sd_emulate_fill(HostSdTri &T, const std::vector<double> &val, const std::vector<double> &dinv)
{
  for (size_t s = 0; s < T.fill_slices.size(); ++s) {
    double *vp = T.stream.data() + unsigned(T.fill_slices[s].x);
    const int *mp = T.map.data() + T.map_off[s];
    for (int k = 0; k < T.fill_slices[s].y * 32; ++k) vp[k] = mp[k] >= 0 ? val[mp[k]] : 0.0;
  }
  for (const int2 &d : T.dfill) T.stream[unsigned(d.x)] = dinv[d.y];
}
static void sd_emulate_parts(const HostSdTri &T, const SdPart *parts, int nparts, int dir, int bs, int ps,
                             const std::vector<int> &order, const double *x, std::vector<double> &yp, double *y)
{
  constexpr int LPR = kSdLpr;
  const unsigned char *sbytes = reinterpret_cast<const unsigned char *>(T.stream.data());
  std::vector<double> ys;
  for (int pi = 0; pi < nparts; ++pi) {
    const SdPart &P = parts[pi];
    ys.assign(size_t(P.ni + P.nring) * bs, 0.0);
    for (int l = 0; l < P.ni; ++l)
      for (int d = 0; d < bs; ++d)
        ys[l * bs + d] = dir == 0 ? x[int64_t(bs) * order[P.row0 + l] + d] : yp[int64_t(ps) * (P.row0 + l) + d];
    for (int j = 0; j < P.nring; ++j)
      for (int d = 0; d < bs; ++d) ys[(P.ni + j) * bs + d] = yp[int64_t(ps) * T.ring_rows[P.ring0 + j] + d];
    for (int r = 0; r < P.nrounds; ++r) {
      const int4 R = T.rounds[P.round0 + r];
      const unsigned char *blk = sbytes + int64_t(unsigned(R.x)) * 16;
      for (int s = 0; s < R.z; ++s) {
        const int4 S = reinterpret_cast<const int4 *>(blk)[s];
        const unsigned char *sb = blk + S.x;
        const double *vp = reinterpret_cast<const double *>(sb + S.w);
        const unsigned short *cp = reinterpret_cast<const unsigned short *>(sb + S.w + size_t(S.y) * 256);
        for (int lr = 0; lr < S.z; ++lr) {
          double acc[3] = {0, 0, 0};
          for (int q = 0; q < LPR; ++q)
            for (int k = 0; k < S.y; ++k) {
              const int l = lr * LPR + q;
              for (int d = 0; d < bs; ++d) acc[d] += vp[k * 32 + l] * ys[int(cp[k * 32 + l]) * bs + d];
            }
          const int row = int(reinterpret_cast<const unsigned short *>(sb)[lr]);
          double *yr = ys.data() + size_t(row) * bs;
          const double di = dir == 1 ? reinterpret_cast<const double *>(sb + 16)[lr] : 0.0;
          for (int d = 0; d < bs; ++d) yr[d] = dir == 0 ? yr[d] - acc[d] : yr[d] * di - acc[d];
        }
      }
    }
    for (int l = 0; l < P.ni; ++l)
      for (int d = 0; d < bs; ++d) {
        yp[int64_t(ps) * (P.row0 + l) + d] = ys[l * bs + d];
        if (dir == 1) y[int64_t(bs) * order[P.row0 + l] + d] = ys[l * bs + d];
      }
  }
}

// CPU check of the whole subdomain-ILU build (ordering -> packed format -> emulated solve) against plain
// substitution on the permuted CSR; returns the largest entry-wise difference relative to max |y|.
double sd_host_check(int n, const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
                     const std::vector<int> &order, const std::vector<int> &part_ptr, const std::vector<int> &pcol_ptr,
                     const std::vector<int> &pcol, const std::vector<int> &sep_colour_ptr, const std::vector<int> &level_part_ptr,
                     int bs, int *stats)
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
  std::vector<SdPart> pl(np), pu(np);
  sd_pack(rowptr, colind, diagpos, part_ptr, pcol_ptr, pcol, level_part_ptr, true, L, pl);
  sd_pack(rowptr, colind, diagpos, part_ptr, pcol_ptr, pcol, level_part_ptr, false, U, pu);
  sd_emulate_fill(L, val, dinv);
  sd_emulate_fill(U, val, dinv);
  std::vector<double> yp(size_t(n) * ps, 0.0);
  const int nl = int(level_part_ptr.size()) - 1;
  for (int l = 0; l < nl; ++l)
    sd_emulate_parts(L, pl.data() + level_part_ptr[l], level_part_ptr[l + 1] - level_part_ptr[l], 0, bs, ps, order, x.data(), yp, nullptr);
  for (int r = n_interior; r < n; ++r) // remaining rows forward (plain CSR: the SELL kernels are not emulated)
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
  for (int l = nl - 1; l >= 0; --l)
    sd_emulate_parts(U, pu.data() + level_part_ptr[l], level_part_ptr[l + 1] - level_part_ptr[l], 1, bs, ps, order, x.data(), yp, y.data());
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
    int ml = 0, mr = 0;
    for (int p = 0; p < np; ++p) {
      ml = std::max(ml, std::max(pl[p].ni + pl[p].nring, pu[p].ni + pu[p].nring));
      mr = std::max(mr, std::max(pl[p].nrounds, pu[p].nrounds));
    }
    stats[3] = ml;
    stats[4] = mr;
    int64_t slots = 0, used = 0;
    for (HostSdTri *T : {&L, &U}) {
      slots += int64_t(T->map.size());
      for (int m : T->map) used += m >= 0;
    }
    stats[5] = int(1000.0 * double(used) / double(std::max<int64_t>(slots, 1))); // slot efficiency, per mille
    stats[6] = nl; stats[7] = nl > 0 ? part_ptr[level_part_ptr[1]] : 0;           // levels, rows of level 1
  }
  return scale > 0 ? err / scale : err;
}

// shared memory the part kernels would need for this ordering (largest rows + ring over all parts and
// both directions; the round table is bounded by 4 KB here), so that ilu_build can shrink the parts
size_t sd_smem_needed(const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
                      const std::vector<int> &part_ptr, int bs)
{
  const int np = int(part_ptr.size()) - 1;
  int max_local = 0;
#pragma omp parallel
  {
    std::vector<int> seen(rowptr.size() - 1, -1);
#pragma omp for schedule(dynamic, 16) reduction(max : max_local)
    for (int p = 0; p < np; ++p) {
      const int r0 = part_ptr[p], r1 = part_ptr[p + 1];
      int lo = 0, hi = 0;
      for (int r = r0; r < r1; ++r)
        for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) {
          const int cc = colind[e];
          if ((cc >= r0 && cc < r1) || seen[cc] == p) continue;
          seen[cc] = p;
          (e < diagpos[r] ? lo : hi)++;
        }
      max_local = std::max(max_local, (r1 - r0) + std::max(lo, hi));
    }
  }
  return sd_smem_bytes(256, max_local, bs);
}

static void sd_set_smem_attr(int bs)
{
  const int lim = 227 * 1024;
  if (bs == 3) {
    NSB_CUDA_SETUP(cudaFuncSetAttribute(k_sd_trsv<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    NSB_CUDA_SETUP(cudaFuncSetAttribute(k_sd_trsv<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  } else if (bs == 2) {
    NSB_CUDA_SETUP(cudaFuncSetAttribute(k_sd_trsv<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    NSB_CUDA_SETUP(cudaFuncSetAttribute(k_sd_trsv<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  } else {
    NSB_CUDA_SETUP(cudaFuncSetAttribute(k_sd_trsv<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    NSB_CUDA_SETUP(cudaFuncSetAttribute(k_sd_trsv<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  }
}

void sd_build(DevIlu &ilu, const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
              const std::vector<int> &part_ptr, const std::vector<int> &pcol_ptr, const std::vector<int> &pcol,
              const std::vector<int> &sep_colour_ptr, const std::vector<int> &level_part_ptr)
{
  const int n = ilu.n, np = int(part_ptr.size()) - 1;
  const int n_interior = part_ptr[np];
  DevSd &sd = ilu.sd;
  sd.n_parts = np;
  sd.n_interior = n_interior;
  sd.level_part_ptr = level_part_ptr;
  std::vector<SdPart> pl(np), pu(np);
  {
    HostSdTri hl;
    sd_pack(rowptr, colind, diagpos, part_ptr, pcol_ptr, pcol, level_part_ptr, true, hl, pl);
    sd_upload(hl, sd.L);
  }
  {
    HostSdTri hu;
    sd_pack(rowptr, colind, diagpos, part_ptr, pcol_ptr, pcol, level_part_ptr, false, hu, pu);
    sd_upload(hu, sd.U);
  }
  sd.parts_f.upload(pl);
  sd.parts_b.upload(pu);
  // shared memory per level and direction: the largest (rows + ring) and the most rounds of its parts
  const int nl = int(level_part_ptr.size()) - 1;
  sd.level_local_f.assign(nl, 0); sd.level_local_b.assign(nl, 0);
  sd.level_rounds_f.assign(nl, 0); sd.level_rounds_b.assign(nl, 0);
  size_t need = 0;
  for (int l = 0; l < nl; ++l) {
    for (int p = level_part_ptr[l]; p < level_part_ptr[l + 1]; ++p) {
      sd.level_local_f[l] = std::max(sd.level_local_f[l], pl[p].ni + pl[p].nring);
      sd.level_local_b[l] = std::max(sd.level_local_b[l], pu[p].ni + pu[p].nring);
      sd.level_rounds_f[l] = std::max(sd.level_rounds_f[l], pl[p].nrounds);
      sd.level_rounds_b[l] = std::max(sd.level_rounds_b[l], pu[p].nrounds);
    }
    need = std::max(need, std::max(sd_smem_bytes(sd.level_rounds_f[l], sd.level_local_f[l], ilu.bs_rhs),
                                   sd_smem_bytes(sd.level_rounds_b[l], sd.level_local_b[l], ilu.bs_rhs)));
  }
  if (need > size_t(227) * 1024) throw StateError("subdomain ILU: a part does not fit shared memory (lower NSB_SD_LEAF)");
  // remaining rows: split L / U CSR restricted to them, SELL-32 per colour
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
  sd_set_smem_attr(ilu.bs_rhs); // set here: ilu_solve launches the kernels inside a stream capture
  ilu.sdmode = true;
}

void sd_fill(Handle &H, DevIlu &ilu)
{
  DevSd &sd = ilu.sd;
  for (DevSdTri *T : {&sd.L, &sd.U}) {
    if (T->n_slices > 0) {
      k_sd_fill<<<unsigned(std::min((T->n_slices * 32 + 255) / 256, kSM_sd * 16)), 256, 0, H.stream>>>(
          T->n_slices, T->fill_slices.p, T->map_off.p, T->map.p, ilu.val.p, T->stream.p);
      H.launches++;
    }
    if (T->dfill.n > 0) {
      k_sd_fill_dinv<<<unsigned(std::min<int64_t>((int64_t(T->dfill.n) + 255) / 256, kSM_sd * 16)), 256, 0, H.stream>>>(
          int64_t(T->dfill.n), T->dfill.p, ilu.dinv.p, T->stream.p);
      H.launches++;
    }
  }
  sell_fill(H, ilu.sellL, ilu.val.p);
  sell_fill(H, ilu.sellU, ilu.val.p);
  NSB_CUDA(cudaGetLastError());
}

int sd_stride(int bs_rhs) { return bs_rhs == 3 ? 4 : bs_rhs; }
int sd_launches(const DevIlu &ilu) { return 2 * (int(ilu.sd.level_part_ptr.size()) - 1) + 2 * (int(ilu.sd.sep_colour_ptr.size()) - 1); }

template <int BS>
static void sd_trsv_t(Handle &H, DevIlu &ilu, double *yp, cudaStream_t s)
{
  DevSd &sd = ilu.sd;
  const TrsvIoSd *io = reinterpret_cast<const TrsvIoSd *>(ilu.io.p);
  const int nsc = int(sd.sep_colour_ptr.size()) - 1, nl = int(sd.level_part_ptr.size()) - 1;
  for (int l = 0; l < nl; ++l) {
    const int p0 = sd.level_part_ptr[l], cnt = sd.level_part_ptr[l + 1] - p0;
    if (cnt <= 0) continue;
    k_sd_trsv<BS, 0><<<cnt, kSdThreads, sd_smem_bytes(sd.level_rounds_f[l], sd.level_local_f[l], BS), s>>>(
        sd.parts_f.p + p0, sd.L.rounds.p, sd.L.stream.p, sd.L.ring_rows.p, yp, ilu.order.p, io, sd.level_rounds_f[l]);
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
  for (int l = nl - 1; l >= 0; --l) {
    const int p0 = sd.level_part_ptr[l], cnt = sd.level_part_ptr[l + 1] - p0;
    if (cnt <= 0) continue;
    k_sd_trsv<BS, 1><<<cnt, kSdThreads, sd_smem_bytes(sd.level_rounds_b[l], sd.level_local_b[l], BS), s>>>(
        sd.parts_b.p + p0, sd.U.rounds.p, sd.U.stream.p, sd.U.ring_rows.p, yp, ilu.order.p, io, sd.level_rounds_b[l]);
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
