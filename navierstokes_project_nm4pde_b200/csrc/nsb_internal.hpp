// nsb_internal.hpp -- device-side state of one engine handle and the kernel launch wrappers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nsb.h"
#include "nsb_host.hpp"

namespace nsb {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
struct ArgError : std::runtime_error { using std::runtime_error::runtime_error; };
struct StateError : std::runtime_error { using std::runtime_error::runtime_error; };
struct NoConvergence : std::runtime_error { using std::runtime_error::runtime_error; };
struct NcclError : std::runtime_error { using std::runtime_error::runtime_error; };

#define NSB_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess)                                                                          \
      throw nsb::CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                           std::to_string(__LINE__) + ")");                                        \
  } while (0)

// ---- programmatic dependent launch (sm_90+) for chains of dependent sweeps (one launch per colour of a
// triangular solve).  A sweep kernel calls pdl_launch_dependents() first -- the next sweep of the chain may then be
// scheduled while this one runs -- reads everything that does NOT depend on the previous sweep (row ids, matrix
// entries, diagonal), and calls pdl_wait() before its first access to the solution vector: the wait returns when the
// preceding grid has completed and its writes are visible.  Without the launch attribute both are no-ops, so the
// same kernels serve ordinary launches.  launch_k() launches with the attribute when `pdl` is set; it is captured
// into CUDA graphs as a programmatic edge.  NSB_PDL=0 switches it off.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args &&...args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  NSB_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}
#endif
bool pdl_enabled(); // kernels_linalg.cu: NSB_PDL (read when a solve is captured)
struct Handle;
void ilu_reset_graphs(Handle &H);

constexpr int kMaxQ = 16;

// reference-element tables, uploaded once (nsb_set_quadrature)
struct FeTables {
  int nq;
  double w[kMaxQ];
  double phi[10][kMaxQ];     // P2 values      [node][q]
  double dphi[10][kMaxQ][3]; // P2 ref. grads  [node][q][d]
  double psi[4][kMaxQ];      // P1 values      [vertex][q]
};

// Pre-contracted reference tensors for the step assembly ("assemble_step"):
//   C_ij = |detJ| * sum_{a,k} T[i][j][a][k] * (J^{-1} U_a)[k]
//   T[i][j][a][k] = sum_q w_q phi_i (dphi_j[k] phi_a  +  temam/2 * phi_j dphi_a[k])
//   rhs_i[c] = |detJ|/dt * sum_a Mh[i][a] U_a[c],  Mh[i][a] = sum_q w_q phi_i phi_a
struct StepTensor {
  double T[10][10][10][3];
  double Mh[10][10];
};

// Setup dry run -- TEST INFRASTRUCTURE (nsb_debug_setup_fingerprint, tests/test_setup_fingerprint.py): while it is on,
// the device buffers of setup are not allocated and every array setup would upload is hashed (FNV-1a over its bytes)
// into `log` instead, so the host-built device data structures (patterns, scatter map, SELL / block-SELL storage, ILU
// orderings) can be pinned on a machine without a GPU.  Nothing is computed in this mode and no entry point of the
// product path ever switches it on: without a device nsb_create still fails ("no CPU fallback").
struct SetupDryRun {
  bool on = false;
  std::vector<uint64_t> log;
  void record(uint64_t tag, const void *data, size_t bytes)
  {
    uint64_t h = 1469598103934665603ull ^ tag;
    const unsigned char *b = static_cast<const unsigned char *>(data);
    // eight bytes at a time (the arrays are hundreds of MB at bench size), tail byte by byte
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
      uint64_t w;
      std::memcpy(&w, b + i, 8);
      h = (h ^ w) * 1099511628211ull;
    }
    for (; i < bytes; ++i) h = (h ^ b[i]) * 1099511628211ull;
    log.push_back(h ^ (uint64_t(bytes) << 1));
  }
};
inline SetupDryRun g_dry;
// raw CUDA calls that setup makes besides DevBuf (attributes, small copies, synchronisation): skipped in a dry run
#define NSB_CUDA_SETUP(call) do { if (!nsb::g_dry.on) NSB_CUDA(call); } while (0)

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  void alloc(size_t count)
  {
    release();
    n = count;
    if (g_dry.on) { g_dry.record(0xA110C, &count, sizeof(count)); return; }
    if (count) NSB_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
  }
  void upload(const T *h, size_t count)
  {
    if (g_dry.on) { n = count; g_dry.record(sizeof(T), h, count * sizeof(T)); return; }
    if (n != count) alloc(count);
    if (n) NSB_CUDA(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice));
  }
  void upload(const std::vector<T> &h) { upload(h.data(), h.size()); }
  void zero(cudaStream_t s = 0)
  {
    if (g_dry.on) return;
    if (n) NSB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
  void release()
  {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
};

// CSR on the device; `bs` doubles per stored entry (1 for F_s, Mp, S; dim for B and Bt).
struct DevCsr {
  int n_rows = 0, n_cols = 0, bs = 1;
  int64_t nnz = 0;
  DevBuf<int> rowptr, colind;
  DevBuf<double> val;
  void upload_pattern(const Csr &h, int bs_)
  {
    n_rows = h.n_rows; n_cols = h.n_cols; bs = bs_; nnz = h.nnz();
    rowptr.upload(h.rowptr);
    colind.upload(h.colind);
    val.alloc(size_t(nnz) * bs);
    val.zero();
  }
};

// SELL-32 storage (kernels_sell.cu)
struct DevSell {
  int n_slices = 0, lanes = 1;    // lanes per row (1 | 4), see k_sell3
  int64_t n_slots = 0;
  DevBuf<int> slice_ptr, rowid, col, map;
  DevBuf<double> val;
  std::vector<int> range_slice; // first slice of each row range (colour)
};

// Block-sequential storage of one triangular factor in the block multicolour ordering (kernels_sell.cu,
// "bsell"): a block is <= 32 consecutive factor rows solved by one warp.  Entries that couple with other
// blocks ("ext", all of them final when the block's colour is swept) are stored in four passes of eight
// rows with four lanes per row (rows sorted by their ext count) and reduced per row by two shuffles; entries
// inside the block ("int") are packed row by row and resolved sequentially by shuffle broadcast.
struct DevBsell {
  int n_blocks = 0, max_int = 0;    // max_int: most intra-block entries of any block
  int max_nx = 0;                   // most distinct outside rows of any block
  std::vector<int> col_max_nx;      // the same per block colour (sizes the shared memory of that colour's launch)
  int64_t n_ext = 0, n_int = 0;
  DevBuf<int> e_ptr, e_map;         // e_ptr[b]: first ext slot of block b (multiple of 32)
  DevBuf<unsigned short> e_lix;     // per ext slot: index into the block's list of distinct outside rows
  DevBuf<int> e_col;                // per ext slot: the same as a factor row (sweeps that do not stage)
  DevBuf<int> x_ptr, x_ids;         // per block: its distinct outside rows (factor rows), staged into shared memory once
  DevBuf<unsigned> e_len;           // steps of the four ext passes of each block (one byte each)
  DevBuf<unsigned char> e_prow;     // [n_blocks][32] local row of each (pass, slot)
  DevBuf<double> e_val;
  DevBuf<int> i_ptr, i_map;         // i_ptr[b]: first intra entry of block b
  DevBuf<unsigned short> i_off;     // [n_blocks][33] offsets of each local row inside the block's entries
  DevBuf<unsigned> i_mask;          // per block: local rows that occur as a column of an intra entry
  DevBuf<unsigned char> i_col;      // local column (0..31); ascending per row (L), descending (U)
  DevBuf<double> i_val;
};

// Subdomain-resident storage of the triangular factors (kernels_sd.cu, ilu_ordering = 3): the rows of
// every part packed in processing order as a sequence of ROUNDS (one bulk copy each), the rows left
// after the last level in SELL-32 (DevIlu::sellL / sellU).
struct SdPart {
  int row0, ni;          // first factor row / number of rows of the part
  int round0, nrounds;   // rounds of the direction this copy is for
  int ring0, nring;      // rows of other levels staged next to the part's own rows (its "ring")
};
struct DevSdTri {
  int n_rounds = 0, n_slices = 0;
  int64_t n_doubles = 0;
  DevBuf<int4> rounds;                 // x: stream offset / 16 B, y: bytes, z: slices, w: colour
  DevBuf<int2> fill_slices;            // x: offset of the slice's values in the stream (doubles), y: steps
  DevBuf<int64_t> map_off;             // first fill-map slot of each slice
  DevBuf<int> map, ring_rows;
  DevBuf<int2> dfill;                  // backward only: (stream offset in doubles, factor row) of every dinv slot
  DevBuf<double> stream;
};
struct DevSd {
  int n_parts = 0, n_interior = 0;
  DevSdTri L, U;
  DevBuf<SdPart> parts_f, parts_b;
  std::vector<int> sep_colour_ptr;
  std::vector<int> level_part_ptr;               // parts of each level
  std::vector<int> level_local_f, level_local_b; // largest (rows + ring) of a level's parts, per direction
  std::vector<int> level_rounds_f, level_rounds_b; // most rounds of a level's parts, per direction
};

// ILU(0) factors in Ifpack's storage convention (strict lower part = a_ij * dinv_j, strict upper
// part scaled by dinv_i, inverse diagonal separate), on the owned-columns pattern, plus the
// level schedules of the two triangular solves.
struct DevIlu {
  int n = 0, bs_rhs = 1;
  int64_t nnz = 0;
  DevBuf<int> rowptr, colind, diagpos; // diagpos[i]: first entry with col > i  (end of L part)
  DevBuf<int> src;                     // position of each entry in the source matrix
  DevBuf<int> order;                   // factor row k = matrix row order[k]
  std::vector<int> h_order;
  DevBuf<double> val, dinv;
  // multicolour mode: split L / U factors and per-colour row blocks for the CSR-stream solves
  bool stream = false, sell = false, bsell = false, sdmode = false;
  DevSd sd;                       // subdomain-resident mode (ilu_ordering = 3)
  DevSell sellL, sellU;
  DevBsell bL, bU;                // block multicolour mode (ilu_ordering = 2)
  DevBuf<int> blk_row;            // first factor row of each block, [n_blocks + 1]
  std::vector<int> colour_blk;    // first block of each block colour
  DevBuf<double> io;              // TrsvIo slot: caller's in / out pointers of the captured solve
  DevBuf<int> Lp, Lc, Up, Uc, mapL, mapU, blkL, blkU;
  DevBuf<double> Lv, Uv;
  std::vector<int> colour_ptr, cblkL, cblkU;
  // forward (L / factorisation) and backward (U) level schedules
  std::vector<int> lvl_ptr_f, lvl_ptr_b;
  DevBuf<int> lvl_rows_f, lvl_rows_b;
  // chunked persistent schedule (sptrsv_kernel = 2)
  DevBuf<int> chunk_ptr_f, chunk_ptr_b, chunk_lvl_ptr_f, chunk_lvl_ptr_b, chunk_rows_f, chunk_rows_b;
  DevBuf<int> chunk_dep_ptr_f, chunk_dep_f, chunk_dep_ptr_b, chunk_dep_b;
  DevBuf<int> chunk_flags, chunk_ticket;
  int n_chunks = 0;
  int max_chunk_rows = 0;
  cudaGraphExec_t graph_f = nullptr, graph_b = nullptr; // per-level launches captured once
  double *graph_x = nullptr;
};

struct Halo; // multi-rank exchange (halo.cu)

// ---- peer-memory transport (halo.cu): mailbox layout shared by all ranks, and the device helpers the
// fused reductions use
constexpr int kMaxRanks = 16, kArSlots = 64;
struct MailHdr {
  long long dir_u[kMaxRanks], dir_p[kMaxRanks]; // where rank r's values go in this rank's inbox (-1: none)
  long long nvals_u, nvals_p;                   // values per inbox buffer (one parity)
  unsigned long long ar_flag[2][kMaxRanks];     // [parity][source rank] = sequence number of the last push
  unsigned long long hu_flag[2][kMaxRanks];
  unsigned long long hp_flag[2][kMaxRanks];
  double ar_slot[2][kMaxRanks][kArSlots];
};
struct PeerTab {
  int nranks, me, n_nb;
  MailHdr *peer[kMaxRanks];                      // mapped mailboxes (peer[me] = own)
  int nb_rank[kMaxRanks];
  int send_ptr_u[kMaxRanks + 1], send_ptr_p[kMaxRanks + 1];
  double *dst_u[kMaxRanks], *dst_p[kMaxRanks];   // parity-0 destination of my values in neighbour k's inbox
  long long stride_u[kMaxRanks], stride_p[kMaxRanks];
};
struct ArArgs {
  const PeerTab *tab; // nullptr: no fused all-reduce
  int parity;
  unsigned long long seq;
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// spin until *f >= seq.  Ranks run the same sequence of exchanges but their hosts may be skewed by seconds
// (I/O between steps, a first-call graph capture): the limit is generous (~2 minutes); a peer that never
// arrives still ends in a trap -- an error on this rank -- instead of a hung GPU.
__device__ __forceinline__ void wait_flag(const unsigned long long *f, unsigned long long seq)
{
  const long long t0 = clock64();
  while (ld_acquire_sys(f) < seq) {
    __nanosleep(40);
    if (clock64() - t0 > 240000000000LL) __trap();
  }
}
// All-reduce (sum) of n <= kArSlots values held one per thread (thread j holds value j) of ONE block;
// every thread of the block must call.  Stores the value into every peer's slot, raises the flags,
// waits for all peers and adds the partials in rank order (bitwise identical on every rank).
__device__ __forceinline__ double ar_exchange_block(const PeerTab *T, int parity, unsigned long long seq, double v, int j, int n)
{
  const int nr = T->nranks, me = T->me;
  if (j < n)
    for (int r = 0; r < nr; ++r) T->peer[r]->ar_slot[parity][me][j] = v;
  __threadfence_system();
  __syncthreads();
  if (j < nr) {
    st_release_sys(&T->peer[j]->ar_flag[parity][me], seq);
    wait_flag(&T->peer[me]->ar_flag[parity][j], seq);
  }
  __syncthreads();
  double s = 0.0;
  if (j < n)
    for (int r = 0; r < nr; ++r) s += __ldcg(&T->peer[me]->ar_slot[parity][r][j]);
  return s;
}
#endif

struct Handle {
  int dim = 0, n2 = 0, nv1 = 0, dpc = 0;
  int device = 0, nranks = 1, rank = 0;
  std::string err;
  nsb_params prm{};
  bool have_mesh = false, have_quad = false, finalized = false, assembled = false, prec_ready = false;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;

  // ---- host copies of the static problem description
  int64_t nc = 0, nc_pad = 0;
  int n_nodes = 0, n_p = 0, n_nodes_owned = 0, n_p_owned = 0;
  std::vector<double> h_vcoords;
  std::vector<int> h_cell_nodes, h_cell_p;
  Csr hFs, hB, hBt, hMp, hS;
  FeTables h_tab{};
  std::vector<int> h_dir_nodes; // constrained P2 nodes (owned)
  std::vector<int> h_dir_rows;  // as given by the caller (dof = dim*node + c)
  std::vector<int> h_dir_slot;  // caller row k -> slot in d_dir_vals (node-major, comp-minor)

  // ---- layout of local vectors: [u owned | p owned | u ghost | p ghost]
  int nu_owned() const { return dim * n_nodes_owned; }
  int n_owned() const { return dim * n_nodes_owned + n_p_owned; }
  int n_local() const { return dim * n_nodes + n_p; }
  // offset added to dim*node when node >= n_nodes_owned
  int ghost_off_u() const { return n_p_owned; }
  // pressure p lives at p_base() + p (+ ghost_off_p() when p >= n_p_owned)
  int p_base() const { return dim * n_nodes_owned; }
  int ghost_off_p() const { return dim * (n_nodes - n_nodes_owned); }

  // ---- device: mesh
  DevBuf<double> d_vcoords;       // [nc_pad/32][(dim+1)*dim][32]  (cell-interleaved, coalesced)
  DevBuf<int> d_cell_nodes;       // [nc_pad/32][n2][32]
  DevBuf<int> d_cell_p;           // [nc_pad/32][nv1][32]
  DevBuf<int> d_mapF;             // [nc_pad/32][n2*n2][32] position in F_s values, -1 = skip
  DevBuf<FeTables> d_tab;
  DevBuf<StepTensor> d_step_tensor;

  // ---- device: matrices (compact layout, DESIGN.md "data layout")
  DevCsr Fs;                      // system_matrix.block(0,0) = I_dim (x) F_s
  DevBuf<double> d_K, d_M, d_A;   // K = M + A, mass/dt, stiffness on the F_s pattern
  DevBuf<double> d_C;             // convection (parity harness only, lazily allocated)
  DevCsr B, Bt, Mp, S;
  DevBuf<int> d_diagF;            // position of the diagonal in each F_s row
  DevSell sellF;                  // F_s in SELL-32 (3D SpMV), refreshed lazily after assembly
  bool sellF_dirty = true;
  DevBuf<double> d_xpad;          // padded (4 doubles per node) copy of the SpMV input
  DevBuf<int> d_blk_Fs, d_blk_S;  // CSR-stream row blocks (kernels_stream.cu)
  int n_blk_Fs = 0, n_blk_S = 0;
  DevBuf<double> d_massdiag, d_masslump;
  DevBuf<double> d_D, d_Dinv, d_negDinv; // per velocity DoF (node-interleaved)
  DevIlu iluF, iluS;
  DevBuf<int> d_spgemm_ws;

  // ---- device: boundary data
  DevBuf<int> d_dir_nodes;
  DevBuf<double> d_dir_vals;      // [n_dir_nodes][dim]
  DevBuf<double> d_neumann;
  bool have_neumann = false;

  // ---- device: obstacle faces of compute_forces (kernels_post.cu), face-minor arrays
  int n_force_faces = 0, force_nq = 0;
  DevBuf<double> d_ff_x;          // [(dim+1)*dim][n_faces] vertex coordinates of the face's cell
  DevBuf<int> d_ff_nodes, d_ff_p; // [n2][n_faces], [nv1][n_faces]
  DevBuf<int> d_ff_opp;           // [n_faces] local vertex opposite to the face
  DevBuf<double> d_ff_q;          // xi[nq][dim-1], w[nq]
  DevBuf<double> d_ff_part;       // [2][n_faces] per-face drag / lift, then the two sums

  // ---- device: vectors (local layout)
  DevBuf<double> d_sol, d_rhs;
  DevBuf<double> d_scratch;       // reductions etc.
  double *h_pinned = nullptr;     // pinned host scratch for scalars
  std::vector<double> h_dir_vals;

  Halo *halo = nullptr;

  // ---- solver workspace (solver.cu)
  struct SolverWs {
    DevBuf<double> V_outer;   // n_tmp x n_local
    DevBuf<double> V_inner;   // n_tmp x max(dim*n_nodes, n_p)
    DevBuf<double> tu[4];     // velocity temporaries (layout U: [owned | ghost])
    DevBuf<double> tp[6];     // pressure temporaries (layout P)
    DevBuf<double> scal;      // device scalars: [0,64) outer GMRES, [64,128) inner, [128,192) CG, 200.. norms
    DevBuf<double> prec_in, prec_out; // staging for the operator-level entry points
  };
  SolverWs *ws = nullptr;

  // ---- statistics of the last solve
  long n_inner_F = 0, n_inner_S = 0, n_F_solves = 0, n_S_solves = 0, n_vmult = 0;
  long cnt_spmv_F = 0, cnt_spmv_S = 0, cnt_spmv_B = 0, cnt_spmv_Bt = 0, cnt_ilu_F = 0, cnt_ilu_S = 0, cnt_dot = 0,
       cnt_sync = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_s0 = nullptr, ev_s1 = nullptr;
  double t_step_dev_ms = 0; // device time of the steps run through nsb_step_host since the last reset
  int last_outer = 0;
  double last_res = 0, t_assemble_ms = 0, t_prec_ms = 0, t_solve_ms = 0;

  ~Handle();
};

// ---------------------------------------------------------------- kernels_assembly.cu
void launch_assemble_first(Handle &H);
void launch_assemble_step(Handle &H, double *F_target);
void launch_apply_dirichlet(Handle &H, bool clear_bt);
void build_step_tensor(const FeTables &tab, int dim, bool temam, StepTensor &out);

// ---------------------------------------------------------------- kernels_linalg.cu
// y_u = F_s x_u (+ Bt x_p when xp != nullptr).  x, y: local vectors (bases of the u blocks).
// goff_u / goff_p: extra offset of ghost entries in the given vector (see Handle::ghost_off_*).
void spmv_F(Handle &H, const double *x_u, int goff_u, const double *x_p, int goff_p, double *y_u);
void spmv_B(Handle &H, const double *x_u, int goff_u, double *y_p);   // y_p = B x_u
void spmv_Bt(Handle &H, const double *x_p, int goff_p, double *y_u);  // y_u = Bt x_p
void spmv_S(Handle &H, const double *x_p, int goff_p, double *y_p);
void vec_copy(Handle &H, int n, const double *x, double *y);
void vec_zero(Handle &H, int n, double *x);
void vec_axpy(Handle &H, int n, double a, const double *x, double *y);             // y += a x
void vec_axpy_dev(Handle &H, int n, const double *a_dev, double sign, const double *x, double *y);
void vec_sadd(Handle &H, int n, double s, double a, const double *x, double *y);   // y = s y + a x
void vec_scale(Handle &H, int n, double a, double *x);
void vec_scale_inv_dev(Handle &H, int n, const double *s_dev, double *x);          // x /= *s_dev
void vec_pointwise(Handle &H, int n, const double *d, double *x);                  // x *= d
void vec_pointwise_out(Handle &H, int n, const double *d, const double *x, double *y); // y = d .* x
// out_dev[0] = sum x_i y_i over n entries (this rank); deterministic two-stage reduction
void vec_dot_dev(Handle &H, int n, const double *x, const double *y, double *out_dev, ArArgs ar = ArArgs{nullptr, 0, 0});
// vv += sign * (*a_dev) * v_prev ; out_dev = vv . v_next      (SolverGMRES add_and_dot)
void vec_add_and_dot_dev(Handle &H, int n, double *vv, const double *a_dev, double sign, const double *v_prev,
                         const double *v_next, double *out_dev, ArArgs ar = ArArgs{nullptr, 0, 0});
// batched classical Gram-Schmidt pieces: h[j] = vv . V_j (j < nv), *self = vv . vv ; vv -= sum h[j] V_j
// allreduce = true (multi-rank, peer-memory transport only): the kernels also sum over the ranks
void vec_multi_dot_dev(Handle &H, int n, const double *vv, const double *V, size_t ld, int nv, double *h_dev,
                       double *self_dev, bool allreduce = false);
void vec_multi_axpy_dev(Handle &H, int n, double *vv, const double *V, size_t ld, int nv, const double *h_dev,
                        double *norm2_dev, bool allreduce = false);
// xyz: support points of the rows [n_rows][gdim] (used by ordering 3; may be nullptr)
void ilu_build(Handle &H, DevIlu &ilu, const Csr &A, int n_owned_cols, int bs_rhs, int ordering, const double *xyz = nullptr,
               int gdim = 0);
void ilu_factor(Handle &H, DevIlu &ilu, const double *A_val);
void ilu_solve(Handle &H, DevIlu &ilu, const double *x, double *y); // y = U^-1 D^-1 L^-1 x
void spgemm_schur(Handle &H); // S = B diag(negDinv) Bt on the static pattern
void extract_diag(Handle &H);  // d_D / d_Dinv / d_negDinv per preconditioner type
void mass_rows(Handle &H);     // d_massdiag / d_masslump from d_M
void flush_l2(Handle &H);

// ---------------------------------------------------------------- kernels_stream.cu
void stream_build_spmv(Handle &H);
void stream_spmv_F(Handle &H, const double *x_u, int goff_u, double *y_u);
void stream_spmv_S(Handle &H, const double *x_p, int goff_p, double *y_p);
void stream_build_ilu(Handle &H, DevIlu &ilu, const std::vector<int> &rowptr, const std::vector<int> &colind,
                      const std::vector<int> &diagpos, const std::vector<int> &colour_ptr);
void stream_split_factors(Handle &H, DevIlu &ilu);
void stream_trsv(Handle &H, DevIlu &ilu, double *y, cudaStream_t s);

// ---------------------------------------------------------------- kernels_sell.cu
void sell_build(const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &src,
                const std::vector<int> &ranges, int window, int lanes, DevSell &out);
void sell_build(const std::vector<int> &rowptr, const int *colind, const int *src, const std::vector<int> &ranges, int window,
                int lanes, DevSell &out);
// host copies of a SELL-32 structure and the L / U split of a factor pattern (shared by stream_build_ilu and the CPU
// emulation of the sweeps, sell_host_check)
struct SellHost { std::vector<int> slice_ptr, rowid, col, map; };
void sell_build_keep(const std::vector<int> &rowptr, const int *colind, const int *src, const std::vector<int> &ranges,
                     int window, int lanes, DevSell &out, SellHost *keep);
void split_lu(const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
              std::vector<int> &Lp, std::unique_ptr<int[]> &Lc, std::unique_ptr<int[]> &mapL, std::vector<int> &Up,
              std::unique_ptr<int[]> &Uc, std::unique_ptr<int[]> &mapU);
int sell_lanes_for(int n_rows);
void sell_fill(Handle &H, DevSell &S, const double *src);
void sell_spmv_F(Handle &H, const double *x_u, int goff_u, double *y_u);
void sell_trsv(Handle &H, DevIlu &ilu, double *yp, cudaStream_t s);
void sell_set_io(Handle &H, DevIlu &ilu, const double *x, double *y);
// block multicolour ILU(0) solves
void bsell_build(DevIlu &ilu, const std::vector<int> &rowptr, const std::vector<int> &colind,
                 const std::vector<int> &diagpos, const std::vector<int> &blk_ptr, const std::vector<int> &colour_blk);
void bsell_fill(Handle &H, DevIlu &ilu);
int bsell_stride(int bs_rhs); // doubles per row of the staging vector
void bsell_trsv(Handle &H, DevIlu &ilu, double *yp, cudaStream_t s);

// ---------------------------------------------------------------- kernels_sd.cu
void sd_build(DevIlu &ilu, const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
              const std::vector<int> &part_ptr, const std::vector<int> &pcol_ptr, const std::vector<int> &pcol,
              const std::vector<int> &sep_colour_ptr, const std::vector<int> &level_part_ptr);
size_t sd_smem_needed(const std::vector<int> &rowptr, const std::vector<int> &colind, const std::vector<int> &diagpos,
                      const std::vector<int> &part_ptr, int bs);
void sd_fill(Handle &H, DevIlu &ilu);
int sd_stride(int bs_rhs);
int sd_launches(const DevIlu &ilu); // kernel launches of one triangular solve pair
void sd_trsv(Handle &H, DevIlu &ilu, double *yp, cudaStream_t s);

// ---------------------------------------------------------------- kernels_post.cu
void force_faces_set(Handle &H, int nf, const int *face_cell, const int *face_opp, int nq, const double *xi, const double *w);
void force_faces_compute(Handle &H, double rho, double *out);

// ---------------------------------------------------------------- solver.cu
void solver_alloc(Handle &H);
void solver_free(Handle &H);
void precond_init(Handle &H);
void precond_vmult(Handle &H, const double *src, double *dst);
void system_vmult(Handle &H, const double *x, double *y);
int solve_outer(Handle &H);
// all-reduce (sum) of `n` device doubles across ranks (no-op on one rank) and fetch to host
void reduce_fetch(Handle &H, double *dev, int n, double *host_out);

// ---------------------------------------------------------------- halo.cu
void halo_create(Handle &H, const void *unique_id);
void halo_destroy(Handle &H);
void get_unique_id(void *out128);
void halo_set_plan(Handle &H, int n_nb, const int *nb_rank, const int *send_node_ptr, const int *send_node_idx,
                   const int *recv_node_cnt, const int *send_p_ptr, const int *send_p_idx, const int *recv_p_cnt);
// fill the ghost entries of a velocity / pressure vector whose ghosts start at x + (owned) + goff
void halo_exchange_u(Handle &H, double *x_u, int goff_u);
void halo_exchange_p(Handle &H, double *x_p, int goff_p);
void halo_allreduce(Handle &H, double *dev, int n);
void halo_p2p_export(Handle &H, void *handle64);
void halo_p2p_attach(Handle &H, const void *handles);
bool halo_is_p2p(const Handle &H);
ArArgs halo_ar_args(Handle &H);

} // namespace nsb
