// nsb_api.cu -- the C ABI (include/nsb.h): argument checking, static setup, layout conversion
// between the reference's block-CSR / DoF numbering and the compact device layout.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>

#include "nsb_internal.hpp"

using namespace nsb;

static thread_local std::string g_create_error;

struct nsb_handle_s {
  Handle H;
};

Handle::~Handle()
{
  solver_free(*this);
  halo_destroy(*this);
  for (DevIlu *ilu : {&iluF, &iluS}) {
    if (ilu->graph_f) cudaGraphExecDestroy(ilu->graph_f);
    if (ilu->graph_x) cudaFree(ilu->graph_x);
  }
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (ev_s0) cudaEventDestroy(ev_s0);
  if (ev_s1) cudaEventDestroy(ev_s1);
  if (h_pinned) cudaFreeHost(h_pinned);
  if (stream) cudaStreamDestroy(stream);
}

template <typename Fn>
static int guarded(nsb_handle h, Fn &&fn)
{
  if (!h) return NSB_ERR_ARG;
  try {
    NSB_CUDA(cudaSetDevice(h->H.device));
    fn(h->H);
    return NSB_OK;
  } catch (const ArgError &e) { h->H.err = e.what(); return NSB_ERR_ARG;
  } catch (const CudaError &e) { h->H.err = e.what(); return NSB_ERR_CUDA;
  } catch (const StateError &e) { h->H.err = e.what(); return NSB_ERR_STATE;
  } catch (const NoConvergence &e) { h->H.err = e.what(); return NSB_ERR_NOCONV;
  } catch (const NcclError &e) { h->H.err = e.what(); return NSB_ERR_NCCL;
  } catch (const std::exception &e) { h->H.err = e.what(); return NSB_ERR_STATE; }
}

static void sync(Handle &H) { NSB_CUDA(cudaStreamSynchronize(H.stream)); }
// Host -> device copy ordered on the engine's stream.  H.stream is a non-blocking stream, so a plain
// cudaMemcpy (legacy default stream) is NOT ordered against it, and for pageable memory cudaMemcpy may
// return while the DMA of its last staging chunk is still in flight: kernels launched on H.stream right
// afterwards could read the tail of the previous content.
static void h2d(Handle &H, void *dst, const void *src, size_t bytes)
{
  if (!bytes) return;
  NSB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, H.stream));
  sync(H);
}

// ------------------------------------------------------------------------------------------------
extern "C" int nsb_default_params(nsb_params *p, int variant)
{
  if (!p || variant < 0 || variant > 2) return NSB_ERR_ARG;
  std::memset(p, 0, sizeof(*p));
  p->variant = variant;
  p->nu = (variant == NSB_VARIANT_CONV) ? 1e-2 : 1e-3;
  p->deltat = (variant == NSB_VARIANT_2D) ? 0.01 : (variant == NSB_VARIANT_3D ? 0.0002 : 0.0004);
  p->precond_type = (variant == NSB_VARIANT_2D) ? NSB_PREC_ASIMPLE : NSB_PREC_YOSIDA;
  p->gmres_tmp = 30;
  p->outer_maxit = 100000;
  p->outer_tol = 1e-4;
  p->inner_maxit = (variant == NSB_VARIANT_2D) ? 10000 : 100000;
  p->inner_rtol = 1e-2;
  p->alpha_simple = 0.5;
  p->alpha_asimple = 1.0;
  p->ilu_ordering_schur = -1; // same as ilu_ordering
  return NSB_OK;
}

extern "C" int nsb_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int nsb_get_unique_id(void *out128)
{
  try { get_unique_id(out128); return NSB_OK; }
  catch (const std::exception &e) { g_create_error = e.what(); return NSB_ERR_NCCL; }
}

extern "C" int nsb_create(nsb_handle *out, int dim, int device_id, int nranks, int rank, const void *unique_id)
{
  if (!out || (dim != 2 && dim != 3) || nranks < 1 || rank < 0 || rank >= nranks) {
    g_create_error = "nsb_create: bad arguments";
    return NSB_ERR_ARG;
  }
  nsb_handle h = nullptr;
  try {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      cudaGetLastError();
      throw CudaError("no CUDA device: this engine has no CPU fallback");
    }
    if (device_id < 0 || device_id >= ndev) throw ArgError("nsb_create: device id out of range");
    NSB_CUDA(cudaSetDevice(device_id));
    h = new nsb_handle_s();
    Handle &H = h->H;
    H.dim = dim; H.n2 = n2_of(dim); H.nv1 = dim + 1; H.dpc = dpc_of(dim);
    H.device = device_id; H.nranks = nranks; H.rank = rank;
    nsb_default_params(&H.prm, dim == 2 ? NSB_VARIANT_2D : NSB_VARIANT_3D);
    NSB_CUDA(cudaStreamCreateWithFlags(&H.stream, cudaStreamNonBlocking));
    NSB_CUDA(cudaMallocHost((void **)&H.h_pinned, sizeof(double) * 256));
    NSB_CUDA(cudaEventCreate(&H.ev0));
    NSB_CUDA(cudaEventCreate(&H.ev1));
    NSB_CUDA(cudaEventCreate(&H.ev_s0));
    NSB_CUDA(cudaEventCreate(&H.ev_s1));
    H.d_scratch.alloc(64 + 1024 + 8 + 9 * 1024 + 8); // dbar | dot partials | ticket | multi-dot partials | ticket
    H.d_scratch.zero();
    NSB_CUDA(cudaDeviceSynchronize());
    halo_create(H, unique_id);
    *out = h;
    return NSB_OK;
  } catch (const ArgError &e) { g_create_error = e.what(); delete h; return NSB_ERR_ARG;
  } catch (const NcclError &e) { g_create_error = e.what(); delete h; return NSB_ERR_NCCL;
  } catch (const std::exception &e) { g_create_error = e.what(); delete h; return NSB_ERR_CUDA; }
}

extern "C" int nsb_destroy(nsb_handle h)
{
  if (!h) return NSB_ERR_ARG;
  cudaSetDevice(h->H.device);
  cudaDeviceSynchronize();
  delete h;
  return NSB_OK;
}

extern "C" const char *nsb_last_error(nsb_handle h) { return h ? h->H.err.c_str() : g_create_error.c_str(); }

extern "C" int nsb_set_params(nsb_handle h, const nsb_params *p)
{
  return guarded(h, [&](Handle &H) {
    if (!p) throw ArgError("nsb_set_params: null");
    if (p->precond_type < 0 || p->precond_type > 3) throw ArgError("Invalid preconditioner type");
    // the device scalar slots of a Krylov solve are 64 doubles wide (solver.cu); the batched
    // Gram-Schmidt splits them into two halves of 32
    if (p->gmres_tmp < 3 || p->gmres_tmp > 60) throw ArgError("nsb_set_params: gmres_tmp must be in [3, 60]");
    if (p->orthogonalisation == 1 && p->gmres_tmp > 30)
      throw ArgError("nsb_set_params: orthogonalisation = 1 needs gmres_tmp <= 30 (the reference uses 30)");
    if (p->ilu_ordering < 0 || p->ilu_ordering > 3) throw ArgError("nsb_set_params: ilu_ordering must be 0, 1, 2 or 3");
    if (p->orthogonalisation < 0 || p->orthogonalisation > 1)
      throw ArgError("nsb_set_params: orthogonalisation must be 0 or 1");
    if (p->ilu_ordering_schur < -1 || p->ilu_ordering_schur > 3)
      throw ArgError("nsb_set_params: ilu_ordering_schur must be -1 (same as ilu_ordering), 0, 1, 2 or 3");
    if (H.finalized && (p->ilu_ordering != H.prm.ilu_ordering || p->ilu_ordering_schur != H.prm.ilu_ordering_schur))
      throw StateError("nsb_set_params: ilu_ordering must be chosen before nsb_finalize_setup");
    if (!(p->deltat > 0) || !(p->nu > 0)) throw ArgError("nsb_set_params: nu and deltat must be positive");
    const bool realloc_ws = H.finalized && p->gmres_tmp != H.prm.gmres_tmp;
    const bool retensor = H.finalized && p->variant != H.prm.variant;
    H.prm = *p;
    if (realloc_ws) solver_alloc(H);
    if (retensor) {
      StepTensor t;
      build_step_tensor(H.h_tab, H.dim, H.prm.variant != NSB_VARIANT_3D, t);
      NSB_CUDA(cudaMemcpy(H.d_step_tensor.p, &t, sizeof(t), cudaMemcpyHostToDevice));
    }
  });
}

// host side of nsb_set_mesh (no device work: also run by nsb_debug_setup_fingerprint)
static void set_mesh_host(Handle &H, int32_t n_cells, const double *vertex_coords, const int32_t *cell_dofs, int32_t n_u,
                          int32_t n_p, int32_t n_u_owned, int32_t n_p_owned)
{
  const int dim = H.dim, n2 = H.n2, nv1 = H.nv1, dpc = H.dpc;
  if (n_cells <= 0 || !vertex_coords || !cell_dofs) throw ArgError("nsb_set_mesh: empty mesh");
  if (n_u % dim || n_u_owned % dim || n_u_owned > n_u || n_p_owned > n_p || n_u <= 0 || n_p <= 0)
    throw ArgError("nsb_set_mesh: inconsistent DoF counts");
  H.nc = n_cells;
  H.n_nodes = n_u / dim; H.n_p = n_p; H.n_nodes_owned = n_u_owned / dim; H.n_p_owned = n_p_owned;
  H.h_vcoords.assign(vertex_coords, vertex_coords + size_t(n_cells) * nv1 * dim);
  H.h_cell_nodes.resize(size_t(n_cells) * n2);
  H.h_cell_p.resize(size_t(n_cells) * nv1);
  for (int64_t c = 0; c < n_cells; ++c) {
    const int32_t *cd = cell_dofs + c * dpc;
    for (int a = 0; a < n2; ++a) {
      const int base = (a < nv1) ? a * (dim + 1) : nv1 * (dim + 1) + (a - nv1) * dim;
      const int d0 = cd[base];
      if (d0 < 0 || d0 >= n_u || d0 % dim) throw ArgError("nsb_set_mesh: velocity DoFs must be node-interleaved (dof = dim*node + c)");
      for (int k = 1; k < dim; ++k)
        if (cd[base + k] != d0 + k) throw ArgError("nsb_set_mesh: velocity DoFs must be node-interleaved (dof = dim*node + c)");
      H.h_cell_nodes[c * n2 + a] = d0 / dim;
    }
    for (int v = 0; v < nv1; ++v) {
      const int p = cd[v * (dim + 1) + dim] - n_u;
      if (p < 0 || p >= n_p) throw ArgError("nsb_set_mesh: pressure DoF out of range");
      H.h_cell_p[c * nv1 + v] = p;
    }
    // orientation / degeneracy check
    const double *x = &H.h_vcoords[c * nv1 * dim];
    double J[3][3];
    for (int r = 0; r < dim; ++r)
      for (int k = 0; k < dim; ++k) J[r][k] = x[(k + 1) * dim + r] - x[r];
    const double det = dim == 2 ? J[0][0] * J[1][1] - J[0][1] * J[1][0]
                                : J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) -
                                      J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                                      J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    if (!(det > 0)) throw ArgError("nsb_set_mesh: cell with non-positive Jacobian");
  }
  H.have_mesh = true;
  H.finalized = H.assembled = H.prec_ready = false;
}

extern "C" int nsb_set_mesh(nsb_handle h, int32_t n_cells, const double *vertex_coords, const int32_t *cell_dofs,
                            int32_t n_u, int32_t n_p, int32_t n_u_owned, int32_t n_p_owned)
{
  return guarded(h, [&](Handle &H) { set_mesh_host(H, n_cells, vertex_coords, cell_dofs, n_u, n_p, n_u_owned, n_p_owned); });
}

static void tabulate(Handle &H, int nq, const double *xi, const double *w)
{
  const int dim = H.dim, nv1 = dim + 1, ne = dim == 2 ? 3 : 6;
  FeTables &t = H.h_tab;
  std::memset(&t, 0, sizeof(t));
  t.nq = nq;
  for (int q = 0; q < nq; ++q) {
    t.w[q] = w[q];
    double l[4], gl[4][3];
    double s = 0;
    for (int d = 0; d < dim; ++d) s += xi[q * dim + d];
    l[0] = 1.0 - s;
    for (int d = 0; d < dim; ++d) l[d + 1] = xi[q * dim + d];
    for (int v = 0; v < nv1; ++v)
      for (int d = 0; d < 3; ++d) gl[v][d] = (d >= dim) ? 0.0 : (v == 0 ? -1.0 : (v - 1 == d ? 1.0 : 0.0));
    for (int v = 0; v < nv1; ++v) {
      t.phi[v][q] = l[v] * (2.0 * l[v] - 1.0);
      for (int d = 0; d < dim; ++d) t.dphi[v][q][d] = (4.0 * l[v] - 1.0) * gl[v][d];
      t.psi[v][q] = l[v];
    }
    for (int e = 0; e < ne; ++e) {
      const int a = kEdgeA[e], b = kEdgeB[e];
      t.phi[nv1 + e][q] = 4.0 * l[a] * l[b];
      for (int d = 0; d < dim; ++d) t.dphi[nv1 + e][q][d] = 4.0 * (l[b] * gl[a][d] + l[a] * gl[b][d]);
    }
  }
}

extern "C" int nsb_set_quadrature(nsb_handle h, int32_t n_q, const double *xi, const double *w)
{
  return guarded(h, [&](Handle &H) {
    if (n_q < 1 || n_q > kMaxQ || !xi || !w) throw ArgError("nsb_set_quadrature: 1 <= n_q <= 16 required");
    tabulate(H, n_q, xi, w);
    H.d_tab.alloc(1);
    NSB_CUDA(cudaMemcpy(H.d_tab.p, &H.h_tab, sizeof(FeTables), cudaMemcpyHostToDevice));
    H.have_quad = true;
    if (H.finalized) {
      StepTensor t;
      build_step_tensor(H.h_tab, H.dim, H.prm.variant != NSB_VARIANT_3D, t);
      NSB_CUDA(cudaMemcpy(H.d_step_tensor.p, &t, sizeof(t), cudaMemcpyHostToDevice));
    }
  });
}

template <typename T>
static std::vector<T> interleave32(const std::vector<T> &in, int64_t nc, int64_t nc_pad, int k, T pad)
{ // [nc][k] -> [nc_pad/32][k][32]
  std::vector<T> out;
  reserve_prefaulted(out, size_t(nc_pad) * k);
  out.assign(size_t(nc_pad) * k, pad);
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < nc_pad / 32; ++g) // one group of 32 cells = one contiguous piece of `out`
    for (int64_t c = g * 32; c < std::min(nc, g * 32 + 32); ++c)
      for (int j = 0; j < k; ++j) out[(g * k + j) * 32 + (c & 31)] = in[c * k + j];
  return out;
}

// dry run only: the host-side members of the setup structures (launch geometry, colour / level boundaries) join
// the fingerprint after the uploads
static void dry_record_host_state(Handle &H)
{
  auto rec_vec = [](const std::vector<int> &v) { g_dry.record(0x4057, v.data(), v.size() * sizeof(int)); };
  auto rec_i64 = [](std::initializer_list<int64_t> v) { g_dry.record(0x4058, v.begin(), v.size() * sizeof(int64_t)); };
  auto rec_sell = [&](const DevSell &S) {
    rec_i64({S.n_slices, S.lanes, S.n_slots});
    rec_vec(S.range_slice);
  };
  auto rec_bsell = [&](const DevBsell &B) {
    rec_i64({B.n_blocks, B.max_int, B.max_nx, B.n_ext, B.n_int});
    rec_vec(B.col_max_nx);
  };
  rec_i64({H.n_blk_Fs, H.n_blk_S, H.nc_pad});
  rec_sell(H.sellF);
  for (const DevIlu *ilu : {&H.iluF, &H.iluS}) {
    rec_i64({ilu->n, ilu->bs_rhs, ilu->nnz, ilu->stream, ilu->sell, ilu->bsell, ilu->sdmode});
    rec_vec(ilu->h_order); rec_vec(ilu->colour_blk); rec_vec(ilu->colour_ptr); rec_vec(ilu->cblkL); rec_vec(ilu->cblkU);
    rec_vec(ilu->lvl_ptr_f); rec_vec(ilu->lvl_ptr_b);
    rec_sell(ilu->sellL); rec_sell(ilu->sellU);
    rec_bsell(ilu->bL); rec_bsell(ilu->bU);
  }
}

// nsb_finalize_setup: everything static is built on the host here and uploaded (the dry run of
// nsb_debug_setup_fingerprint hashes the uploads instead)
static void finalize_setup(Handle &H)
{
  if (!H.have_mesh || !H.have_quad) throw StateError("nsb_finalize_setup: mesh and quadrature must be set first");
  const int dim = H.dim, n2 = H.n2, nv1 = H.nv1;
  const int64_t nc = H.nc;
  H.nc_pad = (nc + 31) / 32 * 32;
  // NSB_VERBOSE=1: where the (host-side, cold-path) setup time goes
  const bool verbose = getenv("NSB_VERBOSE") && atoi(getenv("NSB_VERBOSE")) > 0;
  auto t_last = std::chrono::steady_clock::now();
  auto phase = [&](const char *what) {
    if (!verbose) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[nsb setup rank %d] %-28s %8.2f s\n", H.rank, what, std::chrono::duration<double>(now - t_last).count());
    t_last = now;
  };
  // patterns (DoFTools::make_sparsity_pattern with the coupling table of NavierStokes2D.cpp:109-119)
  {
    RowCells by_node, by_p; // the node -> cells and pressure DoF -> cells lists serve two patterns each
    by_node.build(nc, H.h_cell_nodes.data(), n2, H.n_nodes);
    build_pattern(by_node, H.h_cell_nodes.data(), n2, H.n_nodes_owned, H.n_nodes, H.hFs);
    build_pattern(by_node, H.h_cell_p.data(), nv1, H.n_nodes, H.n_p, H.hBt);
    by_p.build(nc, H.h_cell_p.data(), nv1, H.n_p);
    build_pattern(by_p, H.h_cell_nodes.data(), n2, H.n_p_owned, H.n_nodes, H.hB);
    build_pattern(by_p, H.h_cell_p.data(), nv1, H.n_p_owned, H.n_p, H.hMp);
  }
  phase("sparsity patterns");
  symbolic_product(H.hB, H.hBt, H.hS);
  phase("symbolic Schur product");
  H.Fs.upload_pattern(H.hFs, 1);
  H.Bt.upload_pattern(H.hBt, dim);
  H.B.upload_pattern(H.hB, dim);
  H.Mp.upload_pattern(H.hMp, 1);
  H.S.upload_pattern(H.hS, 1);
  H.d_K.alloc(H.Fs.nnz); H.d_M.alloc(H.Fs.nnz); H.d_A.alloc(H.Fs.nnz); H.d_C.alloc(H.Fs.nnz);
  H.d_K.zero(); H.d_M.zero(); H.d_A.zero(); H.d_C.zero();
  // diagonal positions and the scatter map of the step kernel
  std::vector<int> diag(H.n_nodes_owned);
  for (int i = 0; i < H.n_nodes_owned; ++i) diag[i] = find_in_row(H.hFs, i, i);
  H.d_diagF.upload(diag);
  {
    // [nc_pad / 32][n2 * n2][32] positions in F_s, -1 for padding cells and ghost rows.  1.9 GB at 4.7 M tetrahedra:
    // every group of 32 cells (one contiguous piece) is initialised and filled by the thread that owns it; the n2
    // columns of a cell are sorted once and every row of F_s is walked once (merge) instead of n2 binary searches
    const size_t gsz = size_t(n2) * n2 * 32;
    std::unique_ptr<int[]> map(new int[size_t(H.nc_pad / 32) * gsz]);
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < H.nc_pad / 32; ++g) {
      int *mg = map.get() + size_t(g) * gsz;
      std::fill(mg, mg + gsz, -1);
      for (int64_t c = g * 32; c < std::min<int64_t>(nc, g * 32 + 32); ++c) {
        const int *cn = &H.h_cell_nodes[c * n2];
        int sj[10], sc[10]; // local column indices sorted by node id
        for (int j = 0; j < n2; ++j) sj[j] = j;
        std::sort(sj, sj + n2, [&](int a, int b) { return cn[a] < cn[b]; });
        for (int j = 0; j < n2; ++j) sc[j] = cn[sj[j]];
        for (int i = 0; i < n2; ++i) {
          if (cn[i] >= H.n_nodes_owned) continue;
          int *mi = mg + size_t(i) * n2 * 32 + (c & 31);
          int e = H.hFs.rowptr[cn[i]];
          const int e1 = H.hFs.rowptr[cn[i] + 1];
          for (int j = 0; j < n2; ++j) {
            if (j > 0 && sc[j] == sc[j - 1]) { mi[size_t(sj[j]) * 32] = mi[size_t(sj[j - 1]) * 32]; continue; }
            while (e < e1 && H.hFs.colind[e] < sc[j]) ++e;
            mi[size_t(sj[j]) * 32] = (e < e1 && H.hFs.colind[e] == sc[j]) ? e : -1;
          }
        }
      }
    }
    H.d_mapF.upload(map.get(), size_t(H.nc_pad / 32) * gsz);
  }
  phase("uploads, scatter map");
  H.d_vcoords.upload(interleave32<double>(H.h_vcoords, nc, H.nc_pad, nv1 * dim, 0.0));
  H.d_cell_nodes.upload(interleave32<int>(H.h_cell_nodes, nc, H.nc_pad, n2, -1));
  H.d_cell_p.upload(interleave32<int>(H.h_cell_p, nc, H.nc_pad, nv1, 0));
  H.d_step_tensor.alloc(1);
  {
    StepTensor t;
    build_step_tensor(H.h_tab, dim, H.prm.variant != NSB_VARIANT_3D, t);
    NSB_CUDA_SETUP(cudaMemcpy(H.d_step_tensor.p, &t, sizeof(t), cudaMemcpyHostToDevice));
  }
  // vectors
  const size_t nl = size_t(H.n_local());
  H.d_sol.alloc(nl); H.d_rhs.alloc(nl);
  H.d_sol.zero(); H.d_rhs.zero();
  const size_t nun = size_t(dim) * H.n_nodes;
  H.d_D.alloc(nun); H.d_Dinv.alloc(nun); H.d_negDinv.alloc(nun);
  H.d_D.zero(); H.d_Dinv.zero(); H.d_negDinv.zero();
  H.d_massdiag.alloc(H.n_nodes_owned); H.d_masslump.alloc(H.n_nodes_owned);
  H.d_neumann.alloc(size_t(H.nu_owned()));
  H.d_neumann.zero();
  // ILU(0) schedules (static pattern => symbolic work once)
  ilu_reset_graphs(H); // schedules are rebuilt: drop graphs captured for the old ones
  phase("mesh arrays, vectors");
  stream_build_spmv(H);
  if (dim == 3) {
    sell_build(H.hFs.rowptr, H.hFs.colind, {}, {0, H.hFs.n_rows}, 2048, sell_lanes_for(H.hFs.n_rows), H.sellF);
    H.d_xpad.alloc(size_t(4) * H.n_nodes);
    H.d_xpad.zero();
  }
  H.sellF_dirty = true;
  phase("SpMV formats (SELL, stream)");
  std::vector<double> xyz_n, xyz_p; // support points of the P2 nodes / pressure vertices (subdomain ordering)
  const int ord_s = H.prm.ilu_ordering_schur < 0 ? H.prm.ilu_ordering : H.prm.ilu_ordering_schur;
  if (H.prm.ilu_ordering == 3 || ord_s == 3) {
    xyz_n.assign(size_t(H.n_nodes) * dim, 0.0);
    xyz_p.assign(size_t(H.n_p) * dim, 0.0);
    for (int64_t c = 0; c < nc; ++c) {
      const double *X = &H.h_vcoords[c * nv1 * dim];
      for (int a = 0; a < n2; ++a) {
        const int node = H.h_cell_nodes[c * n2 + a];
        const int va = a < nv1 ? a : kEdgeA[a - nv1], vb = a < nv1 ? a : kEdgeB[a - nv1];
        for (int d = 0; d < dim; ++d) xyz_n[size_t(node) * dim + d] = 0.5 * (X[va * dim + d] + X[vb * dim + d]);
      }
      for (int v = 0; v < nv1; ++v)
        for (int d = 0; d < dim; ++d) xyz_p[size_t(H.h_cell_p[c * nv1 + v]) * dim + d] = X[v * dim + d];
    }
  }
  ilu_build(H, H.iluF, H.hFs, H.n_nodes_owned, dim, H.prm.ilu_ordering, xyz_n.empty() ? nullptr : xyz_n.data(), dim);
  phase("ILU schedule F");
  ilu_build(H, H.iluS, H.hS, H.n_p_owned, 1, ord_s, xyz_p.empty() ? nullptr : xyz_p.data(), dim);
  phase("ILU schedule S");
  solver_alloc(H);
  NSB_CUDA_SETUP(cudaDeviceSynchronize());
  if (g_dry.on) dry_record_host_state(H);
  H.finalized = true;
  H.assembled = H.prec_ready = false;
}

extern "C" int nsb_finalize_setup(nsb_handle h)
{
  return guarded(h, [&](Handle &H) { finalize_setup(H); });
}

// ---- reference-layout views ------------------------------------------------------------------
static void ref_pattern(Handle &H, int blk, std::vector<int> &rowptr, std::vector<int> &colind)
{
  const int dim = H.dim;
  rowptr.clear(); colind.clear();
  rowptr.push_back(0);
  if (blk == NSB_BLK_F) {
    for (int i = 0; i < H.hFs.n_rows; ++i)
      for (int c = 0; c < dim; ++c) {
        for (int k = H.hFs.rowptr[i]; k < H.hFs.rowptr[i + 1]; ++k)
          for (int d = 0; d < dim; ++d) colind.push_back(dim * H.hFs.colind[k] + d);
        rowptr.push_back(int(colind.size()));
      }
  } else if (blk == NSB_BLK_BT) {
    for (int i = 0; i < H.n_nodes_owned; ++i)
      for (int c = 0; c < dim; ++c) {
        for (int k = H.hBt.rowptr[i]; k < H.hBt.rowptr[i + 1]; ++k) colind.push_back(H.hBt.colind[k]);
        rowptr.push_back(int(colind.size()));
      }
  } else if (blk == NSB_BLK_B) {
    for (int i = 0; i < H.hB.n_rows; ++i) {
      for (int k = H.hB.rowptr[i]; k < H.hB.rowptr[i + 1]; ++k)
        for (int d = 0; d < dim; ++d) colind.push_back(dim * H.hB.colind[k] + d);
      rowptr.push_back(int(colind.size()));
    }
  } else if (blk == NSB_BLK_MP || blk == NSB_BLK_S) {
    const Csr &A = blk == NSB_BLK_MP ? H.hMp : H.hS;
    rowptr = A.rowptr;
    colind = A.colind;
  } else
    throw ArgError("unknown block id");
}

extern "C" int nsb_get_pattern_size(nsb_handle h, int blk, int32_t *n_rows, int64_t *nnz)
{
  return guarded(h, [&](Handle &H) {
    if (!H.finalized) throw StateError("setup not finalized");
    std::vector<int> rp, ci;
    ref_pattern(H, blk, rp, ci);
    if (n_rows) *n_rows = int(rp.size()) - 1;
    if (nnz) *nnz = int64_t(ci.size());
  });
}

extern "C" int nsb_get_pattern(nsb_handle h, int blk, int32_t *rowptr, int32_t *colind)
{
  return guarded(h, [&](Handle &H) {
    if (!H.finalized) throw StateError("setup not finalized");
    std::vector<int> rp, ci;
    ref_pattern(H, blk, rp, ci);
    std::memcpy(rowptr, rp.data(), sizeof(int) * rp.size());
    std::memcpy(colind, ci.data(), sizeof(int) * ci.size());
  });
}

extern "C" int nsb_check_pattern(nsb_handle h, int blk, const int32_t *rowptr, const int32_t *colind)
{
  return guarded(h, [&](Handle &H) {
    if (!H.finalized) throw StateError("setup not finalized");
    std::vector<int> rp, ci;
    ref_pattern(H, blk, rp, ci);
    if (std::memcmp(rowptr, rp.data(), sizeof(int) * rp.size()) != 0)
      throw ArgError("nsb_check_pattern: row pointers differ from the derived sparsity pattern");
    if (std::memcmp(colind, ci.data(), sizeof(int) * ci.size()) != 0)
      throw ArgError("nsb_check_pattern: column indices differ from the derived sparsity pattern");
  });
}

extern "C" int nsb_p2p_export(nsb_handle h, void *handle64)
{
  return guarded(h, [&](Handle &H) {
    if (!handle64) throw ArgError("nsb_p2p_export: null output");
    halo_p2p_export(H, handle64);
  });
}
extern "C" int nsb_p2p_attach(nsb_handle h, const void *handles)
{
  return guarded(h, [&](Handle &H) {
    if (!handles) throw ArgError("nsb_p2p_attach: null handles");
    halo_p2p_attach(H, handles);
  });
}

extern "C" int nsb_set_halo(nsb_handle h, int32_t n_nb, const int32_t *nb_ranks, const int32_t *send_node_ptr,
                            const int32_t *send_node_idx, const int32_t *recv_node_cnt, const int32_t *send_p_ptr,
                            const int32_t *send_p_idx, const int32_t *recv_p_cnt)
{
  return guarded(h, [&](Handle &H) {
    if (!H.have_mesh) throw StateError("nsb_set_halo: set the mesh first");
    halo_set_plan(H, n_nb, nb_ranks, send_node_ptr, send_node_idx, recv_node_cnt, send_p_ptr, send_p_idx, recv_p_cnt);
  });
}

// ---- boundary data ---------------------------------------------------------------------------
extern "C" int nsb_set_dirichlet(nsb_handle h, int32_t n_rows, const int32_t *rows)
{
  return guarded(h, [&](Handle &H) {
    if (!H.finalized) throw StateError("setup not finalized");
    const int dim = H.dim;
    std::vector<int> slot_of_node(H.n_nodes, -1);
    H.h_dir_nodes.clear();
    H.h_dir_rows.assign(rows, rows + n_rows);
    H.h_dir_slot.resize(n_rows);
    std::vector<int> seen;
    for (int k = 0; k < n_rows; ++k) {
      const int r = rows[k];
      if (r < 0 || r >= dim * H.n_nodes) throw ArgError("nsb_set_dirichlet: only velocity DoFs can be constrained");
      const int node = r / dim, c = r % dim;
      if (slot_of_node[node] < 0) {
        slot_of_node[node] = int(H.h_dir_nodes.size());
        H.h_dir_nodes.push_back(node);
        seen.push_back(0);
      }
      seen[slot_of_node[node]] |= 1 << c;
      H.h_dir_slot[k] = slot_of_node[node] * dim + c;
    }
    for (int m : seen)
      if (m != (1 << dim) - 1)
        throw ArgError("nsb_set_dirichlet: all velocity components of a node must be constrained together");
    // owned nodes first so that the kernels touching F / rhs see a prefix; ghost nodes only clear Bt rows
    H.d_dir_nodes.upload(H.h_dir_nodes);
    H.d_dir_vals.alloc(std::max<size_t>(1, H.h_dir_nodes.size() * dim));
    H.d_dir_vals.zero();
    H.h_dir_vals.assign(H.h_dir_nodes.size() * dim, 0.0);
    NSB_CUDA(cudaDeviceSynchronize()); // uploads / memsets above ran on the default stream, kernels use H.stream
  });
}

extern "C" int nsb_set_dirichlet_values(nsb_handle h, const double *values)
{
  return guarded(h, [&](Handle &H) {
    if (H.h_dir_rows.empty()) return;
    for (size_t k = 0; k < H.h_dir_rows.size(); ++k) H.h_dir_vals[H.h_dir_slot[k]] = values[k];
    NSB_CUDA(cudaMemcpyAsync(H.d_dir_vals.p, H.h_dir_vals.data(), sizeof(double) * H.h_dir_vals.size(),
                             cudaMemcpyHostToDevice, H.stream));
    sync(H);
  });
}

extern "C" int nsb_set_neumann_rhs(nsb_handle h, const double *rhs_u)
{
  return guarded(h, [&](Handle &H) {
    if (!H.finalized) throw StateError("setup not finalized");
    if (!rhs_u) { H.have_neumann = false; return; }
    h2d(H, H.d_neumann.p, rhs_u, sizeof(double) * H.nu_owned());
    H.have_neumann = true;
  });
}

// ---- CPU-only self check of the subdomain ILU storage (no GPU needed; tests/test_host_cpu.py) ------
namespace nsb {
double sd_debug_check(const Csr &A, const double *xyz, int gdim, const int *leaf_levels, int min_active, int bs, int *stats,
                      int *order_out);
}
extern "C" int nsb_debug_sd_check(int32_t n, const int32_t *rowptr, const int32_t *colind, const double *xyz, int32_t gdim,
                                  const int32_t *leaf_levels, int32_t min_active, int32_t bs, double *rel_err, int32_t *stats,
                                  int32_t *order_out)
{
  try {
    if (n <= 0 || !rowptr || !colind || !rel_err || bs < 1 || bs > 3) return NSB_ERR_ARG;
    Csr A;
    A.n_rows = A.n_cols = n;
    A.rowptr.assign(rowptr, rowptr + n + 1);
    A.colind.assign(colind, colind + rowptr[n]);
    if (!leaf_levels) return NSB_ERR_ARG;
    *rel_err = sd_debug_check(A, xyz, gdim, leaf_levels, min_active, bs, stats, order_out);
    return NSB_OK;
  } catch (const std::exception &) {
    return NSB_ERR_STATE;
  }
}

namespace nsb {
double bsell_debug_check(const Csr &A, int bs, int xcap, int *stats, int *order_out);
}
extern "C" int nsb_debug_bsell_check(int32_t n, const int32_t *rowptr, const int32_t *colind, int32_t bs, int32_t xcap,
                                     double *rel_err, int32_t *stats, int32_t *order_out)
{
  try {
    if (n <= 0 || !rowptr || !colind || !rel_err || bs < 1 || bs > 3 || xcap < 0) return NSB_ERR_ARG;
    Csr A;
    A.n_rows = A.n_cols = n;
    A.rowptr.assign(rowptr, rowptr + n + 1);
    A.colind.assign(colind, colind + rowptr[n]);
    *rel_err = bsell_debug_check(A, bs, xcap, stats, order_out);
    return NSB_OK;
  } catch (const std::exception &) {
    return NSB_ERR_STATE;
  }
}

namespace nsb {
double sell_debug_check(const Csr &A, int bs, int lanes, int window, int *stats, int *order_out);
}
extern "C" int nsb_debug_sell_check(int32_t n, const int32_t *rowptr, const int32_t *colind, int32_t bs, int32_t lanes,
                                    int32_t window, double *rel_err, int32_t *stats, int32_t *order_out)
{
  try {
    if (n <= 0 || !rowptr || !colind || !rel_err || bs < 1 || bs > 3 || (lanes != 1 && lanes != 4) || window < 1) return NSB_ERR_ARG;
    Csr A;
    A.n_rows = A.n_cols = n;
    A.rowptr.assign(rowptr, rowptr + n + 1);
    A.colind.assign(colind, colind + rowptr[n]);
    *rel_err = sell_debug_check(A, bs, lanes, window, stats, order_out);
    return NSB_OK;
  } catch (const std::exception &) {
    return NSB_ERR_STATE;
  }
}

// ---- CPU-only fingerprint of the host side of setup (no GPU needed; tests/test_setup_fingerprint.py) ------
// Runs set_mesh_host + finalize_setup on a handle that owns no device state, with the setup dry run on: every array
// setup would upload is hashed in upload order.  Test infrastructure; computes nothing and is not reachable from the
// product entry points.
extern "C" int nsb_debug_setup_fingerprint(int32_t dim, int32_t n_cells, const double *vertex_coords, const int32_t *cell_dofs,
                                           int32_t n_u, int32_t n_p, int32_t n_u_owned, int32_t n_p_owned,
                                           int32_t ilu_ordering, int32_t ilu_ordering_schur, uint64_t *hashes, int32_t cap,
                                           int32_t *n_hashes)
{
  if ((dim != 2 && dim != 3) || !n_hashes || (cap > 0 && !hashes) || ilu_ordering < 0 || ilu_ordering > 3 ||
      ilu_ordering_schur < -1 || ilu_ordering_schur > 3)
    return NSB_ERR_ARG;
  if (g_dry.on) return NSB_ERR_STATE; // one dry run at a time (the log is process-wide)
  int rc = NSB_OK;
  nsb_handle_s *h = nullptr;
  try {
    h = new nsb_handle_s();
    Handle &H = h->H;
    H.dim = dim; H.n2 = n2_of(dim); H.nv1 = dim + 1; H.dpc = dpc_of(dim);
    nsb_default_params(&H.prm, dim == 2 ? NSB_VARIANT_2D : NSB_VARIANT_3D);
    H.prm.ilu_ordering = ilu_ordering;
    H.prm.ilu_ordering_schur = ilu_ordering_schur;
    g_dry.log.clear();
    g_dry.on = true;
    set_mesh_host(H, n_cells, vertex_coords, cell_dofs, n_u, n_p, n_u_owned, n_p_owned);
    std::memset(&H.h_tab, 0, sizeof(H.h_tab)); // the quadrature tables do not enter any static structure
    H.have_quad = true;
    finalize_setup(H);
  } catch (const ArgError &) { rc = NSB_ERR_ARG;
  } catch (const std::exception &) { rc = NSB_ERR_STATE; }
  delete h; // buffers were never allocated: nothing to free on a device
  g_dry.on = false;
  *n_hashes = int32_t(g_dry.log.size());
  for (int32_t k = 0; k < cap && k < *n_hashes; ++k) hashes[k] = g_dry.log[k];
  g_dry.log.clear();
  return rc;
}

// ---- drag / lift on the device (kernels_post.cu) ---------------------------------------------
extern "C" int nsb_set_force_faces(nsb_handle h, int32_t n_faces, const int32_t *face_cell, const int32_t *face_local,
                                   int32_t n_q, const double *xi, const double *w)
{
  return guarded(h, [&](Handle &H) {
    if (!H.have_mesh) throw StateError("nsb_set_force_faces before nsb_set_mesh");
    if (n_faces > 0 && (!face_cell || !face_local)) throw ArgError("nsb_set_force_faces: null face arrays");
    if (!xi || !w) throw ArgError("nsb_set_force_faces: null quadrature");
    force_faces_set(H, n_faces, face_cell, face_local, n_q, xi, w);
  });
}
extern "C" int nsb_compute_forces(nsb_handle h, double rho, double *out)
{
  return guarded(h, [&](Handle &H) {
    if (!H.finalized) throw StateError("setup not finalized");
    if (!out) throw ArgError("nsb_compute_forces: null output");
    force_faces_compute(H, rho, out);
  });
}

// ---- vectors: caller layout [u (owned, ghost) | p (owned, ghost)] <-> device layout -----------
static void to_device_layout(Handle &H, const double *x, std::vector<double> &out)
{
  const int dim = H.dim, nuo = H.nu_owned(), nu = dim * H.n_nodes, npo = H.n_p_owned, np = H.n_p;
  out.resize(H.n_local());
  std::memcpy(&out[0], x, sizeof(double) * nuo);
  std::memcpy(&out[nuo], x + nu, sizeof(double) * npo);
  std::memcpy(&out[nuo + npo], x + nuo, sizeof(double) * (nu - nuo));
  std::memcpy(&out[nuo + npo + (nu - nuo)], x + nu + npo, sizeof(double) * (np - npo));
}
static void from_device_layout(Handle &H, const std::vector<double> &in, double *x)
{
  const int dim = H.dim, nuo = H.nu_owned(), nu = dim * H.n_nodes, npo = H.n_p_owned, np = H.n_p;
  std::memcpy(x, &in[0], sizeof(double) * nuo);
  std::memcpy(x + nu, &in[nuo], sizeof(double) * npo);
  std::memcpy(x + nuo, &in[nuo + npo], sizeof(double) * (nu - nuo));
  std::memcpy(x + nu + npo, &in[nuo + npo + (nu - nuo)], sizeof(double) * (np - npo));
}
static void upload_vec(Handle &H, const double *x, double *dev)
{
  if (H.n_nodes == H.n_nodes_owned && H.n_p == H.n_p_owned) {
    NSB_CUDA(cudaMemcpyAsync(dev, x, sizeof(double) * H.n_local(), cudaMemcpyHostToDevice, H.stream));
    sync(H);
    return;
  }
  std::vector<double> t;
  to_device_layout(H, x, t);
  h2d(H, dev, t.data(), sizeof(double) * t.size());
}
static void download_vec(Handle &H, const double *dev, double *x)
{
  if (H.n_nodes == H.n_nodes_owned && H.n_p == H.n_p_owned) {
    NSB_CUDA(cudaMemcpyAsync(x, dev, sizeof(double) * H.n_local(), cudaMemcpyDeviceToHost, H.stream));
    sync(H);
    return;
  }
  std::vector<double> t(H.n_local());
  sync(H);
  NSB_CUDA(cudaMemcpy(t.data(), dev, sizeof(double) * t.size(), cudaMemcpyDeviceToHost));
  from_device_layout(H, t, x);
}

extern "C" int nsb_set_solution(nsb_handle h, const double *x)
{
  return guarded(h, [&](Handle &H) {
    if (!H.finalized) throw StateError("setup not finalized");
    upload_vec(H, x, H.d_sol.p);
  });
}
extern "C" int nsb_get_solution(nsb_handle h, double *x)
{
  return guarded(h, [&](Handle &H) {
    if (!H.finalized) throw StateError("setup not finalized");
    if (H.nranks > 1) { // solution = solution_owned: ghost import (NavierStokes2D.cpp:637)
      halo_exchange_u(H, H.d_sol.p, H.ghost_off_u());
      halo_exchange_p(H, H.d_sol.p + H.nu_owned(), H.ghost_off_p());
    }
    download_vec(H, H.d_sol.p, x);
  });
}
extern "C" int nsb_allreduce_sum(nsb_handle h, double *vals, int32_t n)
{
  return guarded(h, [&](Handle &H) {
    if (!H.finalized) throw StateError("setup not finalized");
    if (!vals || n < 1 || n > 64) throw ArgError("nsb_allreduce_sum: 1 <= n <= 64 values required");
    if (H.nranks <= 1) return;
    double *dev = H.ws->scal.p + 192; // device scalars not used by the solvers
    h2d(H, dev, vals, sizeof(double) * size_t(n));
    halo_allreduce(H, dev, n);
    reduce_fetch(H, dev, n, vals);
  });
}
extern "C" int nsb_get_rhs(nsb_handle h, double *x)
{
  return guarded(h, [&](Handle &H) {
    if (!H.finalized) throw StateError("setup not finalized");
    download_vec(H, H.d_rhs.p, x);
  });
}

// ---- the hot path -------------------------------------------------------------------------------
static void refresh_ghosts(Handle &H)
{ // solution = solution_owned (ghost import, NavierStokes2D.cpp:637,709)
  if (H.nranks > 1) {
    halo_exchange_u(H, H.d_sol.p, H.ghost_off_u());
    halo_exchange_p(H, H.d_sol.p + H.nu_owned(), H.ghost_off_p());
  }
}

static void do_assemble_first(Handle &H)
{
  if (!H.finalized) throw StateError("setup not finalized");
  refresh_ghosts(H);
  launch_assemble_first(H);
  mass_rows(H);
  launch_apply_dirichlet(H, true);
  H.sellF_dirty = true;
  H.assembled = true;
  H.prec_ready = false;
}
static void do_assemble_step(Handle &H)
{
  if (!H.assembled) throw StateError("nsb_assemble_step before nsb_assemble_first");
  refresh_ghosts(H);
  launch_assemble_step(H, H.Fs.val.p);
  launch_apply_dirichlet(H, false);
  H.sellF_dirty = true;
  H.prec_ready = false;
}
static void do_solve(Handle &H, int32_t *outer_iters, double *t_prec, double *t_solve)
{
  if (!H.assembled) throw StateError("nsb_solve_step before assembly");
  H.n_inner_F = H.n_inner_S = H.n_F_solves = H.n_S_solves = H.n_vmult = 0;
  H.cnt_spmv_F = H.cnt_spmv_S = H.cnt_spmv_B = H.cnt_spmv_Bt = H.cnt_ilu_F = H.cnt_ilu_S = H.cnt_dot = H.cnt_sync = 0;
  sync(H);
  const auto t0 = std::chrono::steady_clock::now();
  precond_init(H);
  sync(H);
  const auto t1 = std::chrono::steady_clock::now();
  const int rc = solve_outer(H);
  sync(H);
  const auto t2 = std::chrono::steady_clock::now();
  H.t_prec_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  H.t_solve_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
  if (outer_iters) *outer_iters = H.last_outer;
  if (t_prec) *t_prec = H.t_prec_ms * 1e-3;
  if (t_solve) *t_solve = H.t_solve_ms * 1e-3;
  if (rc != 0) throw NoConvergence("SolverControl::NoConvergence: outer GMRES did not reach the tolerance");
}

extern "C" int nsb_assemble_first(nsb_handle h, double)
{
  return guarded(h, [&](Handle &H) { do_assemble_first(H); });
}
extern "C" int nsb_assemble_step(nsb_handle h, double)
{
  return guarded(h, [&](Handle &H) { do_assemble_step(H); });
}
extern "C" int nsb_solve_step(nsb_handle h, int32_t *outer_iters, double *t_prec, double *t_solve)
{
  return guarded(h, [&](Handle &H) { do_solve(H, outer_iters, t_prec, t_solve); });
}

extern "C" int nsb_step_host(nsb_handle h, int first, double, const double *dirichlet_values, double *solution_out,
                             int32_t *outer_iters)
{
  return guarded(h, [&](Handle &H) {
    if (dirichlet_values && !H.h_dir_rows.empty()) {
      for (size_t k = 0; k < H.h_dir_rows.size(); ++k) H.h_dir_vals[H.h_dir_slot[k]] = dirichlet_values[k];
      NSB_CUDA(cudaMemcpyAsync(H.d_dir_vals.p, H.h_dir_vals.data(), sizeof(double) * H.h_dir_vals.size(),
                               cudaMemcpyHostToDevice, H.stream));
    }
    // device time of the step proper (copies excluded), accumulated for bench.py's `value`
    NSB_CUDA(cudaEventRecord(H.ev_s0, H.stream));
    if (first) do_assemble_first(H); else do_assemble_step(H);
    do_solve(H, outer_iters, nullptr, nullptr);
    NSB_CUDA(cudaEventRecord(H.ev_s1, H.stream));
    if (solution_out) download_vec(H, H.d_sol.p, solution_out);
    NSB_CUDA(cudaEventSynchronize(H.ev_s1));
    float ms = 0.f;
    NSB_CUDA(cudaEventElapsedTime(&ms, H.ev_s0, H.ev_s1));
    H.t_step_dev_ms += double(ms);
  });
}

// ---- parity harness -----------------------------------------------------------------------------
extern "C" int nsb_get_matrix_values(nsb_handle h, int mat, int blk, double *vals)
{
  return guarded(h, [&](Handle &H) {
    if (!H.assembled) throw StateError("nothing assembled yet");
    const int dim = H.dim;
    sync(H);
    if (blk == NSB_BLK_F) {
      std::vector<double> v(H.Fs.nnz);
      const double *src = nullptr;
      if (mat == NSB_MAT_SYSTEM) src = H.Fs.val.p;
      else if (mat == NSB_MAT_MASS) src = H.d_M.p;
      else if (mat == NSB_MAT_STIFFNESS) src = H.d_A.p;
      else if (mat == NSB_MAT_CONVECTION) {
        // convection_matrix of the last assemble_time_step, rebuilt from the current solution:
        // only meaningful between nsb_assemble_* and nsb_solve_step.
        H.d_C.zero(H.stream);
        double *saved_rhs = H.d_rhs.p;
        H.d_rhs.p = H.ws->prec_out.p; // dummy rhs target
        launch_assemble_step(H, H.d_C.p);
        H.d_rhs.p = saved_rhs;
        sync(H);
        src = H.d_C.p;
      } else throw ArgError("unknown matrix id");
      NSB_CUDA(cudaMemcpy(v.data(), src, sizeof(double) * v.size(), cudaMemcpyDeviceToHost));
      size_t o = 0;
      for (int i = 0; i < H.hFs.n_rows; ++i)
        for (int c = 0; c < dim; ++c)
          for (int k = H.hFs.rowptr[i]; k < H.hFs.rowptr[i + 1]; ++k)
            for (int d = 0; d < dim; ++d) vals[o++] = (c == d) ? v[k] : 0.0;
    } else if (blk == NSB_BLK_BT || blk == NSB_BLK_B) {
      const DevCsr &A = blk == NSB_BLK_BT ? H.Bt : H.B;
      const Csr &P = blk == NSB_BLK_BT ? H.hBt : H.hB;
      std::vector<double> v(size_t(A.nnz) * dim, 0.0);
      if (mat == NSB_MAT_SYSTEM) NSB_CUDA(cudaMemcpy(v.data(), A.val.p, sizeof(double) * v.size(), cudaMemcpyDeviceToHost));
      size_t o = 0;
      if (blk == NSB_BLK_BT) {
        for (int i = 0; i < H.n_nodes_owned; ++i)
          for (int c = 0; c < dim; ++c)
            for (int k = P.rowptr[i]; k < P.rowptr[i + 1]; ++k) vals[o++] = v[size_t(k) * dim + c];
      } else {
        for (int i = 0; i < P.n_rows; ++i)
          for (int k = P.rowptr[i]; k < P.rowptr[i + 1]; ++k)
            for (int d = 0; d < dim; ++d) vals[o++] = v[size_t(k) * dim + d];
      }
    } else if (blk == NSB_BLK_MP) {
      NSB_CUDA(cudaMemcpy(vals, H.Mp.val.p, sizeof(double) * H.Mp.nnz, cudaMemcpyDeviceToHost));
    } else if (blk == NSB_BLK_S) {
      NSB_CUDA(cudaMemcpy(vals, H.S.val.p, sizeof(double) * H.S.nnz, cudaMemcpyDeviceToHost));
    } else
      throw ArgError("unknown block id");
  });
}

extern "C" int nsb_get_ilu_order(nsb_handle h, int which, int32_t *order)
{
  return guarded(h, [&](Handle &H) {
    if (!H.finalized) throw StateError("setup not finalized");
    const DevIlu &ilu = which == 0 ? H.iluF : H.iluS;
    std::memcpy(order, ilu.h_order.data(), sizeof(int) * ilu.h_order.size());
  });
}

extern "C" int nsb_get_schur_values(nsb_handle h, double *vals)
{
  return guarded(h, [&](Handle &H) {
    if (!H.prec_ready) throw StateError("preconditioner not initialised");
    sync(H);
    NSB_CUDA(cudaMemcpy(vals, H.S.val.p, sizeof(double) * H.S.nnz, cudaMemcpyDeviceToHost));
  });
}

extern "C" int nsb_op_system_vmult(nsb_handle h, const double *x, double *y)
{
  return guarded(h, [&](Handle &H) {
    if (!H.assembled) throw StateError("nothing assembled yet");
    upload_vec(H, x, H.ws->prec_in.p);
    system_vmult(H, H.ws->prec_in.p, H.ws->prec_out.p);
    download_vec(H, H.ws->prec_out.p, y);
  });
}

extern "C" int nsb_op_block_vmult(nsb_handle h, int blk, const double *x, double *y)
{
  return guarded(h, [&](Handle &H) {
    if (!H.assembled) throw StateError("nothing assembled yet");
    if (H.nranks > 1) throw StateError("nsb_op_block_vmult is a single-rank harness entry point");
    const int nu = H.nu_owned(), np = H.n_p_owned;
    double *in = H.ws->prec_in.p, *out = H.ws->prec_out.p;
    const int nin = (blk == NSB_BLK_F || blk == NSB_BLK_B) ? nu : np;
    const int nout = (blk == NSB_BLK_F || blk == NSB_BLK_BT) ? nu : np;
    h2d(H, in, x, sizeof(double) * nin);
    if (blk == NSB_BLK_F) spmv_F(H, in, 0, nullptr, 0, out);
    else if (blk == NSB_BLK_BT) spmv_Bt(H, in, 0, out);
    else if (blk == NSB_BLK_B) spmv_B(H, in, 0, out);
    else if (blk == NSB_BLK_S) { if (!H.prec_ready) throw StateError("preconditioner not initialised"); spmv_S(H, in, 0, out); }
    else throw ArgError("unknown block id");
    sync(H);
    NSB_CUDA(cudaMemcpy(y, out, sizeof(double) * nout, cudaMemcpyDeviceToHost));
  });
}

extern "C" int nsb_op_precond_init(nsb_handle h)
{
  return guarded(h, [&](Handle &H) {
    if (!H.assembled) throw StateError("nothing assembled yet");
    precond_init(H);
    sync(H);
  });
}

extern "C" int nsb_op_ilu_apply(nsb_handle h, int which, const double *x, double *y)
{
  return guarded(h, [&](Handle &H) {
    if (!H.prec_ready) throw StateError("preconditioner not initialised");
    DevIlu &ilu = which == 0 ? H.iluF : H.iluS;
    const int n = ilu.n * ilu.bs_rhs;
    double *in = H.ws->prec_in.p, *out = H.ws->prec_out.p;
    h2d(H, in, x, sizeof(double) * n);
    ilu_solve(H, ilu, in, out);
    sync(H);
    NSB_CUDA(cudaMemcpy(y, out, sizeof(double) * n, cudaMemcpyDeviceToHost));
  });
}

extern "C" int nsb_op_precond_vmult(nsb_handle h, const double *src, const double *dst_in, double *dst)
{
  return guarded(h, [&](Handle &H) {
    if (!H.prec_ready) throw StateError("preconditioner not initialised");
    H.n_inner_F = H.n_inner_S = H.n_F_solves = H.n_S_solves = 0;
    upload_vec(H, src, H.ws->prec_in.p);
    if (dst_in) upload_vec(H, dst_in, H.ws->prec_out.p);
    else H.ws->prec_out.zero(H.stream);
    precond_vmult(H, H.ws->prec_in.p, H.ws->prec_out.p);
    download_vec(H, H.ws->prec_out.p, dst);
  });
}

// ---- measurement --------------------------------------------------------------------------------
extern "C" double nsb_stat(nsb_handle h, const char *name)
{
  if (!h || !name) return -1;
  Handle &H = h->H;
  const std::string n(name);
  if (n == "n_inner_F") return double(H.n_inner_F);
  if (n == "n_inner_S") return double(H.n_inner_S);
  if (n == "n_F_solves") return double(H.n_F_solves);
  if (n == "n_S_solves") return double(H.n_S_solves);
  if (n == "n_vmult") return double(H.n_vmult);
  if (n == "last_outer") return double(H.last_outer);
  if (n == "last_res") return H.last_res;
  if (n == "t_prec_ms") return H.t_prec_ms;
  if (n == "t_solve_ms") return H.t_solve_ms;
  if (n == "t_step_dev_ms") return H.t_step_dev_ms;                                   // accumulated by nsb_step_host
  if (n == "t_step_dev_ms_reset") { const double v = H.t_step_dev_ms; H.t_step_dev_ms = 0; return v; }
  if (n == "nnz_Fs") return double(H.Fs.nnz);
  if (n == "nnz_B") return double(H.B.nnz);
  if (n == "nnz_Bt") return double(H.Bt.nnz);
  if (n == "nnz_S") return double(H.S.nnz);
  if (n == "n_nodes") return double(H.n_nodes);
  if (n == "n_p") return double(H.n_p);
  if (n == "n_cells") return double(H.nc);
  if (n == "cnt_spmv_F") return double(H.cnt_spmv_F);
  if (n == "cnt_spmv_S") return double(H.cnt_spmv_S);
  if (n == "cnt_spmv_B") return double(H.cnt_spmv_B);
  if (n == "cnt_spmv_Bt") return double(H.cnt_spmv_Bt);
  if (n == "cnt_ilu_F") return double(H.cnt_ilu_F);
  if (n == "cnt_ilu_S") return double(H.cnt_ilu_S);
  if (n == "cnt_dot") return double(H.cnt_dot);
  if (n == "cnt_sync") return double(H.cnt_sync);
  if (n == "nnz_iluF") return double(H.iluF.nnz);
  if (n == "nnz_iluS") return double(H.iluS.nnz);
  if (n == "p2p") return halo_is_p2p(H) ? 1.0 : 0.0;
  // dependent launches of one triangular solve (forward or backward)
  if (n == "sweeps_F") return double(H.iluF.colour_ptr.size() > 1 ? H.iluF.colour_ptr.size() - 1 : H.iluF.lvl_ptr_f.size() - 1);
  if (n == "sweeps_S") return double(H.iluS.colour_ptr.size() > 1 ? H.iluS.colour_ptr.size() - 1 : H.iluS.lvl_ptr_f.size() - 1);
  if (n == "ilu_blocks_F") return double(H.iluF.bL.n_blocks);
  if (n == "ilu_blocks_S") return double(H.iluS.bL.n_blocks);
  if (n == "levels_F_fwd") return double(H.iluF.lvl_ptr_f.size()) - 1;
  if (n == "levels_F_bwd") return double(H.iluF.lvl_ptr_b.size()) - 1;
  if (n == "levels_S_fwd") return double(H.iluS.lvl_ptr_f.size()) - 1;
  if (n == "levels_S_bwd") return double(H.iluS.lvl_ptr_b.size()) - 1;
  return -1;
}

extern "C" int nsb_timer_mark(nsb_handle h, int which)
{
  return guarded(h, [&](Handle &H) { NSB_CUDA(cudaEventRecord(which == 0 ? H.ev0 : H.ev1, H.stream)); });
}
extern "C" int nsb_timer_elapsed_ms(nsb_handle h, double *ms)
{
  return guarded(h, [&](Handle &H) {
    NSB_CUDA(cudaEventSynchronize(H.ev1));
    float f = 0;
    NSB_CUDA(cudaEventElapsedTime(&f, H.ev0, H.ev1));
    if (ms) *ms = f;
  });
}

extern "C" int64_t nsb_launch_count(nsb_handle h, int reset)
{
  if (!h) return -1;
  const int64_t v = h->H.launches;
  if (reset) h->H.launches = 0;
  return v;
}

extern "C" int nsb_bench_kernel(nsb_handle h, const char *which, int iters, int flush, double *ms_per_launch,
                                double *bytes_per_launch)
{
  return guarded(h, [&](Handle &H) {
    if (!H.assembled) throw StateError("nothing assembled yet");
    const std::string w(which ? which : "");
    const int dim = H.dim;
    double *a = H.ws->prec_in.p, *b = H.ws->prec_out.p;
    const double nu = double(H.nu_owned()), np = double(H.n_p_owned);
    double bytes = 0;
    std::function<void()> run;
    if (w == "reset_graphs") { // profiling hook: re-capture the triangular solves under the current environment
      sync(H);
      ilu_reset_graphs(H);
      if (ms_per_launch) *ms_per_launch = 0;
      if (bytes_per_launch) *bytes_per_launch = 0;
      return;
    }
    if (w == "spmv_system") {
      // values + column index of every stored entry, row pointers, x read once, y written once
      bytes = 12.0 * H.Fs.nnz + (8.0 * dim + 4.0) * (H.Bt.nnz + H.B.nnz) + 4.0 * (2.0 * H.n_nodes_owned + np) +
              16.0 * (nu + np);
      run = [&] { system_vmult(H, a, b); };
    } else if (w == "spmv_F") {
      bytes = 12.0 * H.Fs.nnz + 4.0 * H.n_nodes_owned + 16.0 * nu;
      run = [&] { spmv_F(H, a, 0, nullptr, 0, b); };
    } else if (w == "spmv_S") {
      if (!H.prec_ready) throw StateError("preconditioner not initialised");
      bytes = 12.0 * H.S.nnz + 4.0 * np + 16.0 * np;
      run = [&] { spmv_S(H, a, 0, b); };
    } else if (w == "assemble_step") {
      // coordinates + node ids + velocity gather + RMW of the scalar block + rhs RMW
      const int n2 = H.n2, nv1 = H.nv1;
      bytes = double(H.nc) * (nv1 * dim * 8.0 + (n2 + nv1) * 4.0 + n2 * dim * 8.0 + n2 * n2 * 16.0 + n2 * dim * 16.0);
      run = [&] { launch_assemble_step(H, H.Fs.val.p); };
    } else if (w == "ilu_F") {
      if (!H.prec_ready) throw StateError("preconditioner not initialised");
      bytes = 12.0 * H.iluF.nnz + 16.0 * nu;
      run = [&] { ilu_solve(H, H.iluF, a, b); };
    } else if (w == "ilu_S") {
      if (!H.prec_ready) throw StateError("preconditioner not initialised");
      bytes = 12.0 * H.iluS.nnz + 16.0 * np;
      run = [&] { ilu_solve(H, H.iluS, a, b); };
    } else if (w == "dot") {
      bytes = 16.0 * (nu + np);
      run = [&] { vec_dot_dev(H, H.n_owned(), a, b, H.ws->scal.p + 220); };
    } else if (w == "axpy") {
      bytes = 24.0 * (nu + np);
      run = [&] { vec_axpy(H, H.n_owned(), 1e-30, a, b); };
    } else if (w == "add_and_dot") {
      // one modified Gram-Schmidt step of the inner GMRES on F: vv -= h v_prev; h' = vv . v_next
      bytes = 32.0 * nu;
      run = [&] {
        vec_add_and_dot_dev(H, H.nu_owned(), b, H.ws->scal.p + 221, -1e-30, a, H.ws->tu[0].p, H.ws->scal.p + 220);
      };
    } else
      throw ArgError("nsb_bench_kernel: unknown kernel name");
    cudaEvent_t e0, e1;
    NSB_CUDA(cudaEventCreate(&e0));
    NSB_CUDA(cudaEventCreate(&e1));
    double total = 0;
    for (int w_ = 0; w_ < 3; ++w_) run();
    for (int it = 0; it < iters; ++it) {
      if (flush) flush_l2(H);
      NSB_CUDA(cudaEventRecord(e0, H.stream));
      run();
      NSB_CUDA(cudaEventRecord(e1, H.stream));
      NSB_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      NSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_per_launch) *ms_per_launch = total / std::max(1, iters);
    if (bytes_per_launch) *bytes_per_launch = bytes;
    if (w == "assemble_step") { // restore a consistent system matrix
      launch_apply_dirichlet(H, false);
      H.sellF_dirty = true;
      sync(H);
    }
  });
}
