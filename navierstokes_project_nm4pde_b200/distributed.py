"""Subdomain decomposition of the hot path across ranks (one process per GPU).

Replaces what the reference gets from `GridTools::partition_triangulation` +
`parallel::fullydistributed::Triangulation` + Epetra maps (Navier-Stokes/src/NavierStokes2D.cpp:16-19,
71-87): every rank holds the whole mesh (as in the reference, :8-14), cells are partitioned, a DoF
belongs to the lowest rank among its cells, and each rank keeps its owned DoFs plus the ghosts of
a TWO-layer cell halo.  With two layers every rank assembles all rows it needs redundantly (the
rows of B^T at ghost nodes that feed its rows of the Schur product included), so assembly needs no
collective; the per-iteration communication is the ghost exchange before each SpMV and the
all-reduce of each dot product (csrc/halo.cu).  ILU(0) is block-Jacobi per rank exactly like
Ifpack with overlap 0, so results depend on the number of ranks the way `mpirun -n P` does.
"""
from __future__ import annotations

import numpy as np

from .engine import Engine
from .problem import NavierStokes
from .quadrature import gauss_simplex


def _node_cols(dim):
    nv1, ne = dim + 1, (3 if dim == 2 else 6)
    return [v * (dim + 1) for v in range(nv1)] + [nv1 * (dim + 1) + e * dim for e in range(ne)]


def _p_cols(dim):
    return [v * (dim + 1) + dim for v in range(dim + 1)]


def build_local_problem(dim, cell_dofs, cell_coords, n_nodes, n_p, cell_part, nranks, rank):
    """Pure numpy: the local (owned + 2-layer ghost) description of rank `rank`.

    Returns a dict with local cells, local cell_dofs (reference layout, local numbering: owned
    first, then ghosts grouped by owner rank and sorted by global id), global ids of the local
    DoFs and, per neighbour, the global ids of the ghosts this rank needs from it."""
    n_u = dim * n_nodes
    nodes = cell_dofs[:, _node_cols(dim)] // dim
    pv = cell_dofs[:, _p_cols(dim)] - n_u
    n2, nv1 = nodes.shape[1], pv.shape[1]
    node_owner = np.full(n_nodes, nranks, np.int32)
    np.minimum.at(node_owner, nodes.ravel(), np.repeat(cell_part, n2))
    p_owner = np.full(n_p, nranks, np.int32)
    np.minimum.at(p_owner, pv.ravel(), np.repeat(cell_part, nv1))
    # layer 1: cells touching an owned DoF; layer 2: cells touching any DoF of layer 1
    touch1 = (node_owner[nodes] == rank).any(axis=1) | (p_owner[pv] == rank).any(axis=1)
    mark_n = np.zeros(n_nodes, bool)
    mark_p = np.zeros(n_p, bool)
    mark_n[nodes[touch1].ravel()] = True
    mark_p[pv[touch1].ravel()] = True
    touch2 = mark_n[nodes].any(axis=1) | mark_p[pv].any(axis=1)
    cells = np.nonzero(touch2)[0]

    def local_numbering(owner, used_ids):
        used = np.unique(used_ids)
        own = used[owner[used] == rank]
        gh = used[owner[used] != rank]
        gh = gh[np.lexsort((gh, owner[gh]))]  # by owner rank, then global id
        loc = np.concatenate([own, gh])
        g2l = np.full(owner.size, -1, np.int64)
        g2l[loc] = np.arange(loc.size)
        nb, cnt = np.unique(owner[gh], return_counts=True)
        return loc, own.size, g2l, gh, dict(zip(nb.tolist(), cnt.tolist()))

    # every owned DoF must be numbered even if (pathologically) no local cell uses it
    loc_n, n_own_n, g2l_n, gh_n, cnt_n = local_numbering(node_owner, np.concatenate([nodes[cells].ravel(), np.nonzero(node_owner == rank)[0]]))
    loc_p, n_own_p, g2l_p, gh_p, cnt_p = local_numbering(p_owner, np.concatenate([pv[cells].ravel(), np.nonzero(p_owner == rank)[0]]))
    n_u_loc = dim * loc_n.size
    cd = np.empty((cells.size, cell_dofs.shape[1]), np.int32)
    ln, lp = g2l_n[nodes[cells]], g2l_p[pv[cells]]
    for a, col in enumerate(_node_cols(dim)):
        for c in range(dim):
            cd[:, col + c] = dim * ln[:, a] + c
    for v, col in enumerate(_p_cols(dim)):
        cd[:, col] = n_u_loc + lp[:, v]
    neighbours = sorted(set(cnt_n) | set(cnt_p))
    need_nodes = {q: gh_n[node_owner[gh_n] == q] for q in neighbours}
    need_p = {q: gh_p[p_owner[gh_p] == q] for q in neighbours}
    return dict(cells=cells, cell_dofs=cd, cell_coords=np.ascontiguousarray(cell_coords[cells]),
                node_gid=loc_n, p_gid=loc_p, n_nodes_owned=n_own_n, n_p_owned=n_own_p,
                g2l_node=g2l_n, g2l_p=g2l_p, neighbours=neighbours, need_nodes=need_nodes, need_p=need_p,
                node_owner=node_owner, p_owner=p_owner)


def exchange_requests(local, nranks, rank, all_to_all):
    """Tell every owner which of its DoFs this rank needs; returns per neighbour the LOCAL owned
    indices to send (in the requester's ghost order).  `all_to_all(list_of_arrays)` sends array q
    to rank q and returns the list received (plumbing: torch.distributed in production)."""
    out = []
    for kind in ("nodes", "p"):
        need = local["need_" + kind]
        send = [need.get(q, np.zeros(0, np.int64)).astype(np.int64) for q in range(nranks)]
        recv = all_to_all(send)
        g2l = local["g2l_node" if kind == "nodes" else "g2l_p"]
        out.append({q: g2l[recv[q]] for q in range(nranks) if q != rank and recv[q].size})
    return out[0], out[1]


def halo_arrays(local, send_nodes, send_p):
    """Flatten the exchange plan into the arrays of nsb_set_halo."""
    nbs = sorted(set(local["neighbours"]) | set(send_nodes) | set(send_p))
    snp, sni, rnc, spp, spi, rpc = [0], [], [], [0], [], []
    for q in nbs:
        a = send_nodes.get(q, np.zeros(0, np.int64))
        b = send_p.get(q, np.zeros(0, np.int64))
        sni.append(a); snp.append(snp[-1] + a.size)
        spi.append(b); spp.append(spp[-1] + b.size)
        rnc.append(local["need_nodes"].get(q, np.zeros(0)).size)
        rpc.append(local["need_p"].get(q, np.zeros(0)).size)
    cat = lambda l: np.concatenate(l).astype(np.int32) if l else np.zeros(0, np.int32)
    return (np.array(nbs, np.int32), np.array(snp, np.int32), cat(sni), np.array(rnc, np.int32),
            np.array(spp, np.int32), cat(spi), np.array(rpc, np.int32))


def torch_all_to_all(arrays):
    """all_to_all of variable-length int64 arrays through torch.distributed (gloo or nccl)."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    n = dist.get_world_size()
    sizes = torch.tensor([a.size for a in arrays], dtype=torch.int64, device=dev)
    rsizes = torch.empty(n, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rsizes, sizes)
    send = torch.from_numpy(np.concatenate(arrays).astype(np.int64)).to(dev)
    recv = torch.empty(int(rsizes.sum().item()), dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=rsizes.tolist(), input_split_sizes=sizes.tolist())
    out = recv.cpu().numpy()
    offs = np.concatenate([[0], np.cumsum(rsizes.cpu().numpy())])
    return [out[offs[q]:offs[q + 1]] for q in range(n)]


class DistributedNavierStokes(NavierStokes):
    """`NavierStokes` under `mpirun -n P` semantics: one instance per rank / GPU."""

    def setup(self):
        self.setup_host()
        d = self.dofs
        self.N_global = d.N
        cell_part = self.mesh.partition(self.nranks)
        loc = build_local_problem(self.dim, d.cell_dofs(copy=False), d.cell_coords(copy=False), d.n_nodes, d.n_p,
                                  cell_part, self.nranks, self.rank)
        send_nodes, send_p = exchange_requests(loc, self.nranks, self.rank, torch_all_to_all)
        self.local = loc
        dim = self.dim
        nn, npl = loc["node_gid"].size, loc["p_gid"].size
        e = self.engine = Engine(dim, self.device, self.nranks, self.rank, self.unique_id)
        e.default_params(self.variant)
        e.set_mesh(loc["cell_coords"], loc["cell_dofs"], dim * nn, npl, dim * loc["n_nodes_owned"], loc["n_p_owned"])
        e.set_quadrature(*gauss_simplex(dim, self.rule))
        e.set_params(deltat=self.deltat, **self.param_overrides)
        e.set_halo(*halo_arrays(loc, send_nodes, send_p))
        e.finalize()
        self.transport = self._enable_peer_memory(e)
        self.nu = e.params.nu
        # local view of the global Dirichlet list (ghost rows included: their B^T rows are cleared too)
        lnode = loc["g2l_node"][self._dir_nodes]
        keep = lnode >= 0
        self._dir_nodes_global = self._dir_nodes
        self._dir_keep = keep
        self._dir_nodes = lnode[keep]
        self._dir_xyz = self._dir_xyz[keep]
        if self._inlet_mask is not None:
            self._inlet_mask = self._inlet_mask[keep]
        self._dir_rows = (dim * self._dir_nodes[:, None] + np.arange(dim)[None, :]).ravel().astype(np.int32)
        e.set_dirichlet(self._dir_rows)
        if self.variant != "conv":  # compute_forces: obstacle faces of the cells this rank owns (is_locally_owned)
            fc, fl = d.boundary_faces(3)
            mine = cell_part[fc] == self.rank
            g2l_cell = np.full(d.n_cells, -1, np.int64)
            g2l_cell[loc["cells"]] = np.arange(loc["cells"].size)
            e.set_force_faces(g2l_cell[fc[mine]], fl[mine], *gauss_simplex(dim - 1, self.rule))
        # local sizes in the caller layout [u (owned, ghost) | p (owned, ghost)]
        self.n_u, self.n_p, self.N = dim * nn, npl, dim * nn + npl
        return self

    def _enable_peer_memory(self, e):
        """Map every rank's mailbox into every other rank (CUDA IPC over NVLink) so that ghost exchange
        and all-reduce become direct stores into peer HBM (csrc/halo.cu).  NSB_P2P=0 keeps NCCL; a box
        without peer mapping falls back to NCCL on every rank (the decision is made collectively)."""
        import os
        import sys

        import torch.distributed as dist

        # Verified on 2, 4 and 8 GPUs (parity tests at 2 and 4 ranks over both transports; the 19.9 M-DoF bench
        # at 8 ranks: 30.8 ms per outer iteration against 33.2 ms over NCCL, profiles/README.md), so peer memory
        # is the default at any rank count the mailbox supports; NSB_P2P=1 / 0 forces either transport.
        want = os.environ.get("NSB_P2P", "1" if self.nranks <= 16 else "0")
        if want == "0" or dist.get_backend() != "nccl":
            return "nccl"
        try:
            mine = e.p2p_export()
        except Exception as ex:  # noqa: BLE001
            mine = None
            print(f"[rank {self.rank}] peer-memory export failed: {ex}", file=sys.stderr)
        handles = [None] * self.nranks
        dist.all_gather_object(handles, mine)
        ok = all(h is not None for h in handles)
        if ok:
            try:
                e.p2p_attach(handles)
            except Exception as ex:  # noqa: BLE001
                ok = False
                print(f"[rank {self.rank}] peer-memory attach failed: {ex}", file=sys.stderr)
        flags = [None] * self.nranks
        dist.all_gather_object(flags, bool(ok))
        if not all(flags):
            if ok:
                raise RuntimeError("peer-memory transport came up on some ranks only")
            return "nccl"
        return "p2p"

    def initial_condition(self):
        if self.variant != "conv":
            return np.zeros(self.N)  # Functions::ZeroFunction u_0 (NavierStokes2D.hpp:198)
        return self.to_local(self._conv_initial())

    def _conv_initial(self):
        from .problem import exact_solution

        d = self.dofs
        u, _ = exact_solution(d.node_xyz, 0.0)
        _, p = exact_solution(d.p_xyz, 0.0)
        return np.concatenate([u.ravel(), p])

    def to_local(self, x_global):
        d, loc, dim = self.dofs, self.local, self.dim
        xu = x_global[: d.n_u].reshape(-1, dim)[loc["node_gid"]].ravel()
        xp = x_global[d.n_u:][loc["p_gid"]]
        return np.concatenate([xu, xp])

    def neumann_rhs(self, time):
        full = NavierStokes.neumann_rhs(self, time) if self._neu is not None else np.zeros(self.dofs.n_u)
        own = self.local["node_gid"][: self.local["n_nodes_owned"]]
        return full.reshape(-1, self.dim)[own].ravel()

    def owned_solution(self):
        """(global node ids, u[owned, dim], global p ids, p[owned]) of this rank."""
        x = self.engine.get_solution()
        loc, dim = self.local, self.dim
        no, po = loc["n_nodes_owned"], loc["n_p_owned"]
        u = x[: dim * loc["node_gid"].size].reshape(-1, dim)[:no]
        p = x[dim * loc["node_gid"].size:][:po]
        return loc["node_gid"][:no], u, loc["p_gid"][:po], p
