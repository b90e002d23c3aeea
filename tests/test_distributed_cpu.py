"""Host logic of the N > 1 path, exercised with world_size 2 and 4 on the gloo backend (CPU only):
partition, ownership, two-layer ghost cells, halo plan and a halo-exchange-driven SpMV that must
reproduce the serial product."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case_name, out_q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys

        here = os.path.dirname(os.path.abspath(__file__))
        sys.path.insert(0, os.path.dirname(here)); sys.path.insert(0, here)
        import helpers as T
        from navierstokes_project_nm4pde_b200 import distributed as D
        from oracle import ns_ref as R

        case = T.Case(case_name)
        d, dim = case.dofs, case.dim
        cell_part = case.mesh.partition(world)
        loc = D.build_local_problem(dim, d.cell_dofs(), d.cell_coords(), d.n_nodes, d.n_p, cell_part, world, rank)
        send_nodes, send_p = D.exchange_requests(loc, world, rank, D.torch_all_to_all)
        nbs, snp, sni, rnc, spp, spi, rpc = D.halo_arrays(loc, send_nodes, send_p)
        no, po = loc["n_nodes_owned"], loc["n_p_owned"]
        res = {}
        # 1. ownership is a partition of the DoFs
        owned = [None] * world
        dist.all_gather_object(owned, (loc["node_gid"][:no].tolist(), loc["p_gid"][:po].tolist()))
        alln = np.concatenate([np.array(o[0]) for o in owned]); allp = np.concatenate([np.array(o[1]) for o in owned])
        res["partition"] = (np.array_equal(np.sort(alln), np.arange(d.n_nodes)) and
                            np.array_equal(np.sort(allp), np.arange(d.n_p)))
        # 2. ghosts are stored neighbour by neighbour, counts match the plan
        res["ghost_counts"] = (int(rnc.sum()) == loc["node_gid"].size - no and int(rpc.sum()) == loc["p_gid"].size - po)
        res["send_owned_only"] = bool((sni < no).all() and (spi < po).all() and (sni >= 0).all())
        # 3. reference system matrix from the oracle (identical on both ranks), random global vector
        o = case.oracle()
        rows, vals = case.bc(3.0)
        o.set_dirichlet(rows, vals)
        o.set_solution(0.2 * case.random_state())
        o.assemble_first()
        A = o.matrix("sys").tocsr()
        x = case.random_state()
        y_ref = A @ x
        # local vector in the caller layout [u(owned, ghost) | p(owned, ghost)], ghosts poisoned
        nn, npl = loc["node_gid"].size, loc["p_gid"].size
        xu = x[: d.n_u].reshape(-1, dim)[loc["node_gid"]].copy()
        xp = x[d.n_u:][loc["p_gid"]].copy()
        xu[no:] = np.nan; xp[po:] = np.nan
        # halo exchange emulated with gloo: pack owned values per neighbour, all_to_all, unpack in arrival order
        import torch

        def halo(vals_owned, ptr, idx, width):
            send = [np.zeros((0, width))] * world
            for k, q in enumerate(nbs):
                send[q] = vals_owned[idx[ptr[k]:ptr[k + 1]]].reshape(-1, width)
            recv = [None] * world
            dist.all_to_all_single  # noqa: B018  (object path below keeps the test backend agnostic)
            objs = [None] * world
            dist.all_gather_object(objs, [s.tolist() for s in send])
            got = [np.array(objs[q][rank]).reshape(-1, width) for q in nbs]
            return np.concatenate(got) if got else np.zeros((0, width))

        xu[no:] = halo(xu[:no], snp, sni, dim)
        xp[po:] = halo(xp[:po], spp, spi, 1).ravel()
        res["halo_values"] = bool(np.array_equal(xu, x[: d.n_u].reshape(-1, dim)[loc["node_gid"]]) and
                                  np.array_equal(xp, x[d.n_u:][loc["p_gid"]]))
        # owned rows of A restricted to local columns must reproduce the serial product
        gu = (dim * loc["node_gid"][:, None] + np.arange(dim)[None, :]).ravel()
        gp = d.n_u + loc["p_gid"]
        gcols = np.concatenate([gu, gp])
        grows = np.concatenate([gu[: dim * no], gp[:po]])
        Aloc = A[grows][:, gcols]
        res["columns_local"] = bool(abs(A[grows]).sum() == abs(Aloc).sum())  # no owned row reaches outside the halo
        y_loc = Aloc @ np.concatenate([xu.ravel(), xp])
        res["spmv"] = bool(np.allclose(y_loc, y_ref[grows], rtol=1e-13, atol=1e-13))
        # 4. two-layer halo: the Schur pattern of owned pressure rows and the B^T rows it needs are local
        nu = d.n_u
        B, Bt = A[nu:, :nu], A[:nu, nu:]
        Bp = sp.csr_matrix((np.ones(B.nnz), B.indices, B.indptr), shape=B.shape)
        Btp = sp.csr_matrix((np.ones(Bt.nnz), Bt.indices, Bt.indptr), shape=Bt.shape)
        S = (Bp @ Btp).tocsr()
        own_p = loc["p_gid"][:po]
        scols = np.unique(S[own_p].indices)
        res["schur_columns_local"] = bool((loc["g2l_p"][scols] >= 0).all())
        # nodes adjacent to owned pressures: every cell containing them must be a local cell
        adj_nodes = np.unique(Bp[own_p].indices // dim)
        nodes = d.cell_dofs()[:, D._node_cols(dim)] // dim
        cells_needed = np.nonzero(np.isin(nodes, adj_nodes).any(axis=1))[0]
        res["bt_rows_complete"] = bool(np.isin(cells_needed, loc["cells"]).all())
        res["local_cells_fraction"] = loc["cells"].size / d.n_cells
        out_q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case_name,world", [("cyl2d", 2), ("cyl3d", 2), ("cyl3d", 4)])
def test_partition_and_halo_over_gloo(case_name, world):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in results:
        for k, v in res.items():
            if k != "local_cells_fraction":
                assert v, (rank, k, res)
        assert res["local_cells_fraction"] < 1.0
