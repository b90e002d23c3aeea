"""Exchange plan of the N > 1 path for 8 ranks, built for all ranks inside one process (no GPU, no
process group): the plan must be symmetric (every rank lists exactly the neighbours that list it),
send and receive counts must agree pairwise, and an emulated ghost exchange + all-reduce following
the mailbox protocol of csrc/halo.cu (per-kind sequence numbers, parity double buffering, flags per
source rank) must deliver exactly the owners' values whatever order the ranks run in."""
import numpy as np
import pytest

import helpers as T
from navierstokes_project_nm4pde_b200 import distributed as D

WORLD = 8


def _plans(case_name):
    case = T.Case(case_name)
    d, dim = case.dofs, case.dim
    part = case.mesh.partition(WORLD)
    locs = [D.build_local_problem(dim, d.cell_dofs(), d.cell_coords(), d.n_nodes, d.n_p, part, WORLD, r) for r in range(WORLD)]
    # in-process all_to_all: rank r's call returns what every q addressed to r
    sends = {}
    for kind in ("nodes", "p"):
        for r in range(WORLD):
            need = locs[r]["need_" + kind]
            sends[(kind, r)] = [need.get(q, np.zeros(0, np.int64)).astype(np.int64) for q in range(WORLD)]
    plans = []
    for r in range(WORLD):
        calls = iter(["nodes", "p"])

        def a2a(_send, r=r, calls=calls):
            kind = next(calls)
            return [sends[(kind, q)][r] for q in range(WORLD)]

        send_nodes, send_p = D.exchange_requests(locs[r], WORLD, r, a2a)
        plans.append(D.halo_arrays(locs[r], send_nodes, send_p))
    return case, locs, plans


@pytest.mark.parametrize("case_name", ["cyl3d", "cyl2d"])
def test_eight_rank_plan_is_symmetric_and_complete(case_name):
    case, locs, plans = _plans(case_name)
    dim = case.dim
    nbsets = [set(p[0].tolist()) for p in plans]
    for r in range(WORLD):
        assert r not in nbsets[r]
        for q in nbsets[r]:
            assert r in nbsets[q], (r, q)  # k_halo_wait on r spins on a flag that only q's push raises
    for r in range(WORLD):
        nbs, snp, sni, rnc, spp, spi, rpc = plans[r]
        no, po = locs[r]["n_nodes_owned"], locs[r]["n_p_owned"]
        assert int(rnc.sum()) == locs[r]["node_gid"].size - no and int(rpc.sum()) == locs[r]["p_gid"].size - po
        for k, q in enumerate(nbs.tolist()):
            kq = plans[q][0].tolist().index(r)
            assert snp[k + 1] - snp[k] == plans[q][3][kq]  # what r sends to q is what q expects from r
            assert spp[k + 1] - spp[k] == plans[q][6][kq]
    # ownership is a partition
    alln = np.concatenate([locs[r]["node_gid"][: locs[r]["n_nodes_owned"]] for r in range(WORLD)])
    assert np.array_equal(np.sort(alln), np.arange(case.dofs.n_nodes))


def test_mailbox_protocol_emulation():
    """Ranks execute their exchange sequence in a random interleaving, constrained only by what the
    protocol enforces (a wait needs the neighbours' flags of that sequence number): with two parity
    buffers no inbox is overwritten before it has been consumed, and every ghost gets its owner's value."""
    case, locs, plans = _plans("cyl3d")
    dim, rng = case.dim, np.random.default_rng(T.SEED)
    n_ex = 6
    # per exchange e, the owner value of global node g is value(e, g)
    value = lambda e, g: 1000.0 * (e + 1) + g  # noqa: E731
    inbox = [[np.full(max(1, int(plans[r][3].sum())), np.nan), np.full(max(1, int(plans[r][3].sum())), np.nan)] for r in range(WORLD)]
    flag = [[{q: 0 for q in plans[r][0].tolist()} for _ in range(2)] for r in range(WORLD)]
    recv_off = [dict(zip(plans[r][0].tolist(), np.concatenate([[0], np.cumsum(plans[r][3])])[:-1].tolist())) for r in range(WORLD)]
    pc = [0] * WORLD  # program counter: 2*e = push of exchange e, 2*e+1 = wait+consume
    done = 0
    while done < WORLD:
        r = int(rng.integers(WORLD))
        if pc[r] >= 2 * n_ex:
            continue
        e, phase = divmod(pc[r], 2)
        seq, par = e + 1, (e + 1) & 1
        nbs, snp, sni, rnc = plans[r][0].tolist(), plans[r][1], plans[r][2], plans[r][3]
        if phase == 0:  # k_halo_push: store into the neighbours' inboxes, then raise their flags
            gid = locs[r]["node_gid"]
            for k, q in enumerate(nbs):
                vals = value(e, gid[sni[snp[k]:snp[k + 1]]])
                o = recv_off[q][r]
                assert np.isnan(inbox[q][par][o:o + vals.size]).all(), "inbox overwritten before it was consumed"
                inbox[q][par][o:o + vals.size] = vals
                flag[q][par][r] = seq
            pc[r] += 1
        else:  # k_halo_wait: needs every neighbour's flag, then copies the inbox into the ghost segment
            if any(flag[r][par][q] < seq for q in nbs):
                continue
            no = locs[r]["n_nodes_owned"]
            ghosts = locs[r]["node_gid"][no:]
            got = inbox[r][par][: ghosts.size].copy()
            assert np.array_equal(got, value(e, ghosts)), (r, e)
            inbox[r][par][:] = np.nan  # consumed
            pc[r] += 1
            if pc[r] == 2 * n_ex:
                done += 1
    assert dim == 3


@pytest.mark.parametrize("case_name,world", [("cyl3d", 8), ("cyl2d", 4), ("cube", 2)])
def test_host_local_problem_matches_python_plan(case_name, world):
    """nsh_local_* (csrc/host_local.cpp, what the C++ NavierStokes class uses under a launcher) derives the
    local problem AND the exchange plan of every rank from the replicated mesh with no message; it must be
    identical to the plan distributed.py negotiates with an all-to-all."""
    import ctypes as C

    from navierstokes_project_nm4pde_b200._lib import lib

    L = lib()
    case = T.Case(case_name)
    d, dim = case.dofs, case.dim
    part = case.mesh.partition(world)
    locs = [D.build_local_problem(dim, d.cell_dofs(), d.cell_coords(), d.n_nodes, d.n_p, part, world, r) for r in range(world)]
    sends = {(kind, r): [locs[r]["need_" + kind].get(q, np.zeros(0, np.int64)).astype(np.int64) for q in range(world)]
             for kind in ("nodes", "p") for r in range(world)}
    for r in range(world):
        calls = iter(["nodes", "p"])
        def a2a(_send, r=r, calls=calls):
            kind = next(calls)
            return [sends[(kind, q)][r] for q in range(world)]

        want = D.halo_arrays(locs[r], *D.exchange_requests(locs[r], world, r, a2a))
        h = C.c_void_p(L.nsh_local_create(case.mesh.h, d.h, world, r))
        assert h
        try:
            arr = lambda p, n: np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.int32)  # noqa: E731
            nc, nn, npl = L.nsh_local_n_cells(h), L.nsh_local_n_nodes(h), L.nsh_local_n_p(h)
            loc = locs[r]
            assert np.array_equal(arr(L.nsh_local_cell_part(h), d.n_cells), part)
            assert np.array_equal(arr(L.nsh_local_cells(h), nc), loc["cells"])
            assert L.nsh_local_n_nodes_owned(h) == loc["n_nodes_owned"] and L.nsh_local_n_p_owned(h) == loc["n_p_owned"]
            assert np.array_equal(arr(L.nsh_local_node_gid(h), nn), loc["node_gid"])
            assert np.array_equal(arr(L.nsh_local_p_gid(h), npl), loc["p_gid"])
            assert np.array_equal(arr(L.nsh_local_cell_dofs(h), nc * d.dpc).reshape(nc, -1), loc["cell_dofs"])
            cc = np.ctypeslib.as_array(L.nsh_local_cell_coords(h), shape=(nc, dim + 1, dim))
            assert np.array_equal(cc, loc["cell_coords"])
            assert np.array_equal(arr(L.nsh_local_g2l_node(h), d.n_nodes), loc["g2l_node"])
            g2c = arr(L.nsh_local_g2l_cell(h), d.n_cells)
            assert np.array_equal(np.nonzero(g2c >= 0)[0], loc["cells"])
            ptrs = [C.POINTER(C.c_int32)() for _ in range(7)]
            nnb = L.nsh_local_halo(h, *[C.byref(p) for p in ptrs])
            assert nnb == want[0].size
            got_nb = arr(ptrs[0], nnb)
            snp, spp = arr(ptrs[1], nnb + 1), arr(ptrs[4], nnb + 1)
            got = (got_nb, snp, arr(ptrs[2], int(snp[-1])), arr(ptrs[3], nnb), spp, arr(ptrs[5], int(spp[-1])), arr(ptrs[6], nnb))
            for a, b in zip(got, want):
                assert np.array_equal(a, b)
        finally:
            L.nsh_local_free(h)
