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
