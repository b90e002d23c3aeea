"""N > 1 path on real GPUs (`-m gpu`, needs >= 2 devices): two ranks -- ghost exchange and all-reduced
dot products over peer memory (NVLink P2P stores, csrc/halo.cu) or over NCCL -- must reproduce the
oracle run with the same block-Jacobi partition (the reference's `mpirun -n 2` semantics)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _solver_overrides(case_name):
    """The 1 093-DoF Ethier-Steinman cube split into subdomains is a STAGNATING solve at the reference's
    inner tolerance 1e-2 (110-300 outer iterations of GMRES(28) for a system of 1 093 unknowns): its
    iteration count is round-off noise -- the oracle alone gives 26 or 24 iterations for step 3 depending
    on its OpenMP thread count, and 248 or 305 for step 1 between two runs on the GPU box
    (profiles/r02_multi_gpu.md) -- which is what made the round-1 `cube-1-p2p` case fail, not the
    transport.  With the inner solves tightened to 1e-6 the preconditioner is (nearly) a fixed operator,
    the solve takes 13 / 13 / 4 iterations on any summation order, and the case tests what it is meant to:
    the CONV variant (Neumann rhs, exact-solution Dirichlet rows, non-zero initial state) across ranks."""
    return {"inner_rtol": 1e-6} if case_name == "cube" else {}


def _worker(rank, world, port, case_name, ordering, orth, transport, out_q):
    import sys

    os.environ["NSB_P2P"] = "1" if transport == "p2p" else "0"

    import torch
    import torch.distributed as dist

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here)); sys.path.insert(0, here)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import helpers as T
        from navierstokes_project_nm4pde_b200 import Engine
        from navierstokes_project_nm4pde_b200.distributed import DistributedNavierStokes

        uid = [Engine.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        case = T.Case(case_name)
        prob = DistributedNavierStokes(case.mesh, case.variant, T=1.0, deltat=case.dt, test_case=2 if case.dim == 3 else 3,
                                       device=rank, nranks=world, rank=rank, unique_id=uid[0], ilu_ordering=ordering,
                                       orthogonalisation=orth, **_solver_overrides(case_name))
        prob.setup()
        e = prob.engine
        assert prob.transport == transport and e.stat("p2p") == (1.0 if transport == "p2p" else 0.0), prob.transport
        e.set_solution(prob.initial_condition())
        its, sols, forces = [], [], []
        t = 0.0
        for step in range(3):
            if case.variant == "conv":
                e.set_neumann_rhs(prob.neumann_rhs(t))
            t += case.dt
            tb = t if case.variant == "conv" else 2.0 + t
            e.set_dirichlet_values(prob.dirichlet_values(tb))
            if step == 0:
                e.assemble_first(t)
            else:
                e.assemble_step(t)
            its.append(e.solve_step()[0])
            gn, u, gp, p = prob.owned_solution()
            sols.append((gn.copy(), u.copy(), gp.copy(), p.copy()))
            # drag / lift on the device, summed over the ranks (each rank integrates the faces of its own cells)
            forces.append(e.compute_forces() if case.variant != "conv" else np.zeros(2))
        loc = prob.local
        out_q.put((rank, its, sols, loc["node_owner"] if rank == 0 else None, loc["p_owner"] if rank == 0 else None,
                   (e.ilu_order(0), e.ilu_order(1), loc["node_gid"][: loc["n_nodes_owned"]], loc["p_gid"][: loc["n_p_owned"]]),
                   forces))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case_name,ordering,orth,transport,world",
                         [("cyl3d", 0, 0, "p2p", 2), ("cyl3d", 1, 1, "p2p", 2), ("cyl2d", 0, 0, "p2p", 2),
                          ("cube", 1, 0, "p2p", 2), ("cube", 1, 0, "nccl", 2), ("cyl3d", 2, 1, "p2p", 2),
                          ("cyl3d", 1, 1, "nccl", 2), ("cyl2d", 0, 0, "nccl", 2), ("cyl2d", 2, 1, "p2p", 2),
                          ("cyl3d", 2, 1, "p2p", 4), ("cyl3d", 1, 1, "nccl", 4), ("cube", 2, 1, "p2p", 4)])
def test_ranks_match_block_jacobi_oracle(case_name, ordering, orth, transport, world):
    """ordering / orth: replay (0, 0) or throughput mode (multicolour / block-multicolour ILU(0), batched
    Gram-Schmidt); `world` ranks, one per GPU, against the oracle with the same block-Jacobi partition."""
    from navierstokes_project_nm4pde_b200 import Engine

    if Engine.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    import helpers as T

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case_name, ordering, orth, transport, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue as _queue
    import time as _time

    res, deadline = [], _time.time() + 240
    while len(res) < len(procs) and _time.time() < deadline:
        try:
            res.append(q.get(timeout=2))
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break  # a rank died: do not wait for the others (they would spin on its flags)
    if len(res) < len(procs):
        for p in procs:
            p.kill()
        pytest.fail(f"ranks did not finish: exit codes {[p.exitcode for p in procs]}")
    res.sort(key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = T.Case(case_name)
    dim = case.dim
    node_owner, p_owner = res[0][3], res[0][4]
    part = np.concatenate([np.repeat(node_owner, dim), p_owner]).astype(np.int32)
    o = case.oracle()
    o.set_partition(part)
    o.set_orthogonalisation(orth)
    if _solver_overrides(case_name):
        o.set_options(**_solver_overrides(case_name))
    if ordering >= 1:
        # global ILU ordering = each rank's multicolour order of its owned block, rank after rank
        ou, op = [], []
        for r in res:
            on, opp, gn, gp = r[5]
            ou.append((dim * gn[on][:, None] + np.arange(dim)[None, :]).ravel())
            op.append(case.n_u + gp[opp])
        o.set_ilu_order(np.concatenate(ou), np.concatenate(op) - case.n_u)
    rows, vals = case.bc(0.0)
    o.set_dirichlet(rows, vals)
    o.set_solution(case.initial())
    ptype = T.VARIANT_PREC[case.variant]
    t = 0.0
    for step in range(3):
        if case.variant == "conv":
            o.set_neumann_rhs(case.neumann(t))
        t += case.dt
        rows, vals = case.bc(t if case.variant == "conv" else 2.0 + t)
        o.set_dirichlet_values(vals)
        if step == 0:
            o.assemble_first()
        else:
            o.assemble_step()
        rc, its_o, _ = o.solve_step(ptype)
        assert rc == 0
        xo = o.array("sol_owned", case.N)
        xe = np.zeros(case.N)
        for r in res:
            gn, u, gp, p = r[2][step]
            xe[: case.n_u].reshape(-1, dim)[gn] = u
            xe[case.n_u:][gp] = p
        assert all(r[1][step] == res[0][1][step] for r in res)
        if case.variant != "conv":  # compute_forces across ranks == the oracle's face loop on the assembled global field
            fc, fl = case.dofs.boundary_faces(3)
            f_ref = o.compute_forces(xe, fc, fl, *T.gauss_simplex(dim - 1))
            for r in res:
                assert np.allclose(r[6][step], f_ref, rtol=1e-10, atol=1e-12 * np.abs(f_ref).max()), (step, r[0], r[6][step], f_ref)
        if res[0][1][step] == its_o:
            assert T.rel_l2(xe[: case.n_u], xo[: case.n_u]) < 1e-8, step
        else:
            # long solves (hundreds of iterations) can cross the absolute 1e-4 stopping threshold one
            # iteration apart because the all-reduced dot products are summed in a different order;
            # the fields then agree to the solver tolerance, not to round-off
            assert abs(res[0][1][step] - its_o) <= 1 and its_o > 100, (step, res[0][1], its_o)
            assert T.rel_l2(xe[: case.n_u], xo[: case.n_u]) < 1e-4, step
            break


# ---------------------------------------------------------------------------------------------------------
# The C++ NavierStokes class and its drivers, one process per GPU (the reference under mpirun, main3D.cpp:9,28)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "navierstokes_project_nm4pde_b200", "csrc", "host", "bin")
LAUNCH = os.path.join(ROOT, "scripts", "nsb_launch.sh")


def _solve_worker(rank, world, port, out_q):
    import sys

    import torch
    import torch.distributed as dist

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here)); sys.path.insert(0, here)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from navierstokes_project_nm4pde_b200 import Engine, HostMesh
        from navierstokes_project_nm4pde_b200.distributed import DistributedNavierStokes

        uid = [Engine.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        prob = DistributedNavierStokes(HostMesh.cylinder3d(1, 3), "3d", T=4.0, deltat=0.0002, test_case=2, device=rank,
                                       nranks=world, rank=rank, unique_id=uid[0])
        prob.forces_after = 0.0
        prob.setup()
        prob.solve(max_steps=3)
        out_q.put((rank, prob.gmres_iterations, list(zip(prob.vec_drag_coeff, prob.vec_lift_coeff)), prob.transport))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_cpp_driver_two_ranks_matches_python_two_ranks(tmp_path, transport):
    """navier_stokes3D started once per GPU by scripts/nsb_launch.sh (rendezvous.hpp: NCCL id and IPC handles over
    TCP, subdomains from nsh_local_*) against the Python multi-rank class on the same mesh: same partition, same
    engine calls, hence the same outer iteration counts and force coefficients."""
    import re
    import subprocess

    from navierstokes_project_nm4pde_b200 import Engine

    if Engine.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, NSB_MAX_STEPS="3", NSB_FORCES_AFTER="0", NSB_P2P="1" if transport == "p2p" else "0",
               NSB_RDV_PORT=str(_free_port()))
    r = subprocess.run([LAUNCH, "2", os.path.join(BIN, "navier_stokes3D"), "gen:cylinder3d:1:3"], capture_output=True, text=True,
                       env=env, cwd=tmp_path, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    assert f"2 ranks, transport {transport}" in out
    its_cpp = [int(m) for m in re.findall(r"Result:\s+(\d+) GMRES iterations", out)]
    coeff_cpp = [(float(a), float(b)) for a, b in re.findall(r"Coeff:\s+(\S+) Coeff:\s+(\S+)", out)]
    assert len(its_cpp) == 3 and len(coeff_cpp) == 3
    assert "Result" not in (tmp_path / "rank1.log").read_text()  # only rank 0 prints (pcout)

    import torch.multiprocessing as mp

    os.environ["NSB_P2P"] = "1" if transport == "p2p" else "0"
    try:
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_solve_worker, args=(k, 2, port, q)) for k in range(2)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=240) for _ in procs)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        os.environ.pop("NSB_P2P", None)
    assert res[0][3] == transport
    assert its_cpp == res[0][1] == res[1][1]
    assert np.allclose(np.array(coeff_cpp), np.array(res[0][2]), rtol=1e-5, atol=1e-9)  # 6 printed digits


def test_cpp_convergence_driver_two_ranks(tmp_path):
    """Ethier-Steinman driver on 2 ranks (Neumann face term, exact Dirichlet data and initial state through the
    local numbering; compute_error summed over the ranks): the errors equal the single-rank run to solver
    tolerance and the orders hold."""
    import math
    import subprocess

    from navierstokes_project_nm4pde_b200 import Engine

    if Engine.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    errs = {}
    for n in (1, 2):
        d = tmp_path / f"n{n}"
        d.mkdir()
        r = subprocess.run([LAUNCH, str(n), os.path.join(BIN, "convergence"), "gen:cube:4", "gen:cube:8"], capture_output=True,
                           text=True, env=dict(os.environ, NSB_RDV_PORT=str(_free_port())), cwd=d, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        rows = (d / "convergence.csv").read_text().strip().splitlines()[1:]
        errs[n] = np.array([[float(v) for v in ln.split(",")] for ln in rows])
    assert errs[2].shape == (2, 3)
    assert np.allclose(errs[1], errs[2], rtol=2e-2)
    assert 2.5 < math.log2(errs[2][0, 1] / errs[2][1, 1]) < 3.6 and 1.6 < math.log2(errs[2][0, 2] / errs[2][1, 2]) < 2.6
