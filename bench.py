#!/usr/bin/env python
"""bench.py -- DoF-timesteps/sec of the per-timestep hot path (assemble_time_step + preconditioner
initialisation + outer GMRES) on the synthetic refined 3D cylinder, and the roofline of its
dominant kernel.

  python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
  python bench.py --impl reference --steps K --warmup W     (CPU restatement of the reference on
                                                            the host cores; the reference itself
                                                            needs deal.II+Trilinos+MPI and cannot
                                                            be built here)

One "step" = one time step of NavierStokes::solve (NavierStokes3D.cpp:714-735): assemble_time_step
+ solve_time_step, entered through nsb_step_host with HOST buffers (Dirichlet values host->device,
solution device->host).  The SAME K steps give both numbers: `value` from the device time of each
step (CUDA events around assembly + solve on the engine's stream, copies excluded, summed; max over
ranks) and `e2e` from the wall clock around the whole loop (copies and host work included; max over
ranks).

State preparation (untimed): the reference starts impulsively from u = 0, which makes time steps 1
and 2 several times harder than every later one (the pressure jumps by O(1/dt) twice).  They are
solved here to the reference's outer tolerance but with the inner solves tightened to 1e-4
(`PREP_INNER_RTOL`), which reaches the same converged state in a fraction of the outer iterations.
Everything after them -- the W warm-up steps and the K timed steps -- runs the reference's literals.

The run watches its own wall clock (NSB_BENCH_BUDGET_S, default 780 s -- the driver's per-run limit
is 870 s): if the K timed steps would not fit, fewer steps are timed and the line says so
(`detail.steps_requested`, `detail.truncated`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

_T0 = time.perf_counter()
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep NCCL's version banner off it
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
    os.environ["NCCL_DEBUG"] = "NONE"
# torchrun exports OMP_NUM_THREADS=1 when it is unset, which would serialise the host side of setup()
# (sparsity, scatter maps, ILU schedules: OpenMP in libnsb) on every rank and cripple the CPU arm.
# Give every rank its share of the host cores instead (NSB_KEEP_OMP=1 keeps the caller's value).
# Must happen before anything loads an OpenMP runtime.
if os.environ.get("OMP_NUM_THREADS") == "1" and "WORLD_SIZE" in os.environ and not os.environ.get("NSB_KEEP_OMP"):
    _world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ["WORLD_SIZE"])))
    _share = os.cpu_count() if "--impl" in sys.argv and "reference" in sys.argv else max(1, (os.cpu_count() or 1) // _world)
    os.environ["OMP_NUM_THREADS"] = str(_share)

import numpy as np  # noqa: E402

METRIC = "DoF-timesteps/sec (assemble_time_step + preconditioner init + outer GMRES)"
UNIT = "DoF-timesteps/s"
# workloads: name -> (variant, generator arguments).  "cyl3d-20M" is BASELINE.json configs[4]
# (refined 3D cylinder, ~20 M DoF, NavierStokes3D + Yosida, main3D.cpp:37-38); "cyl2d-2M" is configs[3]
# (refined 2D cylinder, ~2 M DoF, NavierStokes2D + aSIMPLE, main2D.cpp:21-22).
WORKLOADS = {"cyl3d-20M": ("3d", (8, 40)), "cyl3d-16M": ("3d", (8, 32)), "cyl3d-2M": ("3d", (4, 16)),
             "cyl3d-900k": ("3d", (3, 12)), "cyl3d-500k": ("3d", (3, 7)), "cyl3d-270k": ("3d", (2, 8)),
             "cyl3d-30k": ("3d", (1, 3)), "cyl2d-2M": ("2d", (28,)), "cyl2d-640k": ("2d", (16,)), "cyl2d-160k": ("2d", (8,)),
             "cyl2d-3k": ("2d", (1,))}
DELTAT = {"3d": 2e-4, "2d": 0.01}          # main3D.cpp:38, main2D.cpp:22
PRECOND = {"3d": "yosida", "2d": "asimple"}  # NavierStokes3D.cpp:562, NavierStokes2D.cpp:547
N_DOFS = {"cyl3d-20M": 19923035, "cyl3d-2M": 2059237, "cyl3d-500k": 530456}
AUTO_BLOCK_MIN_DOFS = 1.5e7                  # --ilu-ordering -1: block multicolour ILU above this many DoF per GPU
PREP_INNER_RTOL = 1e-4                      # inner tolerance of the two untimed start-up steps (3D only, see prep_rtol)
PREP_STEPS = 2


def prep_rtol(variant):
    """Inner tolerance of the untimed start-up steps.  3D: tightened (the impulsive start from u = 0 otherwise costs
    ~1 500 outer iterations).  2D: the reference's 1e-2 -- the inlet ramps up smoothly (test case 2) and the inner
    GMRES(28) + ILU(0) on the Schur complement already needs thousands of iterations per solve at 2 M DoF
    (tests/test_gpu_properties.py); a tighter tolerance would hit the reference's 10 000-iteration limit."""
    return PREP_INNER_RTOL if variant == "3d" else 1e-2


def make_mesh(workload):
    from navierstokes_project_nm4pde_b200 import HostMesh

    variant, a = WORKLOADS[workload]
    return (HostMesh.cylinder3d(*a) if variant == "3d" else HostMesh.cylinder2d(*a)), variant


def HostDofsCount(mesh):
    from navierstokes_project_nm4pde_b200 import HostDofs

    return HostDofs(mesh).N


def mesh_label(workload):
    variant, a = WORKLOADS[workload]
    return f"cylinder3d(s={a[0]}, nz={a[1]})" if variant == "3d" else f"cylinder2d(s={a[0]})"


def cpu_sample_for(workload, cores):
    """Bounded sample of the same mesh family for the CPU legs.  DoF-timesteps/s falls with the mesh
    size (outer iterations per step grow: ~25 at 0.27 M DoF, ~50 at 0.5 M, ~130 at 2 M, ~300 at 20 M),
    so the sample is the largest mesh whose time step still costs seconds on the host cores."""
    if WORKLOADS[workload][0] == "2d":
        return "cyl2d-160k"
    return "cyl3d-500k" if cores >= 12 else "cyl3d-270k"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def log(msg):
    """Progress on stderr (rank 0): where the wall time of a run goes; stdout stays one JSON line."""
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench {time.perf_counter() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------- CPU arm
def dof_partition(mesh, dofs, nparts):
    """DoF -> subdomain like the reference under mpirun -n P: cells by coordinate bisection, a DoF
    belongs to the lowest-numbered subdomain among its cells."""
    cell_part = mesh.partition(nparts)
    cd = dofs.cell_dofs(copy=False)
    part = np.full(dofs.N, nparts, np.int32)
    np.minimum.at(part, cd.ravel(), np.repeat(cell_part, cd.shape[1]))
    return part


def run_cpu(workload, steps, warmup, budget_s=None):
    """The oracle (CPU restatement of the reference algorithm, OpenMP over all host cores, block-Jacobi
    ILU with one subdomain per thread like mpirun -n <cores>) on a bounded sample.  Same procedure as
    the GPU arm: PREP_STEPS start-up steps with tightened inner solves, `warmup` untimed steps, `steps`
    timed steps (fewer when `budget_s` runs out)."""
    from navierstokes_project_nm4pde_b200 import NavierStokes
    from oracle import ns_ref as R

    t_begin = time.perf_counter()
    mesh, variant = make_mesh(workload)
    dt = DELTAT[variant]
    dim = 3 if variant == "3d" else 2
    prob = NavierStokes(mesh, variant, T=1.0, deltat=dt, test_case=2)
    prob.setup_host()
    d = prob.dofs
    num = dict(dim=dim, cell_dofs=d.cell_dofs(), N=d.N, n_u=d.n_u, n_p=d.n_p, dpc=d.dpc)
    o = R.Oracle(dim, variant, mesh.vertices, mesh.cells, num, R.system_pattern(num), 1e-3, dt)
    cores = int(R.lib().nso_num_threads())
    o.set_partition(dof_partition(mesh, d, cores))
    tm = dt
    o.set_dirichlet(prob._dir_rows, prob.dirichlet_values(tm))
    o.set_solution(np.zeros(d.N))
    ptype = PRECOND[variant]
    o.set_options(inner_rtol=prep_rtol(variant))
    o.assemble_first()
    o.solve_step(ptype)
    for _ in range(PREP_STEPS - 1):
        tm += dt
        o.set_dirichlet_values(prob.dirichlet_values(tm))
        o.assemble_step(); o.solve_step(ptype)
    o.set_options(inner_rtol=1e-2)  # Preconditioners.hpp:260
    for _ in range(warmup):
        tm += dt
        o.set_dirichlet_values(prob.dirichlet_values(tm))
        o.assemble_step(); o.solve_step(ptype)
    its, done = [], 0
    t0 = time.perf_counter()
    for _ in range(steps):
        tm += dt
        o.set_dirichlet_values(prob.dirichlet_values(tm))
        o.assemble_step()
        rc, k, _ = o.solve_step(ptype)
        its.append(k); done += 1
        el = time.perf_counter() - t0
        if budget_s is not None and done < steps and (time.perf_counter() - t_begin) + el / done > budget_s:
            break
    el = time.perf_counter() - t0
    return dict(value=d.N * done / el, n_dofs=d.N, cores=cores, seconds=el, iterations=its, steps=done, workload=workload,
                sample=f"{workload}: {mesh_label(workload)}, {d.N} DoF, {done} time step(s) after "
                       f"{PREP_STEPS} start-up + {warmup} warm-up steps, block-Jacobi ILU(0) over {cores} subdomains")


# ------------------------------------------------------------------------------------------- GPU arm
def pick_orderings(variant, n_dofs, world, ilu_ordering=-1, ilu_ordering_schur=-1):
    """(ILU ordering of F_s, of the Schur complement) for the throughput mode; -1 = automatic.
    F_s: by the DoFs one GPU holds -- block multicolour (2) above AUTO_BLOCK_MIN_DOFS, point multicolour (1) below
    (measured: the block sweeps need large colours, profiles/README.md).
    Schur complement: in 3D the pressure matrix keeps the point multicolour sweeps when F_s takes the block sweeps
    (session M: 0.41 ms against 0.83 ms per apply at 19.9 M DoF, same CG iteration count).  In 2D (aSIMPLE: restarted
    GMRES on the Schur complement, hundreds to thousands of iterations per solve) the block ordering, whose ILU(0) is as
    good as the natural one -- with the point multicolour factors the inner GMRES hits the reference's 10 000-iteration
    limit from 0.64 M DoF on (sessions M-P)."""
    if ilu_ordering < 0:
        ilu_ordering = 2 if n_dofs / max(world, 1) >= AUTO_BLOCK_MIN_DOFS else 1
    if ilu_ordering_schur < 0:
        ilu_ordering_schur = 2 if variant == "2d" else (1 if ilu_ordering == 2 else ilu_ordering)
    return ilu_ordering, ilu_ordering_schur


class GpuRun:
    """One problem instance on this rank's GPU, stepped through the host-buffer entry point."""

    def __init__(self, workload, args, world, rank, local_rank, uid=None):
        import torch

        from navierstokes_project_nm4pde_b200 import NavierStokes

        self.torch = torch
        self.mesh, self.variant = make_mesh(workload)
        n_dofs = N_DOFS.get(workload) or (HostDofsCount(self.mesh) if args.ilu_ordering < 0 else 0)
        self.ilu_ordering, self.ilu_ordering_schur = pick_orderings(self.variant, n_dofs, world, args.ilu_ordering,
                                                                    args.ilu_ordering_schur)
        self.dt = DELTAT[self.variant]
        kw = dict(T=1.0, deltat=self.dt, test_case=2, device=local_rank, ilu_ordering=self.ilu_ordering,
                  ilu_ordering_schur=self.ilu_ordering_schur, orthogonalisation=args.orthogonalisation)
        if world > 1:
            from navierstokes_project_nm4pde_b200.distributed import DistributedNavierStokes

            self.prob = DistributedNavierStokes(self.mesh, self.variant, nranks=world, rank=rank, unique_id=uid, **kw)
        else:
            self.prob = NavierStokes(self.mesh, self.variant, **kw)
        self.prob.setup()
        self.e = self.prob.engine
        self.n_dofs = self.prob.N_global if world > 1 else self.prob.N
        nd = max(len(self.prob._dir_rows), 1)
        self.dir_host = torch.empty(nd, dtype=torch.float64).pin_memory()
        self.sol_host = torch.empty(self.prob.N, dtype=torch.float64).pin_memory()
        self.dv, self.so = self.dir_host.numpy()[: len(self.prob._dir_rows)], self.sol_host.numpy()
        self.tm = 0.0
        self.steps_done = 0

    def step(self):
        """One time step through nsb_step_host: Dirichlet values of the new time level up, solution down."""
        self.tm += self.dt
        self.dv[:] = self.prob.dirichlet_values(self.tm)
        its = self.e.step_host(self.steps_done == 0, self.tm, self.dv, self.so)
        self.steps_done += 1
        return its

    def prepare(self):
        """u_0 = 0, then the PREP_STEPS start-up steps (tightened inner solves, reference outer tolerance)."""
        e = self.e
        e.set_solution(self.prob.initial_condition())
        e.set_params(inner_rtol=prep_rtol(self.variant))
        its = [self.step() for _ in range(PREP_STEPS)]
        e.set_params(inner_rtol=1e-2)  # Preconditioners.hpp:260
        return its


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    budget = float(os.environ.get("NSB_BENCH_BUDGET_S", "780"))
    # NSB_BENCH_BACKEND=gloo exists for the CPU test of this function's multi-rank control flow (tests/test_bench_contract.py,
    # fake engines); a real run is always NCCL
    backend = os.environ.get("NSB_BENCH_BACKEND", "nccl")
    if world > 1:
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    torch.cuda.set_device(local_rank)
    log(f"process group up: world {world}, OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op="max"):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda" if backend == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.MIN)
        return float(t.item())

    t_setup = time.perf_counter()
    uid = [None]
    if world > 1:
        if rank == 0:
            from navierstokes_project_nm4pde_b200 import Engine

            uid[0] = Engine.unique_id()
        dist.broadcast_object_list(uid, src=0)
    run = GpuRun(args.workload, args, world, rank, local_rank, uid[0])
    e, prob = run.e, run.prob
    setup_s = time.perf_counter() - t_setup
    log(f"setup done: {args.workload}, {run.mesh.n_cells} cells, {run.n_dofs} DoF, ilu_ordering {run.ilu_ordering} / {run.ilu_ordering_schur} (F_s / Schur), transport {getattr(prob, 'transport', 'none')}")

    barrier(); t0 = time.perf_counter()
    its_prep = run.prepare()
    barrier(); prep_s = time.perf_counter() - t0
    log(f"start-up steps (inner rtol {prep_rtol(run.variant):g}): {its_prep} outer iterations, {prep_s:.1f} s")

    warm_its, warm_s = [], []
    for _ in range(args.warmup):
        t0 = time.perf_counter()
        warm_its.append(run.step())
        warm_s.append(time.perf_counter() - t0)
        log(f"warm-up step: {warm_its[-1]} outer iterations, {warm_s[-1]:.1f} s")

    # ---- how many of the K steps fit the wall-clock budget (all ranks take rank 0's decision)
    extras_s = 25.0 + (55.0 if (world == 1 and not args.no_cpu_baseline) else 0.0) + (20.0 if world == 1 else 0.0)
    est = reduce(max(warm_s) if warm_s else prep_s / PREP_STEPS)
    left = budget - (time.perf_counter() - _T0) - extras_s
    steps = args.steps
    if est * steps > left:
        steps = max(1, min(args.steps, int(left / est)))
        log(f"budget {budget:.0f} s: {left:.0f} s left at ~{est:.1f} s/step -> timing {steps} of {args.steps} steps")
    steps = int(reduce(float(steps), "min"))

    # ---- timed region: `steps` time steps through nsb_step_host; device time -> value, wall clock -> e2e
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    its, t_prec, t_solve = [], [], []
    e.launch_count(reset=True)
    e.stat("t_step_dev_ms_reset")
    barrier()
    t0 = time.perf_counter()
    planned, slowest = steps, 0.0
    failure = None
    for i in range(planned):
        ts = time.perf_counter()
        try:
            its.append(run.step())
        except Exception as ex:  # noqa: BLE001  (e.g. SolverControl::NoConvergence of an inner solve: the same scalars decide
            # on every rank, so all ranks arrive here together); the steps finished so far are still a measurement
            failure = f"step {i + 1} of the timed region raised {type(ex).__name__}: {ex}"
            log(failure)
            if not its:
                raise
            steps = len(its)
            break
        t_prec.append(e.stat("t_prec_ms") * 1e-3); t_solve.append(e.stat("t_solve_ms") * 1e-3)
        # a step can cost twice the warm-up estimate (230-700 outer iterations): stop early rather than overrun the
        # budget (every rank takes the same decision; step_host has already synchronised, so this adds no device wait)
        t_last_ok = time.perf_counter()
        slowest = max(slowest, t_last_ok - ts)
        over = float((time.perf_counter() - _T0) + extras_s + slowest > budget)
        if i + 1 < planned and reduce(over) > 0.0:
            steps = i + 1
            log(f"budget {budget:.0f} s: stopping after {steps} of {planned} planned steps (slowest step {slowest:.1f} s)")
            break
    torch.cuda.synchronize()
    wall_ms = ((t_last_ok if failure else time.perf_counter()) - t0) * 1e3  # a failed step is not part of the measurement
    dev_ms = e.stat("t_step_dev_ms_reset")
    barrier()
    launches = e.launch_count(reset=True)
    clocks = sampler.stop() if rank == 0 else None
    dev_ms, wall_ms = reduce(dev_ms), reduce(wall_ms)
    log(f"timed steps: {its} outer iterations, {dev_ms / steps:.0f} ms/step on the device, {wall_ms / steps:.0f} ms/step wall")
    stats = {k: e.stat(k) for k in ("cnt_spmv_F", "cnt_spmv_S", "cnt_spmv_B", "cnt_spmv_Bt", "cnt_ilu_F", "cnt_ilu_S",
                                    "cnt_dot", "cnt_sync", "n_inner_F", "n_inner_S", "n_F_solves", "n_S_solves",
                                    "n_vmult")}
    value = run.n_dofs * steps / (dev_ms * 1e-3)
    e2e = dict(value=run.n_dofs * steps / (wall_ms * 1e-3), unit=UNIT,
               h2d_bytes_per_step=int(len(prob._dir_rows) * 8), d2h_bytes_per_step=int(prob.N * 8))

    # ---- roofline of the dominant kernel (isolated launches, L2 flushed between, CUDA events)
    peak, peak_src = measured_peaks()
    per_step = {"ilu_F": stats["cnt_ilu_F"], "spmv_F": stats["cnt_spmv_F"], "ilu_S": stats["cnt_ilu_S"],
                "spmv_S": stats["cnt_spmv_S"], "assemble_step": 1.0, "spmv_system": stats["n_vmult"]}
    kern = {}
    for name, cnt in per_step.items():
        try:
            kms, kbytes = e.bench_kernel(name, iters=5, flush_l2=True)
        except Exception as ex:  # noqa: BLE001  (a kernel that cannot be timed alone must not cost the whole line)
            log(f"bench_kernel({name}) failed: {ex}")
            continue
        kern[name] = dict(ms=kms, bytes=kbytes, gbs=kbytes / (kms * 1e-3) / 1e9, calls_last_step=cnt,
                          share_last_step=cnt * kms / (1e3 * (t_prec[-1] + t_solve[-1]) + 1e-9))
    if not kern:
        raise RuntimeError("no kernel of the step could be timed in isolation")
    dom = max(kern, key=lambda k: kern[k]["share_last_step"])
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    # (profiles/r02_traffic.json; only valid for the workload / rank count / ordering it was captured on)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            tj = json.load(f).get(dom, {})
        if tj.get("workload") == args.workload and world == 1 and tj.get("ilu_ordering", run.ilu_ordering) == run.ilu_ordering:
            traffic = tj.get("dram_bytes_per_apply", tj.get("dram_bytes_per_launch"))
    except Exception:
        pass
    roofline = dict(bound="hbm", kernel=dom, achieved=kern[dom]["gbs"], peak=peak, unit="GB/s",
                    frac=kern[dom]["gbs"] / peak, traffic=traffic, peak_source=peak_src,
                    kernels={k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                             for k, v in kern.items()})
    variant = run.variant
    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=steps, warmup=args.warmup,
               ms_per_step=dev_ms / steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
               data="synthetic", gpu_launches=int(launches),
               config=dict(workload=args.workload, mesh=mesh_label(args.workload), n_dofs=int(run.n_dofs),
                           n_cells=int(run.mesh.n_cells), variant={"3d": "NavierStokes3D", "2d": "NavierStokes2D"}[variant],
                           preconditioner=PRECOND[variant], deltat=run.dt,
                           quadrature="QGaussSimplex(3) / Witherden-Vincent",
                           ilu_ordering={0: "natural (reference replay)", 1: "multicolour (throughput mode)",
                                         2: "block multicolour, natural order inside 32-row blocks (throughput mode)",
                                         3: "subdomain ordering: parts solved out of shared memory, separators last "
                                            "(throughput mode)"}[run.ilu_ordering],
                           ilu_ordering_schur=int(run.ilu_ordering_schur),
                           orthogonalisation={0: "modified Gram-Schmidt (reference replay)",
                                              1: "batched classical Gram-Schmidt (throughput mode)"}[args.orthogonalisation],
                           l2_policy="working set (>1 GB of matrices) exceeds the 126 MB L2; isolated kernel "
                                     "timings flush L2 between launches",
                           start_up=f"{PREP_STEPS} untimed steps from u=0 with inner rtol {prep_rtol(variant):g}, then reference literals",
                           partition=f"{world} subdomain(s), coordinate bisection",
                           transport=(getattr(prob, "transport", "none") if world > 1 else "none")),
               e2e=e2e, roofline=roofline, clocks=clocks,
               detail=dict(outer_iterations=its, steps_requested=args.steps, truncated=bool(steps < args.steps), failure=failure,
                           start_up_iterations=its_prep, start_up_s=prep_s, warmup_iterations=warm_its,
                           warmup_s=[round(x, 2) for x in warm_s], setup_s=setup_s, t_prec_s=t_prec, t_solve_s=t_solve,
                           last_step_counts=stats, wall_ms_per_step=wall_ms / steps,
                           levels={k: e.stat(k) for k in ("levels_F_fwd", "sweeps_F", "sweeps_S", "ilu_blocks_F",
                                                           "ilu_blocks_S")}))
    left = budget - (time.perf_counter() - _T0)
    if rank == 0 and world == 1:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or cpu_sample_for(args.workload, cores)
        # the same mesh the CPU legs run, on the GPU: a like-for-like ratio next to the headline
        if left > 60 and sample != args.workload:
            try:
                del run, e, prob
                t0 = time.perf_counter()
                small = GpuRun(sample, args, 1, 0, local_rank)
                small.prepare()
                for _ in range(2):
                    small.step()
                small.e.stat("t_step_dev_ms_reset")
                ks = 5
                t1 = time.perf_counter()
                sits = [small.step() for _ in range(ks)]
                torch.cuda.synchronize()
                w = time.perf_counter() - t1
                d_ms = small.e.stat("t_step_dev_ms_reset")
                out["detail"]["same_mesh"] = dict(workload=sample, n_dofs=int(small.n_dofs), steps=ks, outer_iterations=sits,
                                                  value=small.n_dofs * ks / (d_ms * 1e-3), e2e=small.n_dofs * ks / w,
                                                  seconds=w, total_s=time.perf_counter() - t0)
                log(f"same mesh as the CPU legs ({sample}): {sits} outer iterations, {1e3 * w / ks:.0f} ms/step")
                del small
            except Exception as ex:  # noqa: BLE001
                out["detail"]["same_mesh"] = dict(error=str(ex))
        left = budget - (time.perf_counter() - _T0)
        if not args.no_cpu_baseline and left > 45:
            try:
                cpu = run_cpu(sample, 1, 0)
                out["cpu_baseline"] = dict(value=cpu["value"], unit=UNIT, cores=cpu["cores"], kind="port",
                                           sample=cpu["sample"], seconds=cpu["seconds"], outer_iterations=cpu["iterations"])
            except Exception as ex:  # noqa: BLE001
                out["cpu_baseline"] = dict(value=None, unit=UNIT, cores=cores, kind="port", sample=f"failed: {ex}")
        elif not args.no_cpu_baseline:
            out["cpu_baseline"] = dict(value=None, unit=UNIT, cores=cores, kind="port",
                                       sample="skipped: wall-clock budget of the run exhausted (see --impl reference)")
    if rank == 0:
        print(json.dumps(out), flush=True)
    log("done")
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cyl3d-20M", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", default=None, choices=sorted(WORKLOADS),
                    help="mesh of the bounded CPU sample (cpu_baseline leg and --impl reference); default by core count")
    ap.add_argument("--ilu-ordering", type=int, default=-1, choices=[-1, 0, 1, 2, 3],
                    help="0: natural row order (reference replay), 1: multicolour ILU(0), 2: block multicolour ILU(0), "
                         "3: subdomain-resident ILU(0) (throughput modes); -1 (default): 2 above 15 M DoF per GPU, else 1 "
                         "(measured: the block sweeps need large colours, profiles/README.md)")
    ap.add_argument("--ilu-ordering-schur", type=int, default=-1, choices=[-1, 0, 1, 2, 3],
                    help="ordering of the Schur-complement factors; -1 (default): 2 in 2D, 1 in 3D when F_s uses 2, else the same as F_s")
    ap.add_argument("--orthogonalisation", type=int, default=1, choices=[0, 1],
                    help="0: modified Gram-Schmidt as deal.II (reference replay), 1: batched classical Gram-Schmidt")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or cpu_sample_for(args.workload, cores)
        budget = float(os.environ.get("NSB_BENCH_BUDGET_S", "780"))
        cpu = run_cpu(sample, args.steps, args.warmup, budget_s=budget - 20)
        variant = WORKLOADS[sample][0]
        print(json.dumps(dict(
            impl="reference", metric=METRIC, value=cpu["value"], unit=UNIT, n_gpus=args.gpus, steps=cpu["steps"],
            warmup=args.warmup, ms_per_step=cpu["seconds"] * 1e3 / cpu["steps"], higher_is_better=True, scaling="strong",
            vs_baseline=None, dtype="f64", data="synthetic", gpu_launches=0,
            config=dict(workload=sample, mesh=mesh_label(sample), n_dofs=int(cpu["n_dofs"]), sample_of=args.workload,
                        variant={"3d": "NavierStokes3D", "2d": "NavierStokes2D"}[variant],
                        preconditioner=PRECOND[variant], deltat=DELTAT[variant],
                        note="bounded sample of the headline workload (same mesh family, fewer DoF: the CPU cannot run "
                             f"{args.workload} in benchmark time); the GPU arm reports the same mesh under detail.same_mesh; "
                             "the timed loop runs the oracle port only (libnsb.so is mapped for mesh / DoF generation)"),
            cpu_baseline=dict(value=cpu["value"], unit=UNIT, cores=cpu["cores"], kind="port", sample=cpu["sample"]),
            e2e=dict(value=cpu["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
            detail=dict(outer_iterations=cpu["iterations"], steps_requested=args.steps,
                        truncated=bool(cpu["steps"] < args.steps)))), flush=True)
        return
    run_gpu(args)


if __name__ == "__main__":
    main()
