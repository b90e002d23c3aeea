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
+ solve_time_step.  `value` times K steps with every input resident in HBM; `e2e` times K further
steps through nsb_step_host with host buffers (Dirichlet values H2D, solution D2H per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

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
# workloads: name -> (s, nz) of HostMesh.cylinder3d; "cyl3d-20M" is BASELINE.json configs[4]
WORKLOADS = {"cyl3d-20M": (8, 40), "cyl3d-16M": (8, 32), "cyl3d-2M": (4, 16), "cyl3d-900k": (3, 12), "cyl3d-500k": (3, 7), "cyl3d-270k": (2, 8),
             "cyl3d-30k": (1, 3)}
# Bounded sample of the same mesh family for the CPU legs.  DoF-timesteps/s falls with the mesh size
# (outer iterations per step: ~25 at 0.27 M DoF, ~85 at 0.5 M, ~130 at 2 M, ~450 at 20 M), so the
# sample is the largest mesh whose time step still costs ~30 s on 16 host cores.
CPU_SAMPLE = "cyl3d-500k"
DT = 2e-4                  # main3D.cpp:38


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


# ------------------------------------------------------------------------------------------- CPU arm
def dof_partition(mesh, dofs, nparts):
    """DoF -> subdomain like the reference under mpirun -n P: cells by coordinate bisection, a DoF
    belongs to the lowest-numbered subdomain among its cells."""
    cell_part = mesh.partition(nparts)
    cd = dofs.cell_dofs(copy=False)
    part = np.full(dofs.N, nparts, np.int32)
    np.minimum.at(part, cd.ravel(), np.repeat(cell_part, cd.shape[1]))
    return part


def run_cpu(workload, steps, warmup, threads=None):
    """The oracle (CPU restatement of the reference algorithm, OpenMP over all host cores, block-Jacobi
    ILU with one subdomain per thread like mpirun -n <cores>) on a bounded sample."""
    from navierstokes_project_nm4pde_b200 import HostMesh, NavierStokes
    from oracle import ns_ref as R

    s, nz = WORKLOADS[workload]
    mesh = HostMesh.cylinder3d(s, nz)
    prob = NavierStokes(mesh, "3d", T=1.0, deltat=DT, test_case=2)
    prob.setup_host()
    d = prob.dofs
    num = dict(dim=3, cell_dofs=d.cell_dofs(), N=d.N, n_u=d.n_u, n_p=d.n_p, dpc=d.dpc)
    o = R.Oracle(3, "3d", mesh.vertices, mesh.cells, num, R.system_pattern(num), 1e-3, DT)
    cores = int(R.lib().nso_num_threads())
    o.set_partition(dof_partition(mesh, d, cores))
    o.set_dirichlet(prob._dir_rows, prob.dirichlet_values(DT))
    o.set_solution(np.zeros(d.N))
    o.assemble_first()
    o.solve_step("yosida")
    t = DT
    for _ in range(warmup):
        t += DT
        o.assemble_step(); o.solve_step("yosida")
    its = []
    t0 = time.perf_counter()
    for _ in range(steps):
        t += DT
        o.assemble_step()
        rc, k, _ = o.solve_step("yosida")
        its.append(k)
    dt = time.perf_counter() - t0
    return dict(value=d.N * steps / dt, n_dofs=d.N, cores=cores, seconds=dt, iterations=its,
                sample=f"{workload}: cylinder3d(s={s}, nz={nz}), {d.N} DoF, {steps} time step(s) after "
                       f"{warmup + 1} untimed, block-Jacobi ILU(0) over {cores} subdomains")


# ------------------------------------------------------------------------------------------- GPU arm
_T0 = time.perf_counter()


def log(msg):
    """Progress on stderr (rank 0): where the wall time of a run goes; stdout stays one JSON line."""
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench {time.perf_counter() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from navierstokes_project_nm4pde_b200 import HostMesh, NavierStokes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    s, nz = WORKLOADS[args.workload]
    log(f"process group up: world {world}, OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')}")
    t_setup = time.perf_counter()
    mesh = HostMesh.cylinder3d(s, nz)
    log(f"mesh {args.workload}: {mesh.n_cells} cells")
    if world > 1:
        from navierstokes_project_nm4pde_b200.distributed import DistributedNavierStokes

        uid = [None]
        if rank == 0:
            from navierstokes_project_nm4pde_b200 import Engine

            uid[0] = Engine.unique_id()
        dist.broadcast_object_list(uid, src=0)
        prob = DistributedNavierStokes(mesh, "3d", T=1.0, deltat=DT, test_case=2, device=local_rank, nranks=world,
                                       rank=rank, unique_id=uid[0], ilu_ordering=args.ilu_ordering,
                                       orthogonalisation=args.orthogonalisation)
    else:
        prob = NavierStokes(mesh, "3d", T=1.0, deltat=DT, test_case=2, device=local_rank,
                            ilu_ordering=args.ilu_ordering, orthogonalisation=args.orthogonalisation)
    prob.setup()
    e = prob.engine
    n_dofs_global = prob.N_global if world > 1 else prob.N
    setup_s = time.perf_counter() - t_setup
    log(f"setup done: {n_dofs_global} DoF, transport {getattr(prob, 'transport', 'none')}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # step 1: NavierStokes::assemble (first-step path, amortised) -- untimed
    e.set_solution(prob.initial_condition())
    tm = DT
    e.set_dirichlet_values(prob.dirichlet_values(tm))
    # The impulsive start from u = 0 makes this the hardest solve of the run (>1000 outer iterations at
    # 20 M DoF); it is state preparation, not part of any timed region, so it is capped at
    # --first-step-cap outer iterations (0 = run to the reference tolerance).
    from navierstokes_project_nm4pde_b200._lib import NsbError

    barrier(); t0 = time.perf_counter()
    e.assemble_first(tm)
    first_converged = True
    if args.first_step_cap > 0:
        e.set_params(outer_maxit=args.first_step_cap)
    try:
        its_first = e.solve_step()[0]
    except NsbError as ex:
        if ex.code != -4:  # NSB_ERR_NOCONV
            raise
        its_first, first_converged = args.first_step_cap, False
    e.set_params(outer_maxit=100000)  # NavierStokes3D.cpp:551
    barrier(); first_step_s = time.perf_counter() - t0
    log(f"first step: {its_first} outer iterations, converged {first_converged}")
    for _ in range(args.warmup):
        tm += DT
        e.assemble_step(tm)
        log(f"warm-up step: {e.solve_step()[0]} outer iterations")

    # ---- timed region 1: K steps, inputs resident in HBM (CUDA events on the launching stream)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    its, t_prec, t_solve, launches = [], [], [], 0
    e.launch_count(reset=True)
    barrier()
    e.timer_start()
    for _ in range(args.steps):
        tm += DT
        e.assemble_step(tm)
        k, tp, ts = e.solve_step()
        its.append(k); t_prec.append(tp); t_solve.append(ts)
    ms = e.timer_stop_ms()
    barrier()
    log(f"timed steps: {its} outer iterations, {ms / args.steps:.0f} ms/step")
    launches = e.launch_count(reset=True)
    ms = reduce_max(ms)
    stats = {k: e.stat(k) for k in ("cnt_spmv_F", "cnt_spmv_S", "cnt_spmv_B", "cnt_spmv_Bt", "cnt_ilu_F", "cnt_ilu_S",
                                    "cnt_dot", "cnt_sync", "n_inner_F", "n_inner_S", "n_F_solves", "n_S_solves",
                                    "n_vmult")}
    value = n_dofs_global * args.steps / (ms * 1e-3)

    # ---- timed region 2: K steps end to end through nsb_step_host with (pinned) host buffers
    dir_host = torch.empty(max(len(prob._dir_rows), 1), dtype=torch.float64).pin_memory()
    sol_host = torch.empty(prob.N, dtype=torch.float64).pin_memory()
    dv, so = dir_host.numpy(), sol_host.numpy()
    barrier()
    t0 = time.perf_counter()
    e.timer_start()
    for _ in range(args.steps):
        tm += DT
        dv[: len(prob._dir_rows)] = prob.dirichlet_values(tm)
        e.step_host(False, tm, dv[: len(prob._dir_rows)], so)
    ms_e2e = e.timer_stop_ms()
    barrier()
    ms_e2e = reduce_max(max(ms_e2e, (time.perf_counter() - t0) * 1e3))
    clocks = sampler.stop() if rank == 0 else None
    log(f"e2e steps done: {ms_e2e / args.steps:.0f} ms/step")
    e2e = dict(value=n_dofs_global * args.steps / (ms_e2e * 1e-3), unit=UNIT,
               h2d_bytes_per_step=int(len(prob._dir_rows) * 8), d2h_bytes_per_step=int(prob.N * 8))

    # ---- roofline of the dominant kernel (isolated launches, L2 flushed between, CUDA events)
    peak, peak_src = measured_peaks()
    per_step = {"ilu_F": stats["cnt_ilu_F"], "spmv_F": stats["cnt_spmv_F"], "ilu_S": stats["cnt_ilu_S"],
                "spmv_S": stats["cnt_spmv_S"], "assemble_step": 1.0, "spmv_system": stats["n_vmult"]}
    kern = {}
    for name, cnt in per_step.items():
        kms, kbytes = e.bench_kernel(name, iters=5, flush_l2=True)
        kern[name] = dict(ms=kms, bytes=kbytes, gbs=kbytes / (kms * 1e-3) / 1e9, calls_last_step=cnt,
                          share_last_step=cnt * kms / (1e3 * (t_prec[-1] + t_solve[-1]) + 1e-9))
    dom = max(kern, key=lambda k: kern[k]["share_last_step"])
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    # (profiles/r01_traffic.json; only valid for the workload / rank count it was captured on)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tj = json.load(f).get(dom, {})
        if tj.get("workload") == args.workload and world == 1:
            traffic = tj.get("dram_bytes_per_apply", tj.get("dram_bytes_per_launch"))
    except Exception:
        pass
    roofline = dict(bound="hbm", kernel=dom, achieved=kern[dom]["gbs"], peak=peak, unit="GB/s",
                    frac=kern[dom]["gbs"] / peak, traffic=traffic, peak_source=peak_src,
                    kernels={k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                             for k, v in kern.items()})
    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
               ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
               data="synthetic", gpu_launches=int(launches),
               config=dict(workload=args.workload, mesh=f"cylinder3d(s={s}, nz={nz})", n_dofs=int(n_dofs_global),
                           n_cells=int(mesh.n_cells), variant="NavierStokes3D", preconditioner="Yosida",
                           deltat=DT, quadrature="QGaussSimplex(3) / Witherden-Vincent 14 pt",
                           ilu_ordering={0: "natural (reference replay)", 1: "multicolour (throughput mode)"}[args.ilu_ordering],
                           orthogonalisation={0: "modified Gram-Schmidt (reference replay)",
                                              1: "batched classical Gram-Schmidt (throughput mode)"}[args.orthogonalisation],
                           l2_policy="working set (>1 GB of matrices) exceeds the 126 MB L2; isolated kernel "
                                     "timings flush L2 between launches",
                           partition=f"{world} subdomain(s), coordinate bisection",
                           transport=(getattr(prob, "transport", "none") if world > 1 else "none")),
               e2e=e2e, roofline=roofline, clocks=clocks,
               detail=dict(outer_iterations=its, first_step_s=first_step_s, first_step_iterations=its_first,
                           first_step_converged=first_converged,
                           setup_s=setup_s, t_prec_s=t_prec, t_solve_s=t_solve, last_step_counts=stats,
                           levels={k: e.stat(k) for k in ("levels_F_fwd", "levels_F_bwd", "levels_S_fwd",
                                                           "levels_S_bwd")}))
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_cpu(args.cpu_sample, 1, 0)
        out["cpu_baseline"] = dict(value=cpu["value"], unit=UNIT, cores=cpu["cores"], kind="port",
                                   sample=cpu["sample"], seconds=cpu["seconds"])
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cyl3d-20M", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", default=CPU_SAMPLE, choices=sorted(WORKLOADS),
                    help="mesh of the bounded CPU sample (cpu_baseline leg and --impl reference)")
    ap.add_argument("--ilu-ordering", type=int, default=1, choices=[0, 1],
                    help="0: natural row order (reference replay), 1: multicolour ILU(0) (throughput mode, default)")
    ap.add_argument("--orthogonalisation", type=int, default=1, choices=[0, 1],
                    help="0: modified Gram-Schmidt as deal.II (reference replay), 1: batched classical Gram-Schmidt")
    ap.add_argument("--first-step-cap", type=int, default=560,
                    help="outer GMRES iterations allowed in the untimed first step (20 restart cycles; 0 = unlimited)")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        cpu = run_cpu(args.cpu_sample, args.steps, args.warmup)
        s, nz = WORKLOADS[args.workload]
        print(json.dumps(dict(
            impl="reference", metric=METRIC, value=cpu["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps,
            warmup=args.warmup, ms_per_step=cpu["seconds"] * 1e3 / args.steps, higher_is_better=True, scaling="strong",
            vs_baseline=None, dtype="f64", data="synthetic", gpu_launches=0,
            config=dict(workload=args.workload, mesh=f"cylinder3d(s={s}, nz={nz})", variant="NavierStokes3D",
                        preconditioner="Yosida", deltat=DT,
                        note="each step is a bounded sample of the workload (same mesh family, fewer DoFs)"),
            cpu_baseline=dict(value=cpu["value"], unit=UNIT, cores=cpu["cores"], kind="port", sample=cpu["sample"]),
            e2e=dict(value=cpu["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
            detail=dict(outer_iterations=cpu["iterations"]))))
        return
    run_gpu(args)


if __name__ == "__main__":
    main()
