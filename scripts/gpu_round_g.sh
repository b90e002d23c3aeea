#!/bin/bash
# GPU session G (1 GPU): tests with the persistent solves; A/B of the persistent kernels at 2M; bench at 2M;
# ncu --set full of the top kernels at 20M (small captures); ncu launch list of a 2M bench run
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
tail -6 gpurun_out/g_pytest.log
for p in 1 0; do
  echo "== 2M persistent=$p"; NSB_TRSV_PERSISTENT=$p timeout 300 python scripts/prof_kernels.py cyl3d-2M 1 10 ilu_F,ilu_S,spmv_F,spmv_S 2>&1 | tee gpurun_out/g_prof_2M_pers$p.log
done
timeout 600 python bench.py --workload cyl3d-2M --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_2M.json 2> gpurun_out/g_bench_2M.err
python - <<PY
import json
d=json.load(open("gpurun_out/g_bench_2M.json"))
print("2M", d["value"], d["ms_per_step"], d["detail"]["outer_iterations"], d["detail"]["first_step_iterations"], d["e2e"]["value"])
print({k:(v["ms"],v["gbs"]) for k,v in d["roofline"]["kernels"].items()})
PY
# ncu (each command first runs plain on the box, so keep them short)
timeout 800 ncu --set full --import-source on --clock-control none -k regex:"k_sell3" --launch-skip 102 -c 34 \
  -o gpurun_out/g_ncu_ilu -f python scripts/prof_kernels.py cyl3d-20M 1 1 ilu_F > gpurun_out/g_ncu_ilu.log 2>&1
tail -2 gpurun_out/g_ncu_ilu.log
timeout 800 ncu --set full --import-source on --clock-control none -k regex:"k_sell3|assemble_step_t" --launch-skip 6 -c 2 \
  -o gpurun_out/g_ncu_spmv_asm -f python scripts/prof_kernels.py cyl3d-20M 1 1 assemble_step,spmv_F > gpurun_out/g_ncu_spmv_asm.log 2>&1
tail -2 gpurun_out/g_ncu_spmv_asm.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60000 -c 40000 --csv \
  --log-file gpurun_out/g_launches.csv python bench.py --workload cyl3d-2M --steps 1 --warmup 3 --no-cpu-baseline --first-step-cap 56 \
  > gpurun_out/g_ncu_launches.log 2>&1
tail -2 gpurun_out/g_ncu_launches.log | cut -c1-300
python - <<'PY'
import csv, collections, os
if os.path.exists("gpurun_out/g_launches.csv"):
    rows = list(csv.reader(l for l in open("gpurun_out/g_launches.csv") if not l.startswith("==")))
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= vi: continue
        t = float(r[vi].replace(",", ""))
        t = t / 1000.0 if r[ui] in ("ns", "nsecond") else (t * 1000.0 if r[ui] in ("ms", "msecond") else t)
        a = agg[r[ki].split("(")[0]]; a[0] += 1; a[1] += t
    tot = sum(v[1] for v in agg.values())
    with open("gpurun_out/g_launch_summary.csv", "w") as f:
        f.write("kernel,launches,total_us,share\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k}",{v[0]},{v[1]:.1f},{v[1]/tot:.4f}\n')
    print(open("gpurun_out/g_launch_summary.csv").read())
    os.system("gzip -f gpurun_out/g_launches.csv")
ls = os.popen("du -sh gpurun_out; ls -la gpurun_out").read(); print(ls)
PY
