#!/bin/bash
# GPU session F (1 GPU): full GPU test suite, ncu --set full of the top kernels, ncu launch list of a short bench,
# then the default bench (never under ncu)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -4 gpurun_out/f_pytest.log
timeout 900 python bench.py > gpurun_out/f_bench_20M.json 2> gpurun_out/f_bench_20M.err
cat gpurun_out/f_bench_20M.json; tail -3 gpurun_out/f_bench_20M.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
cat gpurun_out/f_bench_ref.json
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_sell3|assemble_step_t" -c 76 \
  -o gpurun_out/f_ncu_top -f python scripts/prof_kernels.py cyl3d-20M 1 1 assemble_step,spmv_F,ilu_F > gpurun_out/f_ncu_top.log 2>&1
tail -3 gpurun_out/f_ncu_top.log
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 150000 -c 30000 --csv \
  --log-file gpurun_out/f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --first-step-cap 28 \
  > gpurun_out/f_ncu_launches.log 2>&1
tail -2 gpurun_out/f_ncu_launches.log | cut -c1-300
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/f_launches.csv") if not l.startswith("==")))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= vi: continue
    t = float(r[vi].replace(",", "")); t = t / 1000.0 if r[ui] in ("ns", "nsecond") else t
    a = agg[r[ki].split("(")[0]]; a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
with open("gpurun_out/f_launch_summary.csv", "w") as f:
    f.write("kernel,launches,total_us,share\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f'"{k}",{v[0]},{v[1]:.1f},{v[1]/tot:.4f}\n')
print(open("gpurun_out/f_launch_summary.csv").read())
PY
