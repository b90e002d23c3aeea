"""Regenerates profiles/r02/resource_usage.txt: registers / stack (spills) / static shared memory per kernel of
libnsb.so from `cuobjdump --dump-resource-usage`.   python scripts/resource_usage.py"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "navierstokes_project_nm4pde_b200", "libnsb.so")
txt = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True).stdout
names = re.findall(r"Function (\S+):\n\s*(REG:\d+[^\n]*)", txt)
dem = subprocess.run(["c++filt"], input="\n".join(n for n, _ in names), capture_output=True, text=True).stdout.splitlines()
rows = []
for (n, r), d in zip(names, dem):
    d = re.sub(r"\(.*\)$", "", d.replace("void ", "").replace("nsb::", "").replace("(anonymous namespace)::", ""))
    keep = " ".join(f for f in r.split() if f.split(":")[0] in ("REG", "STACK", "SHARED", "LOCAL"))
    rows.append((d, keep))
with open(os.path.join(ROOT, "profiles", "r02", "resource_usage.txt"), "w") as f:
    f.write("# Registers / stack (spills) / static shared memory per kernel of libnsb.so (sm_100a):\n"
            "# cuobjdump --dump-resource-usage, names by c++filt.  The occupancy statements of profiles/README.md and\n"
            "# DESIGN.md rest on these (65 536 registers per SM).  Regenerate: python scripts/resource_usage.py\n\n")
    for d, r in sorted(rows):
        f.write(f"{d:64s} {r}\n")
print(len(rows), "kernels")
