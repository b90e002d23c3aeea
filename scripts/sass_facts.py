"""profiles/r02/sass_facts.txt: per-kernel counts of the SASS mnemonics that carry the design,
from `cuobjdump -sass libnsb.so`.  Run from the repo root after `make -C navierstokes_project_nm4pde_b200/csrc`."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "navierstokes_project_nm4pde_b200", "libnsb.so")
PATS = {"LDG.E.ENL2.256 (256-bit gathers)": r"LDG\.E\.ENL2\.256", "REDG.E.ADD.F64 (FP64 reductions to global)": r"REDG?\.E\.ADD\.F64",
        "PREEXIT (griddepcontrol.launch_dependents)": r"\bPREEXIT\b", "ACQBULK (griddepcontrol.wait)": r"\bACQBULK\b",
        "CCTL.E.PF2 (prefetch.global.L2)": r"CCTL\.E\.PF2", "UBLKCP (cp.async.bulk)": r"UBLKCP", "SYNCS (mbarrier)": r"\bSYNCS\b",
        "LDG.E.EF (evict-first streaming loads)": r"LDG\.E\.EF", "DFMA": r"\bDFMA\b", "SHFL": r"\bSHFL\b",
        "STG/LDG.STRONG.SYS (peer-memory flags and slots)": r"(STG|LDG)\.E(\.64)?\.STRONG\.SYS", "MEMBAR.SYS": r"MEMBAR\.[A-Z]+\.SYS",
        "HMMA/UTCMMA (tensor cores)": r"\b(HMMA|UTCMMA|IMMA|DMMA)\b"}
KEEP = ["k_sell3<0, 8, false, 1>", "k_sell3<1, 8, true, 1>", "k_sell3<2, 8, true, 1>", "k_bsell<3, 0, false, false>",
        "k_bsell<3, 1, false, false>", "k_bsell<1, 0, true, true>", "k_stream<1, 1>", "k_stream<1, 2>", "assemble_step_t_kernel<3>",
        "assemble_first_kernel<3>", "k_sd_trsv<3, 0>", "k_multi_dot<8>", "k_halo_push<3>", "k_halo_wait", "k_allreduce_p2p",
        "k_face_forces<3>"]

txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
out = ["# SASS facts of libnsb.so (sm_100a): `cuobjdump -sass navierstokes_project_nm4pde_b200/libnsb.so`, instruction counts per kernel of",
       "# the mnemonics that carry the design (DESIGN.md section 4).  Regenerate: python scripts/sass_facts.py", "",
       f"{len(funcs)} kernels in the library.", ""]
tot = collections.Counter()
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(.*", "", dem).replace("void ", "").replace("nsb::", "")
    cnt = {k: len(re.findall(p, f)) for k, p in PATS.items()}
    tot.update(cnt)
    if any(short == k for k in KEEP):
        n_inst = len(re.findall(r"/\*[0-9a-f]{4}\*/", f))
        out.append(f"{short}: {n_inst} instructions; " + ", ".join(f"{k.split(' ')[0]} {v}" for k, v in cnt.items() if v))
out += ["", "whole library: " + "; ".join(f"{k}: {v}" for k, v in tot.items())]
path = os.path.join(ROOT, "profiles", "r02", "sass_facts.txt")
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out))
