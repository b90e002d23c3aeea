"""profiles/r02/sass_excerpts.txt: the hot loops of the kernels that carry a time step, cut out of `cuobjdump -sass libnsb.so`
(the innermost backward branch with the most DFMA / gather instructions of each kernel, opcodes only).
Run from the repo root after `make -C navierstokes_project_nm4pde_b200/csrc`."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "navierstokes_project_nm4pde_b200", "libnsb.so")
KEEP = {"k_sell3<0, 8, false, 1>": "SpMV F_s (SELL-32, one thread per row, pipeline depth 8)",
        "k_sell3<1, 8, true, 1>": "forward colour sweep of the point multicolour ILU(0) apply",
        "k_bsell<3, 0, false, false>": "forward sweep of the block multicolour ILU(0) apply (dominant kernel at 19.9 M DoF on one GPU)",
        "k_stream<1, 1>": "forward colour sweep of the ILU(0) apply of the pressure Schur complement",
        "assemble_step_t_kernel<3>": "step assembly, tensor-contracted"}

txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
out = ["# Hot loops of libnsb.so (sm_100a), from `cuobjdump -sass`: per kernel the loop (backward branch target .. branch) with the most",
       "# FP64 FMAs + global loads, addresses and encodings stripped.  Regenerate: python scripts/sass_excerpts.py", ""]
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(.*", "", dem).replace("void ", "").replace("nsb::", "")
    if short not in KEEP:
        continue
    ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4})\*/\s+(.*?);", f)]
    addr = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a or tgt not in addr:
            continue
        body = ins[addr[tgt]: i + 1]
        # the arithmetic loop, not a staging loop: FP64 FMAs and 256-bit gathers weigh more than plain loads
        score = sum(8 if re.search(r"\bDFMA\b|ENL2\.256|\bREDG?\b", x) else 1 if re.search(r"\bLDG\b", x) else 0 for _, x in body)
        if score and (best is None or score > best[0] or (score == best[0] and len(body) < len(best[1]))):
            best = (score, body)
    out.append(f"## {short} -- {KEEP[short]}: {len(ins)} instructions in the kernel")
    if best is None:
        out.append("(no loop found: fully unrolled)")
        body = [x for x in ins if re.search(r"\bDFMA\b|\bLDG\b|\bREDG?\b|\bSHFL\b", x[1])][:60]
    else:
        body = best[1]
        out.append(f"loop of {len(body)} instructions at 0x{body[0][0]:04x}:")
    shown = body if len(body) <= 90 else body[:60] + [(0, f"... {len(body) - 80} instructions ...")] + body[-20:]
    for _, t in shown:
        out.append("    " + re.sub(r"\s+", " ", t))
    out.append("")
path = os.path.join(ROOT, "profiles", "r02", "sass_excerpts.txt")
open(path, "w").write("\n".join(out) + "\n")
print(f"{len(out)} lines -> {path}")
