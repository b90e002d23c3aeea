"""Regenerates tests/golden/setup_fingerprints.json: nsb_debug_setup_fingerprint of small meshes for every ILU ordering.
The file pins the device data structures setup builds on the host (patterns, scatter map, SELL / block-SELL storage, ILU
orderings); regenerate it only together with a change of those formats that the GPU parity tests have passed.
    python tests/golden/make_setup_fingerprints.py"""
import json
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [_ROOT, os.path.join(_ROOT, "tests")]
from test_setup_fingerprint import (CASES, LOCAL_CASES, LOCAL_ORDERINGS, ORDERINGS, fingerprint, fingerprint_local,
                                    owned_prefix)  # noqa: E402

out = {}
for name, make in CASES.items():
    mesh = make()
    for o1, o2 in ORDERINGS:
        out[f"{name}/{o1}/{o2}"] = fingerprint(mesh, o1, o2)
        if o1 != 3:
            out[f"{name}/{o1}/{o2}/ghosts"] = fingerprint(mesh, o1, o2, owned_prefix(mesh))
for name, nranks in LOCAL_CASES:
    mesh = CASES[name]()
    for rank in range(nranks):
        for o1, o2 in LOCAL_ORDERINGS:
            out[f"local/{name}/{nranks}/{rank}/{o1}/{o2}"] = fingerprint_local(mesh, nranks, rank, o1, o2)
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "setup_fingerprints.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print(len(out), "fingerprints written")
