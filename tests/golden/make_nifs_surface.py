#!/usr/bin/env python
"""Writes tests/golden/nifs_surface.json: the (name, arity) surface of the reference's `Vettore.Nifs`
(parsed out of /root/reference/lib/vettore_nifs.ex) and the subset on the scan path that the erl_nif shim
(nif/vettore_b200_nif.c) must export with identical names and arities. Run in the build
container (the reference tree is not on the GPU box); the JSON is the committed fixture."""
import json
import os
import re

REF = "/root/reference/lib/vettore_nifs.ex"
HERE = os.path.dirname(os.path.abspath(__file__))

src = open(REF).read()
surface = []
for m in re.finditer(r"^\s*def\s+([a-z_0-9]+)(\(([^)]*)\))?\s*,?\s*(do:|$)", src, re.M):
    name, args = m.group(1), (m.group(3) or "").strip()
    arity = 0 if not args else len([a for a in args.split(",") if a.strip()])
    surface.append([name, arity])
# multi-line heads: def name(\n  _a,\n  _b ... )
for m in re.finditer(r"^\s*def\s+([a-z_0-9]+)\(\s*\n(.*?)\)\s*,?\s*\n?\s*do:", src, re.M | re.S):
    name = m.group(1)
    arity = len([a for a in m.group(2).split(",") if a.strip()])
    if [name, arity] not in surface:
        surface.append([name, arity])
scan_prefixes = ("flat_", "vector_top_k", "binary_top_k", "multi_vector_", "compress_sign_bits", "muvera_")   # + SURVEY.md §8(f) rank 4
scan = [e for e in surface if e[0].startswith(scan_prefixes)]
out = {"source": "lib/vettore_nifs.ex (reference v0.3.2)", "all": sorted(surface), "scan_path": sorted(scan)}
json.dump(out, open(os.path.join(HERE, "nifs_surface.json"), "w"), indent=1)
print(len(surface), "functions,", len(scan), "on the scan path")
