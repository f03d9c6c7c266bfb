#!/usr/bin/env python
"""Print a kernel_table.json as a compact table (optionally next to a second one for comparison)."""
import json
import sys


def load(p):
    return {(r["case"], r["dtype"]): r for r in json.load(open(p))["rows"]}


a = load(sys.argv[1])
b = load(sys.argv[2]) if len(sys.argv) > 2 else None
for key, r in a.items():
    s = f"{key[0]:18s} {key[1]:5s}"
    for k in ("merge_a", "merge_b", "unmerge"):
        if k in r:
            s += f" | {k[:7]:7s} {r[k]['ms']:7.3f}ms {r[k]['GBps']:6.0f}GB/s"
            if b and key in b and k in b[key]:
                s += f" ({b[key][k]['GBps']:6.0f})"
        else:
            s += " | " + " " * (31 + (9 if b else 0))
    s += f" | dot {r['dot']['ms']:8.3f}ms {r['dot']['TFLOPs']:6.2f}TF"
    if b and key in b:
        s += f" ({b[key]['dot']['TFLOPs']:6.2f})"
    print(s)
