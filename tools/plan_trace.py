#!/usr/bin/env python
"""Where does plan creation time go inside a real sweep?  Wraps CopyPlan / GemmPlan / EwPlan creation with a host timer
during a short DMRG run and prints the distribution and the slowest plans with their shape information."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from yastn_b200 import plans  # noqa: E402

log = {"CopyPlan": [], "GemmPlan": [], "EwPlan": []}


def wrap(cls_name):
    cls = getattr(plans, cls_name)
    init = cls.__init__

    def timed(self, *a, **k):
        t = time.perf_counter()
        init(self, *a, **k)
        dt = (time.perf_counter() - t) * 1e6
        info = self.info() if hasattr(self, "info") else {}
        log[cls_name].append((dt, len(log[cls_name]), info))
    cls.__init__ = timed


for n in log:
    wrap(n)
sys.argv = ["dmrg_bench.py"] + sys.argv[1:]
import runpy  # noqa: E402
try:
    runpy.run_path(os.path.join(ROOT, "tools", "dmrg_bench.py"), run_name="__main__")
except SystemExit:
    pass
for name, rows in log.items():
    if not rows:
        continue
    dts = np.array([r[0] for r in rows])
    print(json.dumps({"plan": name, "count": len(rows), "total_ms": round(dts.sum() / 1e3, 1), "median_us": round(float(np.median(dts)), 1),
                      "p90_us": round(float(np.percentile(dts, 90)), 1), "max_us": round(float(dts.max()), 1),
                      "first20_total_ms": round(dts[:20].sum() / 1e3, 1),
                      "slowest": [{"us": round(r[0], 1), "seq": r[1], "info": r[2]} for r in sorted(rows, key=lambda r: -r[0])[:6]]}))
