#!/usr/bin/env python
"""One-line summary of a bench.py JSON line (A/B runs)."""
import json
import sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    st = d.get("stage_ms", {})
    print(sys.argv[2] if len(sys.argv) > 2 else "", "value %.1f" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.1f" % d["e2e"]["value"],
          "stages", {k[3:]: round(v, 1) for k, v in st.items()}, "launches", d.get("gpu_launches"), "clk", d.get("clocks", {}).get("sm_mhz"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
