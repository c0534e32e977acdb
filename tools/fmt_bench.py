"""Reads bench.py's JSON line on stdin and prints the headline fields."""
import json, sys
for l in sys.stdin:
    try:
        d = json.loads(l)
    except Exception:
        print(l.strip()[:200]); continue
    r = d.get("roofline") or {}
    print(d["config"]["workload"], "| value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]),
          "| roof", r.get("kernel"), round(r.get("frac", 0), 3), "| launches", d.get("gpu_launches"), "| clocks", d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons"))
    print("   ", {k: round(v, 3) for k, v in d.get("kernels_ms_per_step", {}).items()}, d.get("counters"), d.get("cpu_baseline"))
