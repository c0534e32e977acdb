import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2] if len(sys.argv) > 2 else "", round(d["ms_per_step"], 3), round(d["kernels_ms_per_step"].get("tensor_filter", 0), 3),
      round(d["roofline"]["frac_sustained"], 3) if d.get("roofline") else None, d["verified"]["identical"] if d.get("verified") else None,
      d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["counters"])
