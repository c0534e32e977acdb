"""Reads sweep_c5.py's JSON lines on stdin and prints one compact line per point."""
import json, sys
for l in sys.stdin:
    try:
        r = json.loads(l)
    except Exception:
        print(l.strip()[:200]); continue
    km = {k: round(v, 3) for k, v in r["kernels_ms"].items() if v > 0.05}
    print(r["dim"], r["nq"], r["path"], round(r["ms"], 3), int(r["qps"]), "hbm", round(r["hbm_frac"], 3), "tc", round(r["tensor_frac"], 3), r["counters"], km)
