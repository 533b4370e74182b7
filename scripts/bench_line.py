"""Condense bench.py's JSON line (stdin) into one readable line. Usage: python bench.py ... | python scripts/bench_line.py LABEL"""
import json
import sys

label = " ".join(sys.argv[1:])
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        k = d["roofline"]["kernels"]
        g = d.get("general_path") or {}
        print("%s | fwd %.3f adj %.3f step %.3f ms  frac %.3f | general fwd %.3f adj %.3f | e2e %s" % (
            label, k["fwd"]["ms"], k["adj"]["ms"], d["ms_per_step"], d["roofline"]["step_frac"], g.get("fwd_ms", 0), g.get("adj_ms", 0),
            (d.get("e2e") or {}).get("ms_per_step")), flush=True)
    elif "rror" in l or "Traceback" in l:
        print(l.rstrip(), flush=True)
