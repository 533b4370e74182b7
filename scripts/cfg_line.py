"""Condense scripts/bench_configs.py JSON lines (stdin) into one readable line each."""
import json
import sys

for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        print(d["case"], "fwd %.3f adj %.3f ms" % (d["fwd_ms"], d["adj_ms"]), "frac %.3f" % d["step_frac"] if "step_frac" in d else "",
              "tiles %s" % d.get("tiles"), d.get("options", ""), flush=True)
    elif "rror" in l:
        print(l.rstrip(), flush=True)
