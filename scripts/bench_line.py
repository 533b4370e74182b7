"""Condense bench.py's JSON line (stdin) into readable lines. Usage: python bench.py ... | python scripts/bench_line.py LABEL"""
import json
import sys

label = " ".join(sys.argv[1:])
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        k = d["roofline"]["kernels"]
        g = d.get("general_path") or {}
        print("%s | N=%d fwd %.3f adj %.3f step %.3f ms  %.0f Melem/s  frac %.3f | general fwd %.3f adj %.3f | e2e %s ms | cpu %s" % (
            label, d["n_gpus"], k["fwd"]["ms"], k["adj"]["ms"], d["ms_per_step"], d["value"], d["roofline"]["step_frac"], g.get("fwd_ms", 0), g.get("adj_ms", 0),
            (d.get("e2e") or {}).get("ms_per_step"), (d.get("cpu_baseline") or {}).get("value")), flush=True)
        for x in (d.get("extra") or {}).get("configs", []):
            if "error" in x:
                print("   %s ERROR %s" % (x["case"], x["error"]), flush=True)
            else:
                print("   %s N=%d E=%d fwd %.3f adj %.3f step %.3f ms  %.0f Melem/s  step_frac %.3f (fwd %.3f adj %.3f) plan %.0f B/elem setup %.1fs" % (
                    x["case"], x["n_gpus"], x["elements_total"], x["fwd_ms"], x["adj_ms"], x["ms_per_step"], x["Melem_per_s"], x["roofline"]["step_frac"],
                    x["roofline"]["fwd_frac"], x["roofline"]["adj_frac"], x["plan_bytes_per_elem"], x["setup_s_untimed"]), flush=True)
    elif "rror" in l or "Traceback" in l:
        print(l.rstrip(), flush=True)
