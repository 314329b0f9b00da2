"""Print the headline fields of bench.py JSON lines: python tools/summ_bench.py a.json [b.json ...]"""
import json, sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d.get("impl", "ours"), "value", d["value"], "ms/step", d["ms_per_step"], "n_gpus", d["n_gpus"])
    if d.get("roofline"):
        r = d["roofline"]
        print("  roofline", r["kernel"], r["achieved"], "GB/s frac", r["frac"], "launch ms", r["avg_launch_ms"], "whole step", r.get("whole_step", {}).get("frac"))
    if d.get("e2e"):
        print("  e2e", {k: v for k, v in d["e2e"].items() if k in ("value", "d2h_gbs", "delivered_gbs", "pinned_d2h_copy_gbs", "host_threads", "d2h_bytes_per_step")})
    if d.get("cpu_baseline"):
        print("  cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "cores")
    print("  launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
    for s in d.get("secondary") or []:
        print("  ", s["workload"], s["value"], s.get("hbm_frac"))
