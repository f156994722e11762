#!/usr/bin/env python
"""print the step breakdown of a bench.py JSON line: tools/bench_breakdown.py gpurun_out/x.json"""
import json, sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e", d["e2e"].get("value") and round(d["e2e"]["value"], 1), "gemm frac", round(r["frac"], 3),
          "clk", d["clocks"]["sm_mhz"] if d.get("clocks") else None)
    print("  ", {k: v for k, v in r["step_breakdown_ms"].items() if v >= 0.05})
    print("  ", r["attention_tflops"])
