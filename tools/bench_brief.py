"""Condense a bench.py JSON line (stdin) to the numbers watched while tuning."""
import sys, json
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    c = d.get("config", {})
    print(c.get("workload"), "ms/step", round(d["ms_per_step"], 2), "P1", c.get("partitions"), "frac", round(d["roofline"]["frac"], 4),
          {a: round(b, 2) for a, b in d.get("phases_ms", {}).items()}, "e2e ms", round(d["e2e"]["ms_per_step"], 2), "launches", d.get("gpu_launches"))
