"""Prints the headline fields of a bench.py JSON line read from stdin. usage: python bench.py ... | python tools/print_bench.py [tag]"""
import json, sys
lines = [l for l in sys.stdin.read().strip().splitlines() if l.startswith("{")]
d = json.loads(lines[-1])
tag = sys.argv[1] if len(sys.argv) > 1 else ""
print(tag, "ms/step", round(d["ms_per_step"], 2) if d.get("ms_per_step") else None, "wall", round(d.get("wall_ms_per_step", 0), 2), "value M", round(d["value"] / 1e6, 2),
      "e2e M", round(d["e2e"]["value"] / 1e6, 2), "frac", round(d["roofline"]["frac"], 3) if d.get("roofline") else None, "launches/step", d.get("gpu_launches", 0) // max(d.get("steps", 1), 1))
