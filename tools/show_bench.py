"""Print the headline numbers of one or more bench.py JSON lines (files given on the command line)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as ex:
        print(path, "unreadable:", ex)
        continue
    r = d.get("roofline", {})
    stages = {k: (round(v["ms"], 1), round(v["share"], 3)) for k, v in r.get("stages", {}).items()}
    print("%s: %.2f Mmut/s  %.0f Mrays/s  chains %s  rounds/step %s  stages %s  e2e %.2f Mmut/s  cpu %s" % (
        path, d["value"] / 1e6, r.get("mrays_per_s", 0), d["config"].get("chains_per_gpu"), r.get("rounds_per_step"), stages,
        d["e2e"]["value"] / 1e6, (d.get("cpu_baseline") or {}).get("value")))
