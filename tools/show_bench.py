"""Print the key numbers of bench.py JSON lines found in the given log files."""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f):
        line = line.strip()
        if not line.startswith("{"):
            continue
        try:
            d = json.loads(line)
        except Exception:
            continue
        r = d.get("roofline") or {}
        e = d.get("e2e") or {}
        print("%s: ms/step=%.3f value=%.4g kernel_ms=%s achieved=%s frac=%s e2e_ms=%s launches=%s clocks=%s" % (
            f, d.get("ms_per_step", 0), d.get("value", 0), r.get("kernel_ms_per_step"), r.get("achieved"), r.get("frac"),
            e.get("ms_per_step"), d.get("gpu_launches"), d.get("clocks")))
