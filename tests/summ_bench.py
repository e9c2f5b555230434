"""Condenses bench.py JSON lines (files given as arguments, or stdin)."""
import json
import sys


def show(name, text):
    lines = [l for l in text.strip().splitlines() if l.startswith("{")]
    if not lines:
        print(name, "no JSON line")
        return
    d = json.loads(lines[-1])
    print(name, "%.1f GTEPS %.4f ms" % (d["value"] / 1e9, d["ms_per_step"]), "e2e %.1f GTEPS" % (d["e2e"]["value"] / 1e9))
    print("  ", d["config"].get("ms_per_superstep"), d["config"].get("superstep_mode"))
    if "roofline" in d:
        print("   roof frac %.3f %s" % (d["roofline"]["frac"], d["roofline"]["kernel"]))


if len(sys.argv) > 1:
    for f in sys.argv[1:]:
        show(f, open(f).read())
else:
    show("-", sys.stdin.read())
