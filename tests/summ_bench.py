import json,sys
for f in sys.argv[1:]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "%.1f GTEPS %.4f ms" % (d["value"]/1e9, d["ms_per_step"]), "e2e %.1f GTEPS" % (d["e2e"]["value"]/1e9))
    print("  ", d["config"]["ms_per_superstep"], d["config"]["superstep_mode"])
    print("   roof frac %.3f %s" % (d["roofline"]["frac"], d["roofline"]["kernel"]))
