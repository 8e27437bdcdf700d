"""Print the headline fields of a bench.py JSON line."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.0f  ms/step %.4f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
r = d["roofline"]
print("kernel ms %.4f  frac %.3f   step frac %.3f   writer frac %.3f" % (
    r["ms_per_launch"], r["frac"], d["roofline_step"]["frac"], d["roofline_grid_writer"]["frac"]))
c4 = d.get("config4_saturated_cloud") or {}
print("config4 ms %.4f" % c4.get("ms_per_cloud", -1), " full_inference %.0f sweeps/s" % d["full_inference"]["value"])
print("clocks", d["clocks"])
