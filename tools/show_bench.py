import json,sys
d=json.load(open(sys.argv[1]))
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]), "parse core-s/step", round(d["e2e"]["host_parse_core_seconds_per_step"],2), "ms/step", round(d["e2e"]["ms_per_step"]))
for k,v in d["roofline"]["kernels"].items(): print("  ",k, "ms/launch",round(v["ms_per_launch"],3),"GB/s",round(v["GBps"]),"frac",round(v["frac_of_peak"],4),"share",round(v["share_of_step"],3))
