import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
r=d["roofline"]; print("  kernel_ms", r["kernel_ms"], "achieved", r["achieved"], "frac", r["frac"], "ffma frac", r["frac_of_ffma_peak"], "nfev", r["nfev"])
print("  clocks", d["clocks"])
if "other_mode" in d:
    o=d["other_mode"]; print("other", o["mlp_mode"], o["value"], o["ms_per_step"], o["roofline"]["kernel_ms"], o["roofline"]["frac"])
print("cpu", d.get("cpu_baseline"))
