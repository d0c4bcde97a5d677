import sys, json
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if d.get("impl") == "reference":
        print("REFERENCE value", d["value"], d["cpu_baseline"]); continue
    print("value %.4g rows/s  ms/step %.3f  e2e %.4g  ratio %.3f  sel %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["lz4_ratio"], d["config"]["selectivity"]))
    print(" kernels", d["roofline"]["all_kernels_ms"])
    print(" roofline", d["roofline"]["kernel"], "%.1f GB/s frac %.4f" % (d["roofline"]["achieved"], d["roofline"]["frac"]), "clocks", d["clocks"], "launches", d["gpu_launches"])
    if "cpu_baseline" in d: print(" cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
