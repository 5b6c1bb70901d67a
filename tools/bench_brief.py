"""One-screen summary of a bench.py JSON line.  usage: python tools/bench_brief.py file.json"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.lstrip().startswith("{")][-1])  # (NCCL may print a banner first)
print("value %.0f fps (%.3f ms/step)  e2e %.0f fps (%.3f ms/step)  launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("gpu_launches")))
print("repeats", [round(x, 3) for x in d["repeat_stats"]["ms_per_step"]], "e2e repeats", [round(x, 3) for x in d["e2e"]["repeats_ms_per_step"]])
print("kernels", {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["kernels"].items()})
print("dominant", d["roofline"]["kernel"], "frac %.4f" % d["roofline"]["frac"], "clocks", d.get("clocks"))
if d.get("tracking"):
    t = d["tracking"]
    print("tracking %.0f fps device, %.0f fps e2e, stages %s" % (t["value"], t["e2e"]["value"], {k: round(v, 3) for k, v in t["stage_ms"].items()}))
