"""Timeline of one graph-replayed CycleGAN step (CUPTI through torch.profiler): per stream busy time, the union of all
kernel intervals (time with at least one kernel running), the time with two or more kernels running, idle gaps, and the
kernels around the largest gaps.  Answers: is the step bound by kernel time (union ~ step) or by dependencies / launch
latency (large idle share)?   usage: python tools/step_timeline.py [out.json]"""
import contextlib, io, json, os, random, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from cycle_depth_estimation_b200.cycle_gan_model import CycleGANModel

torch.manual_seed(0); random.seed(1234)
model = CycleGANModel()
with contextlib.redirect_stdout(io.StringIO()):
    model.initialize(bench.make_opt("cuda", True))
a, b = bench.synthetic_batch(8, 256, 1234)
dev = {"img_source": a.cuda(), "img_target": b.cuda()}
for _ in range(4):
    model.set_input(dev); model.optimize_parameters("train")
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.set_input(dev); model.optimize_parameters("train")
    torch.cuda.synchronize()
path = "/tmp/step_trace.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
per_stream = collections.defaultdict(float)
for e in ev:
    per_stream[e["args"].get("stream")] += e["dur"]
# sweep
pts = []
for e in ev:
    pts.append((e["ts"], 1)); pts.append((e["ts"] + e["dur"], -1))
pts.sort()
busy1 = busy2 = 0.0
depth, last = 0, pts[0][0]
gaps = []
for t, d in pts:
    if depth >= 1: busy1 += t - last
    if depth >= 2: busy2 += t - last
    if depth == 0 and t - last > 0: gaps.append((t - last, last))
    depth += d; last = t
res = {"kernels": len(ev), "span_us": t1 - t0, "sum_kernel_us": sum(e["dur"] for e in ev), "busy_any_us": busy1,
       "busy_two_or_more_us": busy2, "idle_us": (t1 - t0) - busy1,
       "per_stream_us": {str(k): round(v, 1) for k, v in sorted(per_stream.items(), key=lambda kv: -kv[1])}}
gaps.sort(reverse=True)
res["gaps_over_2us"] = sum(1 for g in gaps if g[0] > 2.0)
res["gap_us_total_over_2us"] = round(sum(g[0] for g in gaps if g[0] > 2.0), 1)
res["largest_gaps"] = []
for g, at in gaps[:12]:
    before = max((e for e in ev if e["ts"] + e["dur"] <= at + 0.01), key=lambda e: e["ts"] + e["dur"], default=None)
    after = min((e for e in ev if e["ts"] >= at + g - 0.01), key=lambda e: e["ts"], default=None)
    res["largest_gaps"].append({"gap_us": round(g, 1), "after": before["name"][:50] if before else None,
                                "before": after["name"][:50] if after else None})
# gap histogram by the kernel that follows
by_next = collections.defaultdict(lambda: [0, 0.0])
for g, at in gaps:
    after = min((e for e in ev if e["ts"] >= at + g - 0.01), key=lambda e: e["ts"], default=None)
    if after:
        k = after["name"].split("(")[0].replace("void ", "")[:40]
        by_next[k][0] += 1; by_next[k][1] += g
res["idle_before_kernel"] = {k: [v[0], round(v[1], 1)] for k, v in sorted(by_next.items(), key=lambda kv: -kv[1][1])[:12]}
print(json.dumps(res, indent=1))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
