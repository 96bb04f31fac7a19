"""Row-sharing weight gradient (wgrad_rowshare_kernel) against the generic pair kernel and a float64 reference on the
3x3 256->256 layer at 64x64 (x materialised-padded to 66x66), then timings of both."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
from cycle_depth_estimation_b200 import ops
torch.manual_seed(0)
n = int(sys.argv[1])
x = torch.randn((n, 66, 66, 256), device="cuda").to(torch.bfloat16)
dy = torch.randn((n, 64, 64, 256), device="cuda").to(torch.bfloat16)
dw = torch.empty((256, 256, 3, 3), device="cuda")
ops.conv2d_wgrad(ops.geom(3, 3), x, dy, dw, False)
torch.cuda.synchronize()
if n <= 2:
    xf, dyf = x.double().permute(0, 3, 1, 2), dy.double().permute(0, 3, 1, 2)
    ref = torch.nn.grad.conv2d_weight(xf, (256, 256, 3, 3), dyf)
    print("rel l2 vs float64: %%.3e" %% float((dw.double() - ref).norm() / ref.norm()))
torch.save(dw.cpu(), sys.argv[2])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.conv2d_wgrad(ops.geom(3, 3), x, dy, dw, False)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print("batch %%d: %%.1f us  %%.0f TFLOP/s" %% (n, us, 2.0 * n * 64 * 64 * 256 * 256 * 9 / us / 1e6))
''' % ROOT
import torch
for n in (2, 8, 16, 24):
    outs = []
    for mode in ("1", "0"):
        path = "/tmp/dw_%s.pt" % mode
        r = subprocess.run([sys.executable, "-c", code, str(n), path], env=dict(os.environ, CDB_WGRAD_ROWSHARE=mode),
                           capture_output=True, text=True)
        print("rowshare=%s |" % mode, " | ".join(r.stdout.strip().splitlines()) or r.stderr[-600:], flush=True)
        outs.append(path)
    try:
        a, b = torch.load(outs[0]), torch.load(outs[1])
        print("   rowshare vs generic: rel l2 %.3e" % float((a.double() - b.double()).norm() / b.double().norm()))
    except Exception as e:
        print("   compare failed", e)
