"""CTA-0 timeline of igemm_flat_kernel (CDB_FLAT_DEBUG=1) on the R256 shape: last lines only."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
from cycle_depth_estimation_b200 import ops
n = int(sys.argv[1])
x = torch.randn((n, 66, 66, 256), device="cuda").to(torch.bfloat16)
w = (torch.randn((256, 256, 3, 3), device="cuda") * 0.02).contiguous()
wp, rp, kp = ops.pack_conv_weight(w, True)
y = ops.alloc_flat_output(n, 64, 64, 66, 256, "cuda")
st = torch.zeros((n, 256, 2), device="cuda")
for _ in range(6):
    ops.conv2d_fwd(ops.geom(3, 3), x, wp, rp, kp, ops.out_view_nhwc(y, 256), None, 0, 0.0, st if sys.argv[2] == "1" else None)
torch.cuda.synchronize()
''' % ROOT
for env in sys.argv[1:] or [""]:
    e = dict(os.environ, CDB_FLAT_DEBUG="1")
    for kv in env.split(","):
        if kv:
            k, v = kv.split("=")
            e[k] = v
    for n in (8, 16):
        for st in ("1", "0"):
            r = subprocess.run([sys.executable, "-c", code, str(n), st], env=e, capture_output=True, text=True)
            lines = [l for l in r.stderr.splitlines() if "flat dbg" in l]
            print(env, "batch", n, "stats", st, "|", lines[-1] if lines else r.stderr[-300:], flush=True)
