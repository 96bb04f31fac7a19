"""Per-role wait cycles of CTA 0 of igemm_kernel (CDB_IGEMM_DEBUG=1) on the transposed 128->64 layer (four parity launches)
and the stride-2 64->128 layer, batch 8: which of TMA producer / MMA issuer / epilogue paces the kernel."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
from cycle_depth_estimation_b200 import ops
n = 8
def run(ci, co, h, k, stride, pad, transposed, stats):
    x = torch.randn((n, h, h, ci), device='cuda').to(torch.bfloat16)
    if transposed:
        w = (torch.randn((ci, co, k, k), device='cuda') * 0.02).contiguous(); ho = (h - 1) * stride - 2 * pad + k + 1
    else:
        w = (torch.randn((co, ci, k, k), device='cuda') * 0.02).contiguous(); ho = (h + 2 * pad - k) // stride + 1
    wp, rows, kpad = ops.pack_conv_weight(w, not transposed)
    y = torch.empty((n, ho, ho, co), dtype=torch.bfloat16, device='cuda')
    st = torch.zeros((n, co, 2), device='cuda') if stats else None
    g = ops.geom(k, k, stride, pad, pad, 1, transposed)
    for _ in range(3):
        ops.conv2d_fwd(g, x, wp, rows, kpad, ops.out_view_nhwc(y, co), None, 0, 0.0, st)
    torch.cuda.synchronize()
which = sys.argv[1]
if which == 'convT': run(128, 64, 128, 3, 2, 1, True, False)
if which == 'convT_stats': run(128, 64, 128, 3, 2, 1, True, True)
if which == 'd128': run(64, 128, 256, 3, 2, 1, False, True)
if which == 'd256': run(128, 256, 128, 3, 2, 1, False, True)
''' % ROOT
for which in ('convT', 'convT_stats', 'd128', 'd256'):
    e = dict(os.environ, CDB_IGEMM_DEBUG="1")
    r = subprocess.run([sys.executable, "-c", code, which], env=e, capture_output=True, text=True)
    lines = [l for l in r.stderr.splitlines() if "igemm dbg" in l]
    print(which)
    for l in lines[-4:] if 'convT' in which else lines[-1:]:
        print("   ", l)
    if not lines:
        print(r.stderr[-400:])
