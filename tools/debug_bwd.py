"""Stage-by-stage backward comparison of the fused engine against torch autograd (debug aid)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import quiet, rel_l2, seeded_image, true_fp32  # noqa: E402
from cycle_depth_estimation_b200 import engine, networks as N  # noqa: E402


def main():
    torch.manual_seed(1)
    with quiet():
        net = N.define_D(3, 64, 'basic', 3, 'instance', False, 'normal', 0.02, ['cuda'])
    x = seeded_image(2, 3, 128, 128).requires_grad_(True)
    gout = seeded_image(2, 1, 14, 14, seed=9)
    engine.DEBUG_RECORD = {}
    out = net(x)
    (out * gout).sum().backward()
    rec = engine.DEBUG_RECORD
    sd = net.state_dict()
    # reference with intermediates
    xr = x.detach().clone().requires_grad_(True)
    inter = {}
    with true_fp32():
        h = F.conv2d(xr, sd['model.0.weight'], sd['model.0.bias'], stride=2, padding=1)
        inter['y0'] = h
        h = F.leaky_relu(h, 0.2)
        inter['v1'] = h
        for i, (idx, stride) in enumerate(((2, 2), (5, 2), (8, 1))):
            y = F.conv2d(h, sd['model.%d.weight' % idx], sd['model.%d.bias' % idx], stride=stride, padding=1)
            inter['y%d' % (i + 1)] = y
            h = F.leaky_relu(F.instance_norm(y), 0.2)
            inter['v%d' % (i + 2)] = h
        y = F.conv2d(h, sd['model.11.weight'], sd['model.11.bias'], stride=1, padding=1)
        inter['y4'] = y
        for t in inter.values():
            t.retain_grad()
        (y * gout).sum().backward()
    print("out", rel_l2(out, y))
    for i in range(5):
        got = rec[('dy', i)]
        ref = inter['y%d' % i].grad
        c = ref.shape[1]
        print("dy stage", i, rel_l2(got[..., :c].permute(0, 3, 1, 2).float(), ref))
    for v in range(1, 5):
        got = rec[('dfull', v)]
        ref = inter['v%d' % v].grad
        c = ref.shape[1]
        print("d value", v, rel_l2(got[..., :c].permute(0, 3, 1, 2).float(), ref))
    print("gx", rel_l2(x.grad, xr.grad))
    named = dict(net.named_parameters())
    for k in ('model.0.weight', 'model.2.weight', 'model.5.weight', 'model.8.weight', 'model.11.weight',
              'model.0.bias', 'model.11.bias'):
        pass
    ref_params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    sys.path.insert(0, ROOT)
    from oracle import networks_oracle as O
    xr2 = x.detach().clone().requires_grad_(True)
    with true_fp32():
        (O.nlayer_discriminator(ref_params, xr2) * gout).sum().backward()
    for k, v in ref_params.items():
        print(k, rel_l2(named[k].grad, v.grad), float(v.grad.norm()))


if __name__ == "__main__":
    main()
