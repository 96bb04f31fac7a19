import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rel_l2, seeded_image, true_fp32
from oracle import networks5_oracle as O5
from cycle_depth_estimation_b200 import networks5_ds as N, graph, ops
import torch.nn.functional as F
for cin, h, w in ((512, 24, 32), (256, 48, 64), (128, 96, 128), (512, 24, 80)):
    net = N._Discriminator(input_nc=cin)
    sd = O5.synth_state_dict(net.state_dict(), 4); sd['model.1.weight'] = sd['model.10.weight']
    net.load_state_dict(sd); net = net.cuda().train(); sd = {k: v.cuda() for k, v in sd.items()}
    x = seeded_image(2, cin, h, w, seed=70)
    with torch.no_grad():
        out = net(x)
        with true_fp32():
            ref = O5.discriminator({k: v.clone() for k, v in sd.items()}, x)
    print(cin, h, w, tuple(out.shape), "err", rel_l2(out, ref), flush=True)
    # layer by layer
    with torch.no_grad(), true_fp32():
        hs = []
        hcur = F.conv2d(x, sd['model.0.weight'], None, stride=2, padding=1); hs.append(hcur)
        hcur = F.prelu(hcur, sd['model.1.weight'])
        for ci, bi, pi, st in ((2, 3, 4, 2), (5, 6, 7, 2), (8, 9, 10, 1)):
            hcur = F.conv2d(hcur, sd['model.%d.weight' % ci], None, stride=st, padding=1); hs.append(hcur)
            hcur = F.prelu(O5._bn(hcur, {k: v.clone() for k, v in sd.items()}, 'model.%d' % bi), sd['model.%d.weight' % pi])
    # our convs individually on the oracle's inputs
    tape = graph.Tape(True, x.device, False); tape.input_wants = [False]
    v = tape.input_nchw(x)
    c0 = tape.stage(v, net.model[0], None, ops.ACT_NONE)
    o, _ = tape.output_nchw(c0)
    print("   conv0 err", rel_l2(o, hs[0]))
    prev = F.prelu(hs[0], sd['model.1.weight'])
    for idx, (ci, bi) in enumerate(((2, 3), (5, 6), (8, 9))):
        v = tape.input_nchw(prev.contiguous())
        y = tape.stage(v, net.model[ci], net.model[bi], ops.ACT_NONE)
        o, _ = tape.output_nchw(y)
        r = O5._bn(hs[idx + 1], {k: v_.clone() for k, v_ in sd.items()}, 'model.%d' % bi)
        print("   conv%d+bn out %s err %.4f" % (ci, tuple(o.shape), rel_l2(o, r)))
        prev = F.prelu(r, sd['model.%d.weight' % (bi + 1)])
