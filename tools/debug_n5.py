import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rel_l2, seeded_image, true_fp32
from oracle import networks5_oracle as O5
from cycle_depth_estimation_b200 import networks5_ds as N
net = N.General_net()
sd = O5.synth_state_dict(net.state_dict(), 2)
net.load_state_dict(sd); net = net.cuda().train()
sd = {k: v.cuda() for k, v in sd.items()}
x = seeded_image(2, 64, 32, 64, seed=43)
with torch.no_grad():
    head, feats = net(x, 'S')
    with true_fp32():
        rhead, rfeats = O5.general_net({k: v.clone() for k, v in sd.items()}, x, 'S')
print("head", rel_l2(head, rhead))
for i, (f, r) in enumerate(zip(feats, rfeats)):
    print("feat", i, tuple(f.shape), rel_l2(f, r))
    c = f.shape[1]
    step = 32
    errs = [round(rel_l2(f[:, j:j + step], r[:, j:j + step]), 4) for j in range(0, c, step)]
    print("   per-32ch", errs[:6], "...", errs[-6:])
with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
    ahead, afeats = O5.general_net({k: v.clone() for k, v in sd.items()}, x, 'S')
print("autocast head", rel_l2(ahead.float(), rhead), [round(rel_l2(a.float(), r), 4) for a, r in zip(afeats, rfeats)])
for size in ((64, 128), (96, 320)):
    x2 = seeded_image(2, 64, size[0], size[1], seed=47)
    with torch.no_grad():
        head, feats = net(x2, 'S')
        with true_fp32():
            rhead, rfeats = O5.general_net({k: v.clone() for k, v in sd.items()}, x2, 'S')
        with torch.autocast('cuda', dtype=torch.bfloat16):
            ahead, afeats = O5.general_net({k: v.clone() for k, v in sd.items()}, x2, 'S')
    print(size, "ours head", round(rel_l2(head, rhead), 4), [round(rel_l2(a, r), 4) for a, r in zip(feats, rfeats)],
          "autocast head", round(rel_l2(ahead.float(), rhead), 4), [round(rel_l2(a.float(), r), 4) for a, r in zip(afeats, rfeats)])
