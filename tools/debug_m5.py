import os, sys, torch, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rel_l2, quiet, true_fp32
from oracle import networks5_oracle as O5
from test_model5_step_gpu import _inputs
from cycle_depth_estimation_b200.model5 import Seg_Depth
torch.manual_seed(0)
model = Seg_Depth()
with quiet():
    model.initialize(argparse.Namespace(lr=2e-4, beta1=0.5, pool_size=50))
strip = O5.strip_module_prefix
init_sds = [{k: v.clone() for k, v in strip(getattr(model, 'net_' + n).state_dict()).items()} for n in ('G_1', 'G_2', 'R_D', 'FD1', 'FD2', 'FD3')]
oracle = O5.SegDepthStepOracle(*init_sds)
envo = O5.SegDepthStepOracle(*init_sds)
data = _inputs(2, 192, 256, 90)
cu = {k: v.cuda() for k, v in data.items()}
model.set_input(data, 'train'); model.optimize_parameters('train')
got = model.get_current_losses()
with true_fp32():
    ref = oracle.step(cu['img_syn'], cu['img_real'], cu['seg_l_syn'].squeeze(1), cu['seg_l_real'].squeeze(1), cu['dep_l_syn'].squeeze(1), cu['depth_l_s'])
print({k: (round(got[k], 4), round(ref[k], 4)) for k in got})
for i in range(3):
    print("real_feats", i, rel_l2(model.real_feats[i], oracle.real_feats[i]), "syn_feats", rel_l2(model.syn_feats[i], oracle.syn_feats[i]))
print("real head", rel_l2(model.real_features1, oracle_head) if False else "")
with torch.no_grad(), true_fp32():
    for i, n in enumerate(('FD1', 'FD2', 'FD3')):
        net = getattr(model, 'net_' + n)
        sd = strip(net.state_dict())
        for name, f in (("real", oracle.real_feats[i]), ("syn", oracle.syn_feats[i])):
            a = net(f); b = O5.discriminator({k: v.clone() for k, v in sd.items()}, f)
            print(n, name, "D on oracle feats: ours", a.flatten()[:4].tolist(), "ref", b.flatten()[:4].tolist())
with torch.autocast('cuda', dtype=torch.bfloat16):
    e = envo.step(cu['img_syn'], cu['img_real'], cu['seg_l_syn'].squeeze(1), cu['seg_l_real'].squeeze(1), cu['dep_l_syn'].squeeze(1), cu['depth_l_s'])
for i in range(3):
    print("AUTOCAST real_feats", i, rel_l2(envo.real_feats[i].float(), oracle.real_feats[i]), "syn_feats", rel_l2(envo.syn_feats[i].float(), oracle.syn_feats[i]))
