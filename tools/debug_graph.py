import argparse, os, random, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import quiet, seeded_image
from cycle_depth_estimation_b200 import cycle_gan_model as M

def run(variant):
    opt = argparse.Namespace(input_nc=3, output_nc=3, ngf=64, ndf=64, netG='resnet_6blocks', netD='basic',
                             n_layers_D=3, norm='instance', no_dropout=True, init_type='normal', init_gain=0.02,
                             no_lsgan=False, pool_size=5, lr=2e-4, beta1=0.5, lambda_A=10.0, lambda_B=10.0,
                             lambda_identity=0.5, isTrain=True, device='cuda', direction='AtoB', cuda_graph=(variant == 'graph'))
    torch.manual_seed(0); random.seed(77)
    model = M.CycleGANModel()
    with quiet():
        model.initialize(opt)
    if variant == 'devstep':
        for o in (model.optimizer_G, model.optimizer_D):
            o.device_step = True
    hist = []
    for step in range(4):
        a, b = seeded_image(2, 3, 64, 64, seed=100 + step), seeded_image(2, 3, 64, 64, seed=200 + step)
        model.set_input({'img_source': a, 'img_target': b})
        if variant == 'planned':
            model._draw_pool_plans(); model._plan_dev.copy_(model._plan_host)
            model._eager_step(True)
        elif variant == 'sidestream':
            s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                model._eager_step(True)
            torch.cuda.current_stream().wait_stream(s)
        else:
            model.optimize_parameters('train')
        l = model.get_current_losses()
        hist.append((round(l['G_A'], 4), round(l['D_A'], 4), round(l['cycle_A'], 4)))
    return hist

for v in ('eager', 'eager', 'devstep', 'planned', 'sidestream', 'graph'):
    print(v, run(v), flush=True)
