"""Generates tests/golden/*.pt by running the REFERENCE's own modules (imported by file path from
/root/reference, which exists only in the build container) on seeded inputs. The fixtures pin the
oracle restatement (oracle/networks_oracle.py) wherever the reference itself cannot travel.

    python oracle/make_golden.py            # rewrites tests/golden/
"""
import contextlib
import importlib.util
import io
import os
import random
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
REF = os.environ.get("CDB_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def load_ref(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def available():
    return os.path.exists(os.path.join(REF, "models", "networks.py"))


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def image(n, c, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand((n, c, h, w), generator=g) * 2 - 1


def make_networks(N):
    fx = {}
    torch.manual_seed(0)
    with quiet():
        g = N.define_G(3, 3, 8, 'resnet_6blocks', 'instance', False, 'normal', 0.02, ['cpu'])
    x = image(2, 3, 32, 32, 1234)
    xg = x.clone().requires_grad_(True)
    out = g(xg)
    (out * image(2, 3, 32, 32, 5)).sum().backward()
    fx['resnet'] = dict(sd=g.state_dict(), x=x, out=out.detach(), gx=xg.grad,
                        gw={k: p.grad for k, p in g.named_parameters()})
    for norm in ('instance', 'batch'):
        torch.manual_seed(1)
        with quiet():
            d = N.define_D(3, 8, 'basic', 3, norm, norm == 'batch', 'normal', 0.02, ['cpu'])
        sd0 = {k: v.clone() for k, v in d.state_dict().items()}
        x = image(2, 3, 64, 64, 77)
        out = d(x)
        fx['nlayer_' + norm] = dict(sd=sd0, x=x, out=out.detach(), sd_after=d.state_dict())
    torch.manual_seed(2)
    with quiet():
        u = N.define_G(3, 3, 4, 'unet_128', 'batch', False, 'normal', 0.02, ['cpu'])
    sd0 = {k: v.clone() for k, v in u.state_dict().items()}
    x = image(2, 3, 128, 128, 21)
    fx['unet'] = dict(sd=sd0, x=x.clone(), out=u(x.clone()).detach())
    torch.manual_seed(3)
    with quiet():
        pd = N.define_D(3, 8, 'pixel', 3, 'instance', False, 'normal', 0.02, ['cpu'])
    x = image(1, 3, 16, 16, 8)
    fx['pixel'] = dict(sd=pd.state_dict(), x=x, out=pd(x).detach())
    pred = image(2, 1, 6, 6, 9)
    fx['gan_loss'] = dict(pred=pred,
                          lsgan_real=float(N.GANLoss(True)(pred, True)), lsgan_fake=float(N.GANLoss(True)(pred, False)),
                          bce_real=float(N.GANLoss(False)(torch.sigmoid(pred), True)),
                          bce_fake=float(N.GANLoss(False)(torch.sigmoid(pred), False)))
    return fx


def make_pool(P):
    """Each image carries its own id, so the ids the reference pool returns are its decision trace."""
    random.seed(1234)
    pool = P.ImagePool(7)
    returned = []
    nxt = 0
    for q in range(64):
        b = 1 + (q % 4)
        batch = torch.stack([torch.full((1, 2, 2), float(nxt + i)) for i in range(b)])
        nxt += b
        out = pool.query(batch)
        returned.append([int(v) for v in out[:, 0, 0, 0]])
    return dict(pool_size=7, seed=1234, returned=returned)


def make_metrics(E):
    rows, pairs = [], []
    for i in range(6):
        rng = np.random.default_rng(2019 + i)
        h, w = (375, 1242) if i < 2 else (31 + 7 * i, 45 + 3 * i)
        gt = rng.integers(0, 80, (h, w), dtype=np.uint8)
        gt[rng.random((h, w)) < 0.3] = 0
        pred = rng.integers(0, 256, (h, w), dtype=np.uint8)
        p = pred / 255 * 80
        p[p < 1] = 1
        p[p > 50] = 50
        mask = np.logical_and(gt > 1, gt < 50)
        with quiet():
            rows.append([float(v) for v in E.compute_errors(gt[mask], p[mask])])
        pairs.append((h, w, 2019 + i))
    return dict(pairs=pairs, rows=rows)


def make_networks5(N5):
    """Outputs of the reference's own G_1 / General_net / R_dep / _Discriminator classes (CPU, training-mode
    BatchNorm) on name-keyed synthetic weights (networks5_oracle.synth_state_dict) at small spatial sizes."""
    import builtins
    from oracle import networks5_oracle as O5
    fx = {}
    real_print = builtins.print
    builtins.print = lambda *a, **k: None       # R_dep.forward prints (networks5_ds.py:791)
    try:
        def load(net, seed):
            sd = O5.synth_state_dict(net.state_dict(), seed)
            if 'model.10.weight' in sd and 'model.1.weight' in sd:   # the shared nn.PReLU of _Discriminator
                sd['model.1.weight'] = sd['model.10.weight']
            net.load_state_dict(sd, strict=True)
            return net.train()
        g1 = load(N5.G_1(), 1)
        x = image(2, 3, 32, 64, 31)
        ss = g1(x)
        fx['g1'] = dict(x=x, out=ss.detach())
        g2 = load(N5.General_net(), 2)
        head_s, feats_s = g2(ss.detach(), 'S')
        xr = image(2, 3, 32, 64, 32)
        head_r, feats_r = g2(xr, 'R')
        fx['g2_S'] = dict(head=head_s.detach(), feat_means=[float(f.mean()) for f in feats_s],
                          feat_abs=[float(f.abs().mean()) for f in feats_s], feat3=feats_s[3].detach())
        fx['g2_R'] = dict(x=xr, head=head_r.detach(), feat_abs=[float(f.abs().mean()) for f in feats_r])
        rd = load(N5.R_dep(), 3)
        feats, seg, (dep4, dep1) = rd(feats_s, head_s.detach())
        fx['rd'] = dict(out0=feats[0].detach(), out1=feats[1].detach(), out2=feats[2].detach()[:, :, ::2, ::2],
                        seg=seg.detach()[:, :, ::2, ::2], dep4=[d.detach() for d in dep4], dep1=dep1.detach())
        fd = load(N5._Discriminator(input_nc=128), 4)
        xd = image(2, 128, 32, 64, 33)
        fx['fd'] = dict(x=xd, out=fd(xd).detach())
        t = image(2, 4, 8, 8, 34)
        t[t > 0.5] = 1.0
        t[t < -0.5] = -1.0
        o_m, z_m = N5.get_masks(t)
        xin = torch.tanh(image(2, 1, 8, 8, 35))
        fx['bcedep'] = dict(x=xin, t=t, o_m=o_m, z_m=z_m, loss=float(N5.BCEDepLoss()(xin, t, o_m, z_m)))
    finally:
        builtins.print = real_print
    return fx


def make_encoder_decoder(ED):
    """Outputs and input gradients of the reference's own _UNetEncoder / _UNetDecoder (models/encoder_decoder.py,
    CPU, training-mode BatchNorm, ngf=8) on name-keyed synthetic weights with the shared PReLU tied."""
    from oracle import encoder_decoder_oracle as OE
    from oracle import networks5_oracle as O5
    enc, dec = ED._UNetEncoder(input_nc=3, ngf=8), ED._UNetDecoder(output_nc=5, ngf=8)
    enc.load_state_dict(OE.tie_prelu(O5.synth_state_dict(enc.state_dict(), 11)), strict=True)
    dec.load_state_dict(OE.tie_prelu(O5.synth_state_dict(dec.state_dict(), 12)), strict=True)
    enc.train()
    dec.train()
    x = image(1, 3, 96, 96, 51).requires_grad_(True)
    feats = enc(x)
    outs = dec(feats)
    gout = image(1, 5, 96, 96, 52)
    (outs[-1] * gout).sum().backward()
    return dict(x=x.detach(), gout=gout, feats=[f.detach() for f in feats], outs=[o.detach() for o in outs[1:]],
                gx=x.grad, enc_keys=list(enc.state_dict().keys()), dec_keys=list(dec.state_dict().keys()),
                g_enc_slope=enc.conv1[3].weight.grad.clone(), g_dec_slope=dec.deconv_center.model[3].weight.grad.clone(),
                g_dec_out1=dec.output1.model[1].weight.grad.clone())


def make_seg_network(SN):
    """Outputs and gradients of the reference's own models/seg_network.py::_UNetGenerator (CPU, training-mode BatchNorm,
    ngf=8) for both heads on name-keyed synthetic weights with the shared PReLU tied."""
    from oracle import encoder_decoder_oracle as OE
    from oracle import networks5_oracle as O5
    net = SN._UNetGenerator(input_nc=3, output_nc=22, ngf=8)
    net.load_state_dict(OE.tie_prelu(O5.synth_state_dict(net.state_dict(), 21)), strict=True)
    net.train()
    fx = dict(keys=list(net.state_dict().keys()), heads={})
    for head, nc, seed in (('syn', 22, 61), ('real', 28, 62)):
        net.zero_grad()
        x = image(1, 3, 96, 96, seed).requires_grad_(True)          # inputs are regenerated from their seeds by the test
        center_in, out1 = net(x, head)
        gout = image(1, nc, 96, 96, seed + 10)
        (out1 * gout).sum().backward()
        fx['heads'][head] = dict(seed=seed, center_in=center_in.detach(), out1=out1.detach()[:, :, ::3, ::3].clone(),
                                 gx=x.grad.clone(), g_slope=net.conv1[3].weight.grad.clone(),
                                 g_out1=getattr(net, 'output1_' + head).model[1].weight.grad.clone(),
                                 g_conv1=net.conv1[1].weight.grad.clone())
    # _MultiscaleDiscriminator(num_D=1) / _Discriminator (models/seg_network.py:561-627), ndf=8
    D = SN._MultiscaleDiscriminator(input_nc=5, ndf=8)
    D.load_state_dict(OE.tie_prelu_prefixed(O5.synth_state_dict(D.state_dict(), 23)), strict=True)
    D.train()
    x = image(2, 5, 64, 64, 63).requires_grad_(True)
    out = D(x)
    assert isinstance(out, list) and len(out) == 1
    gout = image(*out[0].shape, 64)
    (out[0] * gout).sum().backward()
    fx['disc'] = dict(keys=list(D.state_dict().keys()), x=x.detach(), out=out[0].detach(), gout=gout, gx=x.grad.clone(),
                      g_slope=D.scale0.model[1].weight.grad.clone(), g_w0=D.scale0.model[0].weight.grad.clone())
    return fx


def _ref_lines(rel, first, last):
    """Source lines [first, last] (1-based) of a reference file with comment-only lines dropped, dedented."""
    import textwrap
    with open(os.path.join(REF, rel)) as f:
        lines = f.read().split("\n")[first - 1:last]
    lines = [ln for ln in lines if ln.strip() and not ln.strip().startswith("#")]
    return textwrap.dedent("\n".join(lines))


def _ref_dict(rel, name):
    """The literal of ``self.<name> = {...}`` in a reference file, evaluated (ignore_label as the file defines it)."""
    import re
    src = open(os.path.join(REF, rel)).read()
    ign = int(re.search(r"^ignore_label\s*=\s*(\d+)", src, re.M).group(1))
    start = src.index("self.%s = {" % name) + len("self.%s = " % name)
    end = src.index("}", start) + 1
    return eval(src[start:end], {"ignore_label": ign})


def make_input_pipeline():
    """Runs the reference's OWN loader statements (new_multi/try_data.py:199-211 and :240-272,
    datasets/dataset_synthia.py:172-175) on seeded arrays by executing the source lines read from /root/reference."""
    import torchvision.transforms as transforms
    rng = np.random.default_rng(77)
    fx = {'depth': [], 'labels': {}}
    depth_block = _ref_lines("new_multi/try_data.py", 240, 272)
    cases = [rng.uniform(0, 65535, (24, 40)), rng.uniform(0, 9000, (24, 40)), rng.uniform(1500, 3500, (16, 16)),
             np.full((8, 8), 1234.0), rng.uniform(4100, 4900, (8, 12)), rng.uniform(0, 900, (8, 12))]
    for c in cases:
        d = c.astype(np.float32)
        ns = {'np': np, 'depth_source': d.copy()}
        with np.errstate(all='ignore'):
            exec(depth_block, ns)
        fx['depth'].append(dict(depth=d, dep_l=ns['depth_source'], depth_l_s=ns['depth_labels']))
    real_map = _ref_dict("new_multi/try_data.py", "real_id_to_trainid")
    syn_map = _ref_dict("datasets/dataset_synthia.py", "syn_id_to_trainid")
    lab = rng.integers(0, 40, (32, 48), dtype=np.uint8)
    lab[0, :8] = np.arange(248, 256, dtype=np.uint8)

    class _Self:
        real_id_to_trainid = real_map
        syn_id_to_trainid = syn_map
    ns = {'np': np, 'self': _Self, 'lab_source': lab.copy(), 'lab_target': lab.copy()}
    exec(_ref_lines("new_multi/try_data.py", 199, 211), ns)
    fx['labels']['sequential'] = dict(lab=lab, mapping=real_map, zero_to=7, out=ns['lab_source'].astype(np.uint8),
                                      out_target=ns['lab_target'].astype(np.uint8))
    ns = {'np': np, 'self': _Self, 'lab_source': lab.copy()}
    exec(_ref_lines("datasets/dataset_synthia.py", 172, 175), ns)
    fx['labels']['masked'] = dict(lab=lab, mapping=syn_map, out=ns['lab_source_copy'].astype(np.uint8))
    img = rng.integers(0, 256, (20, 28, 3), dtype=np.uint8)
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
    fx['normalize'] = dict(img=img, out=tf(img))
    return fx


def make_pil_resize():
    """Calls Pillow itself (the reference's third-party dependency) with the reference's call-site arguments
    (datasets/dataset_synthia.py:154-167: resize([640, 192], BILINEAR / NEAREST); new_multi/try_data.py:164-167:
    resize([576, 192], BILINEAR); paired_transform's F.hflip = transpose(FLIP_LEFT_RIGHT)) on small seeded images whose
    aspect ratios are those of the datasets (Synthia 1280x760, KITTI 1242x375 scaled down by 1/5), plus an upscaling and
    a one-axis case."""
    import PIL
    from PIL import Image
    rng = np.random.default_rng(2024)
    fx = {'pillow': PIL.__version__, 'cases': []}
    for (h, w, dw, dh) in [(152, 256, 128, 38), (75, 248, 128, 38), (75, 248, 115, 38), (40, 56, 90, 71), (48, 64, 64, 20),
                           (33, 47, 47, 33), (760, 1280, 640, 192)]:
        full = h >= 700      # full Synthia size: the input is regenerated from its seed, the outputs are stored subsampled
        r = np.random.default_rng(9) if full else rng
        img = r.integers(0, 256, (h, w, 3), dtype=np.uint8)
        lab = r.integers(0, 34, (h, w), dtype=np.uint8)
        pim, plab = Image.fromarray(img).convert('RGB'), Image.fromarray(lab)
        bil = np.array(pim.resize([dw, dh], Image.BILINEAR))
        near = np.array(plab.resize([dw, dh], Image.NEAREST))
        flip = np.array(pim.resize([dw, dh], Image.BILINEAR).transpose(Image.FLIP_LEFT_RIGHT))
        if full:
            case = dict(size=(dw, dh), shape=(h, w), seed=9, sub=(7, 11), bilinear=bil[::7, ::11].copy(),
                        nearest=near[::7, ::11].copy(), bilinear_flipped=flip[::7, ::11].copy())
        else:
            case = dict(size=(dw, dh), shape=(h, w), img=img, lab=lab, bilinear=bil, nearest=near, bilinear_flipped=flip)
        fx['cases'].append(case)
    return fx


def main():
    if "--seg-network-only" in sys.argv:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        SN = load_ref("ref_seg_network", "models/seg_network.py")
        torch.save(make_seg_network(SN), os.path.join(OUT, "seg_network.pt"))
        print("seg_network.pt", os.path.getsize(os.path.join(OUT, "seg_network.pt")))
        return
    if "--pil-only" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        torch.save(make_pil_resize(), os.path.join(OUT, "pil_resize.pt"))
        print("pil_resize.pt", os.path.getsize(os.path.join(OUT, "pil_resize.pt")))
        return
    if not available():
        raise SystemExit("reference not found at %s" % REF)
    os.makedirs(OUT, exist_ok=True)
    N = load_ref("ref_networks", "models/networks.py")
    P = load_ref("ref_image_pool", "util/image_pool.py")
    E = load_ref("ref_my_eval", "new_multi/my_eval.py")
    torch.save(make_networks(N), os.path.join(OUT, "networks.pt"))
    torch.save(make_pool(P), os.path.join(OUT, "image_pool.pt"))
    torch.save(make_metrics(E), os.path.join(OUT, "metrics.pt"))
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    N5 = load_ref("ref_networks5_ds", "new_multi/networks5_ds.py")
    torch.save(make_networks5(N5), os.path.join(OUT, "networks5.pt"))
    ED = load_ref("ref_encoder_decoder", "models/encoder_decoder.py")
    torch.save(make_encoder_decoder(ED), os.path.join(OUT, "encoder_decoder.pt"))
    SN = load_ref("ref_seg_network", "models/seg_network.py")
    torch.save(make_seg_network(SN), os.path.join(OUT, "seg_network.pt"))
    torch.save(make_input_pipeline(), os.path.join(OUT, "input_pipeline.pt"))
    torch.save(make_pil_resize(), os.path.join(OUT, "pil_resize.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
