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


def main():
    if not available():
        raise SystemExit("reference not found at %s" % REF)
    os.makedirs(OUT, exist_ok=True)
    N = load_ref("ref_networks", "models/networks.py")
    P = load_ref("ref_image_pool", "util/image_pool.py")
    E = load_ref("ref_my_eval", "new_multi/my_eval.py")
    torch.save(make_networks(N), os.path.join(OUT, "networks.pt"))
    torch.save(make_pool(P), os.path.join(OUT, "image_pool.pt"))
    torch.save(make_metrics(E), os.path.join(OUT, "metrics.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
