"""ORACLE — test infrastructure only (imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by the product package).

A plain-PyTorch fp32, functional restatement of the arithmetic of the reference's hot path, driven by
a reference-layout ``state_dict``:

* ``resnet_generator``      models/networks.py:145-191 (+ ResnetBlock :195-236)
* ``nlayer_discriminator``  models/networks.py:320-364
* ``pixel_discriminator``   models/networks.py:367-389
* ``unet_generator``        models/networks.py:243-316 (incl. the in-place LeakyReLU skip quirk, SURVEY B-5)
* ``gan_loss`` / ``l1``     models/networks.py:119-138; models/cycle_gan_model.py:63-64
* ``ImagePoolOracle``       util/image_pool.py:5-32
* ``compute_errors`` / ``eval_metric``  new_multi/my_eval.py:7-31 / :35-108
* ``CycleGANStepOracle``    models/cycle_gan_model.py:80-160 (forward, backward_G, the 4x D loop)

Parity pin: the reference has no golden vectors or tests for this path, so the
restatement is pinned against the reference's OWN modules imported from /root/reference in the build
container (tests/test_oracle_pin.py, skipped where the reference is absent) and against fixtures those
modules generated (oracle/make_golden.py -> tests/golden/*.pt).  The convolution / normalisation
arithmetic itself lives in the third-party dependency torch (requirements.txt: torch>=0.4.0, unpinned;
evaluated here with the torch of this image), exactly as in the reference.
"""
import random

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-5

# Optional emulation of the product path's STORAGE rounding (bf16 activations and weights, fp32
# accumulation) with a straight-through gradient. With it the oracle takes the same ReLU / LeakyReLU
# branches as the bf16 engine, which isolates the backward kernels from the activation-sign flips that
# any reduced-precision forward pass has (DESIGN.md "tolerances").
_EMULATE_BF16 = [False]


class emulate_bf16_storage:
    def __enter__(self):
        self.prev = _EMULATE_BF16[0]
        _EMULATE_BF16[0] = True

    def __exit__(self, *a):
        _EMULATE_BF16[0] = self.prev


def _q(t):
    if not _EMULATE_BF16[0] or t is None:
        return t
    return t + (t.to(torch.bfloat16).float() - t).detach()


# ------------------------------------------------------------------------------------------------
# networks
# ------------------------------------------------------------------------------------------------
def _norm(x, sd, key, norm, training=True, momentum=0.1):
    """InstanceNorm2d(affine=False) or BatchNorm2d(affine=True) as get_norm_layer builds them
    (models/networks.py:12-22)."""
    if norm == 'instance':
        return F.instance_norm(x, eps=EPS)
    if norm == 'batch':
        return F.batch_norm(x, sd.get(key + '.running_mean'), sd.get(key + '.running_var'), sd[key + '.weight'],
                            sd[key + '.bias'], training, momentum, EPS)
    raise NotImplementedError(norm)


def resnet_generator(sd, x, n_blocks=9, norm='instance', training=True, prefix='model.'):
    """models/networks.py:157-191. Layer indices follow the reference's nn.Sequential."""
    g = lambda k: _q(sd.get(prefix + k)) if k.endswith('weight') else sd.get(prefix + k)
    h = _q(F.conv2d(F.pad(_q(x), (3, 3, 3, 3), mode='reflect'), g('1.weight'), g('1.bias')))
    h = _q(F.relu(_norm(h, sd, prefix + '2', norm, training)))
    for idx in (4, 7):
        h = _q(F.conv2d(h, g('%d.weight' % idx), g('%d.bias' % idx), stride=2, padding=1))
        h = _q(F.relu(_norm(h, sd, prefix + '%d' % (idx + 1), norm, training)))
    idx = 10
    for _ in range(n_blocks):  # ResnetBlock: x + conv_block(x), models/networks.py:234-236
        b = '%d.conv_block.' % idx
        r = _q(F.conv2d(F.pad(h, (1, 1, 1, 1), mode='reflect'), g(b + '1.weight'), g(b + '1.bias')))
        r = _q(F.relu(_norm(r, sd, prefix + b + '2', norm, training)))
        r = _q(F.conv2d(F.pad(r, (1, 1, 1, 1), mode='reflect'), g(b + '5.weight'), g(b + '5.bias')))
        r = _norm(r, sd, prefix + b + '6', norm, training)
        h = _q(h + r)
        idx += 1
    for _ in range(2):
        h = _q(F.conv_transpose2d(h, g('%d.weight' % idx), g('%d.bias' % idx), stride=2, padding=1,
                                  output_padding=1))
        h = _q(F.relu(_norm(h, sd, prefix + '%d' % (idx + 1), norm, training)))
        idx += 3
    idx += 1  # ReflectionPad2d(3)
    h = F.conv2d(F.pad(h, (3, 3, 3, 3), mode='reflect'), g('%d.weight' % idx), g('%d.bias' % idx))
    return torch.tanh(h)


def nlayer_discriminator(sd, x, norm='instance', use_sigmoid=False, training=True, prefix='model.'):
    """models/networks.py:330-364 with n_layers=3 (define_D hard-wires it, :100)."""
    g = lambda k: _q(sd.get(prefix + k)) if k.endswith('weight') else sd.get(prefix + k)
    h = _q(F.leaky_relu(F.conv2d(_q(x), g('0.weight'), g('0.bias'), stride=2, padding=1), 0.2))
    for idx, stride in ((2, 2), (5, 2), (8, 1)):
        h = _q(F.conv2d(h, g('%d.weight' % idx), g('%d.bias' % idx), stride=stride, padding=1))
        h = _q(F.leaky_relu(_norm(h, sd, prefix + '%d' % (idx + 1), norm, training), 0.2))
    h = F.conv2d(h, g('11.weight'), g('11.bias'), stride=1, padding=1)
    return torch.sigmoid(h) if use_sigmoid else h


def pixel_discriminator(sd, x, norm='instance', use_sigmoid=False, training=True, prefix='net.'):
    """models/networks.py:375-389."""
    g = lambda k: sd.get(prefix + k)
    h = F.leaky_relu(F.conv2d(x, g('0.weight'), g('0.bias')), 0.2)
    h = F.leaky_relu(_norm(F.conv2d(h, g('2.weight'), g('2.bias')), sd, prefix + '3', norm, training), 0.2)
    h = F.conv2d(h, g('5.weight'), g('5.bias'))
    return torch.sigmoid(h) if use_sigmoid else h


def unet_generator(sd, x, num_downs=8, norm='batch', training=True, dropout_masks=None):
    """models/networks.py:243-316. The reference's in-place LeakyReLU(0.2, True) mutates a block's
    input before ``torch.cat([x, model(x)], 1)`` materialises, so the skip carries leaky_relu(x)
    (SURVEY B-5); dropout (three middle blocks when enabled) is applied with the supplied masks
    (already scaled by 1/(1-p)) or skipped when dropout_masks is None."""
    def block(prefix, h, depth):
        g = lambda k: sd.get(prefix + k)
        outermost, innermost = depth == 0, depth == num_downs - 1
        if outermost:
            d = F.conv2d(h, g('0.weight'), g('0.bias'), stride=2, padding=1)
            u = block(prefix + '1.model.', d, depth + 1)
            u = F.conv_transpose2d(F.relu(u), g('3.weight'), g('3.bias'), stride=2, padding=1)
            return torch.tanh(u)
        a = F.leaky_relu(h, 0.2)  # also what the skip connection carries
        d = F.conv2d(a, g('1.weight'), g('1.bias'), stride=2, padding=1)
        if innermost:
            u = F.conv_transpose2d(F.relu(d), g('3.weight'), g('3.bias'), stride=2, padding=1)
            u = _norm(u, sd, prefix + '4', norm, training)
        else:
            d = _norm(d, sd, prefix + '2', norm, training)
            u = block(prefix + '3.model.', d, depth + 1)
            u = F.conv_transpose2d(F.relu(u), g('5.weight'), g('5.bias'), stride=2, padding=1)
            u = _norm(u, sd, prefix + '6', norm, training)
            if dropout_masks is not None and depth in dropout_masks:
                u = u * dropout_masks[depth]
        return torch.cat([a, u], 1)

    return block('model.model.', x, 0)


def gan_loss(pred, target_is_real, use_lsgan=True, real_label=1.0, fake_label=0.0):
    """GANLoss.__call__ (models/networks.py:129-138): MSE / BCE against the expanded label buffer."""
    t = torch.full_like(pred, real_label if target_is_real else fake_label)
    return F.mse_loss(pred, t) if use_lsgan else F.binary_cross_entropy(pred, t)


def l1(a, b):
    return F.l1_loss(a, b)


# ------------------------------------------------------------------------------------------------
# ImagePool (util/image_pool.py:5-32)
# ------------------------------------------------------------------------------------------------
class ImagePoolOracle:
    """History buffer: fill until pool_size, then with probability 1/2 swap the incoming image with a
    random stored one. Draws from Python's global ``random`` exactly like the reference
    (uniform(0,1) then randint(0, pool_size-1), inclusive). ``trace`` records the decisions."""

    def __init__(self, pool_size):
        self.pool_size = pool_size
        self.num_imgs = 0
        self.images = []
        self.trace = []

    def query(self, images):
        if self.pool_size == 0:
            return images
        out = []
        for image in images:
            image = torch.unsqueeze(image.detach(), 0)
            if self.num_imgs < self.pool_size:
                self.num_imgs += 1
                self.images.append(image)
                out.append(image)
                self.trace.append(('fill', self.num_imgs - 1))
            elif random.uniform(0, 1) > 0.5:
                slot = random.randint(0, self.pool_size - 1)
                out.append(self.images[slot].clone())
                self.images[slot] = image
                self.trace.append(('swap', slot))
            else:
                out.append(image)
                self.trace.append(('pass', -1))
        return torch.cat(out, 0)


# ------------------------------------------------------------------------------------------------
# depth metrics (new_multi/my_eval.py)
# ------------------------------------------------------------------------------------------------
def compute_errors(ground_truth, predication):
    """new_multi/my_eval.py:7-31 on already-masked 1-D arrays (gt uint8, pred float64). Note
    np.log(uint8) is evaluated in float16 by numpy (SURVEY B-1) — kept, it is the reference result."""
    predication = (predication - predication.min()) / (predication.max() - predication.min()) * 49 + 1
    ratio = np.maximum(ground_truth / predication, predication / ground_truth)
    a1, a2, a3 = [(ratio < 1.25 ** k).mean() for k in (1, 2, 3)]
    rmse = np.sqrt(((ground_truth - predication) ** 2).mean())
    rmse_log = np.sqrt(((np.log(ground_truth) - np.log(predication)) ** 2).mean())
    abs_rel = np.mean(np.abs(ground_truth - predication) / ground_truth)
    sq_rel = np.mean(((ground_truth - predication) ** 2) / ground_truth)
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3


def eval_metric_arrays(gts, preds):
    """new_multi/my_eval.py:35-108 with the PNG reads replaced by in-memory uint8 arrays of equal size
    (so the cv2.resize at :55 is the identity). Returns (7 float32 means, per-image float32 [n,7])."""
    n = len(gts)
    per = np.zeros((max(n, 1000), 7), np.float32)  # the reference accumulates in float32 arrays of 1000
    for i in range(n):
        gt = gts[i]
        pred = preds[i] / 255 * 80                   # :56
        pred[pred < 1] = 1                           # :78
        pred[pred > 50] = 50                         # :79
        mask = np.logical_and(gt > 1, gt < 50)       # :84
        per[i] = compute_errors(gt[mask], pred[mask])  # :100 (stored as float32)
    means = tuple(per[:, k].sum() / n for k in range(7))  # :108
    return means, per[:n]


# ------------------------------------------------------------------------------------------------
# CycleGAN training step (models/cycle_gan_model.py)
# ------------------------------------------------------------------------------------------------
class CycleGANStepOracle:
    """Restated glue of CycleGANModel (models/cycle_gan_model.py:46-160) around functional networks.
    Parameters are leaf tensors in dicts with the reference's state_dict keys; Adam(lr, betas=(beta1,
    0.999)) as at :66-69; D is updated 4 times per G update (:151), the pools are queried in every D
    iteration (:101-109), the loss_D terms are (real + fake) * 0.5 (:97)."""

    def __init__(self, sd_G_A, sd_G_B, sd_D_A, sd_D_B, lr=2e-4, beta1=0.5, lambda_A=10.0, lambda_B=10.0,
                 lambda_idt=0.5, pool_size=50, n_blocks=9, d_iters=4):
        mk = lambda sd: {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        self.G_A, self.G_B, self.D_A, self.D_B = mk(sd_G_A), mk(sd_G_B), mk(sd_D_A), mk(sd_D_B)
        self.lambda_A, self.lambda_B, self.lambda_idt = lambda_A, lambda_B, lambda_idt
        self.n_blocks, self.d_iters = n_blocks, d_iters
        self.fake_A_pool, self.fake_B_pool = ImagePoolOracle(pool_size), ImagePoolOracle(pool_size)
        params = lambda *ds: [p for d in ds for p in d.values() if p.requires_grad]
        self.opt_G = torch.optim.Adam(params(self.G_A, self.G_B), lr=lr, betas=(beta1, 0.999))
        self.opt_D = torch.optim.Adam(params(self.D_A, self.D_B), lr=lr, betas=(beta1, 0.999))
        self.losses = {}

    def _g(self, sd, x):
        return resnet_generator(sd, x, self.n_blocks, 'instance')

    def _d(self, sd, x):
        return nlayer_discriminator(sd, x, 'instance')

    def _set_requires_grad(self, sds, flag):
        for sd in sds:
            for p in sd.values():
                if p.is_floating_point():
                    p.requires_grad_(flag)

    def _d_loss(self, sd, real, fake):  # backward_D_basic :87-99
        return (gan_loss(self._d(sd, real), True) + gan_loss(self._d(sd, fake), False)) * 0.5

    def step(self, real_A, real_B, train=True, apply_updates=True):
        # forward :80-85
        fake_B = self._g(self.G_A, real_A)
        rec_A = self._g(self.G_B, fake_B)
        fake_A = self._g(self.G_B, real_B)
        rec_B = self._g(self.G_A, fake_A)
        # backward_G :111-137
        self._set_requires_grad([self.D_A, self.D_B], False)
        self.opt_G.zero_grad()
        L = self.losses
        if self.lambda_idt > 0:
            L['idt_A'] = l1(self._g(self.G_A, real_B), real_B) * self.lambda_B * self.lambda_idt
            L['idt_B'] = l1(self._g(self.G_B, real_A), real_A) * self.lambda_A * self.lambda_idt
        else:
            L['idt_A'] = L['idt_B'] = 0
        L['G_A'] = gan_loss(self._d(self.D_A, fake_B), True)
        L['G_B'] = gan_loss(self._d(self.D_B, fake_A), True)
        L['cycle_A'] = l1(rec_A, real_A) * self.lambda_A
        L['cycle_B'] = l1(rec_B, real_B) * self.lambda_B
        loss_G = L['G_A'] + L['G_B'] + L['cycle_A'] + L['cycle_B'] + L['idt_A'] + L['idt_B']
        L['G'] = loss_G
        if train:
            loss_G.backward()
            if apply_updates:
                self.opt_G.step()
        # D_A and D_B, four times :151-160
        for _ in range(self.d_iters):
            self._set_requires_grad([self.D_A, self.D_B], True)
            self.opt_D.zero_grad()
            L['D_A'] = self._d_loss(self.D_A, real_B, self.fake_B_pool.query(fake_B))
            L['D_B'] = self._d_loss(self.D_B, real_A, self.fake_A_pool.query(fake_A))
            if train:
                L['D_A'].backward()
                L['D_B'].backward()
                if apply_updates:
                    self.opt_D.step()
        self.fake_A, self.fake_B, self.rec_A, self.rec_B = fake_A, fake_B, rec_A, rec_B
        return {k: float(v) for k, v in L.items()}


# ------------------------------------------------------------------------------------------------
# pix2pix training step (models/pix2pix_model.py)
# ------------------------------------------------------------------------------------------------
class Pix2PixStepOracle:
    """Restated glue of Pix2PixModel (models/pix2pix_model.py:24-111): U-Net generator, PatchGAN on
    cat(A, B) with BatchNorm + sigmoid, BCE GAN loss + lambda_L1 * L1, Adam; D first (:101-105), then G
    (:107-111). pool_size is 0 in the reference defaults (:16), so the pool returns its input."""

    def __init__(self, sd_G, sd_D, num_downs=8, lr=2e-4, beta1=0.5, lambda_L1=100.0, use_lsgan=False):
        is_param = lambda k, v: v.is_floating_point() and 'running_' not in k
        mk = lambda sd: {k: v.detach().clone().requires_grad_(is_param(k, v)) for k, v in sd.items()}
        self.G, self.D = mk(sd_G), mk(sd_D)
        self.num_downs, self.lambda_L1, self.use_lsgan = num_downs, lambda_L1, use_lsgan
        params = lambda d: [p for p in d.values() if p.requires_grad]
        self.D_params = params(self.D)
        self.opt_G = torch.optim.Adam(params(self.G), lr=lr, betas=(beta1, 0.999))
        self.opt_D = torch.optim.Adam(params(self.D), lr=lr, betas=(beta1, 0.999))
        self.losses = {}

    def _d(self, x):
        return nlayer_discriminator(self.D, x, 'batch', use_sigmoid=not self.use_lsgan)

    def step(self, real_A, real_B, apply_updates=True, dropout_masks=None):
        L = self.losses
        fake_B = unet_generator(self.G, real_A, self.num_downs, 'batch', True, dropout_masks)
        for p in self.D_params:
            p.requires_grad_(True)
        self.opt_D.zero_grad()
        L['D_fake'] = gan_loss(self._d(torch.cat((real_A, fake_B), 1).detach()), False, self.use_lsgan)
        L['D_real'] = gan_loss(self._d(torch.cat((real_A, real_B), 1)), True, self.use_lsgan)
        ((L['D_fake'] + L['D_real']) * 0.5).backward()
        if apply_updates:
            self.opt_D.step()
        for p in self.D_params:
            p.requires_grad_(False)
        self.opt_G.zero_grad()
        L['G_GAN'] = gan_loss(self._d(torch.cat((real_A, fake_B), 1)), True, self.use_lsgan)
        L['G_L1'] = l1(fake_B, real_B) * self.lambda_L1
        (L['G_GAN'] + L['G_L1']).backward()
        if apply_updates:
            self.opt_G.step()
        self.fake_B = fake_B
        return {k: float(v.detach()) for k, v in L.items()}


# ------------------------------------------------------------------------------------------------
# validation path of new_multi/train5.py:85-110 (PNG round trip restated in memory)
# ------------------------------------------------------------------------------------------------
def prediction_png_u8(dep):
    """What ends up in the prediction PNG for one depth map ``dep`` (float array [H,W], network convention
    [-1,1]): util/util.py:59-65 (tensor2im: float32 arithmetic, astype(uint8)), new_multi/train5.py:100
    (``img / img.max()``), :110 (``* 255`` and cv2.imwrite of the float64 array: saturate_cast<uchar> = round
    half to even)."""
    image_numpy = np.asarray(dep, dtype=np.float32)
    image_numpy = (image_numpy + 1) / 2.0 * 255.0
    with np.errstate(invalid='ignore'):
        img = image_numpy.astype(np.uint8)
    arr = img / img.max() * 255
    return np.clip(np.rint(arr), 0, 255).astype(np.uint8)


def eval_metric_from_predictions(deps, gts):
    """new_multi/train5.py:97-110 + my_eval.py:52-108 for in-memory inputs: PNG quantisation, cv2.resize to the
    ground-truth size (my_eval.py:55, OpenCV INTER_LINEAR), then the metrics."""
    import cv2
    preds = []
    for dep, gt in zip(deps, gts):
        p8 = prediction_png_u8(dep)
        preds.append(cv2.resize(p8, (gt.shape[1], gt.shape[0])))
    return eval_metric_arrays(list(gts), preds)
