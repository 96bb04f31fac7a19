"""The pix2pix training step of the reference (``models/pix2pix_model.py:24-111``) on the B200 networks:
U-Net generator (BatchNorm, dropout), 70x70 PatchGAN on ``cat(A, B)`` (6 channels, BatchNorm, sigmoid),
vanilla GAN loss (BCE) + ``lambda_L1`` * L1, Adam(lr, (beta1, 0.999)); D is updated first, then G (:99-111).

``Pix2PixModel`` mirrors the reference class (``initialize(opt)``, ``set_input``, ``forward``,
``backward_D``, ``backward_G``, ``optimize_parameters()``); like ``cycle_gan_model.CycleGANModel`` it does
not inherit the reference's BaseModel (broken as shipped, SURVEY B-12).  Under ``torch.distributed`` the
gradients are averaged with the all-reduce of ``cycle_gan_model.GradBuckets`` and every BatchNorm layer normalises over
the shards of ALL ranks (``ops.bn_world``: per-layer all-reduce of the sums forward and backward), so a
sharded step equals the single-device step on the global batch (``tools/dp_bn_parity.py``; CDB_BN_SYNC=0 keeps the
per-replica statistics of the reference's own nn.DataParallel).
"""
from collections import OrderedDict

import torch

from . import losses, networks
from .cycle_gan_model import FusedAdam, GradBuckets
from .image_pool import ImagePool


class Pix2PixModel:
    def name(self):
        return 'Pix2PixModel'

    def initialize(self, opt):
        """opt: namespace with input_nc, output_nc, ngf, ndf, netG, netD, n_layers_D, norm, no_dropout,
        init_type, init_gain, no_lsgan, pool_size, lr, beta1, lambda_L1, isTrain, direction (+ device)."""
        self.opt = opt
        self.isTrain = opt.isTrain
        self.device = torch.device(getattr(opt, 'device', 'cuda'))
        gpu_ids = [self.device]
        self.loss_names = ['G_GAN', 'G_L1', 'D_real', 'D_fake']
        self.model_names = ['G', 'D'] if self.isTrain else ['G']
        self.netG = networks.define_G(opt.input_nc, opt.output_nc, opt.ngf, opt.netG, opt.norm, not opt.no_dropout,
                                      opt.init_type, opt.init_gain, gpu_ids)
        if self.isTrain:
            use_sigmoid = opt.no_lsgan
            self.netD = networks.define_D(opt.input_nc + opt.output_nc, opt.ndf, opt.netD, opt.n_layers_D, opt.norm,
                                          use_sigmoid, opt.init_type, opt.init_gain, gpu_ids)
            self.fake_AB_pool = ImagePool(opt.pool_size)
            self.criterionGAN = networks.GANLoss(use_lsgan=not opt.no_lsgan).to(self.device)
            self.criterionL1 = losses.L1Loss()
            adam = FusedAdam if getattr(opt, 'fused_adam', True) else torch.optim.Adam
            self.optimizer_G = adam(self.netG.parameters(), lr=opt.lr, betas=(opt.beta1, 0.999))
            self.optimizer_D = adam(self.netD.parameters(), lr=opt.lr, betas=(opt.beta1, 0.999))
            self.optimizers = [self.optimizer_G, self.optimizer_D]
            self._buckets_G = GradBuckets(self.netG.parameters())
            self._buckets_D = GradBuckets(self.netD.parameters())

    def set_input(self, input):
        AtoB = getattr(self.opt, 'direction', 'AtoB') == 'AtoB'
        self.real_A = input['A' if AtoB else 'B'].to(self.device, non_blocking=True)
        self.real_B = input['B' if AtoB else 'A'].to(self.device, non_blocking=True)
        self.image_paths = input.get('A_paths' if AtoB else 'B_paths')

    def set_requires_grad(self, nets, requires_grad=False):
        if not isinstance(nets, list):
            nets = [nets]
        for net in nets:
            if net is not None:
                for param in net.parameters():
                    param.requires_grad = requires_grad

    def forward(self):
        self.fake_B = self.netG(self.real_A)

    def backward_D(self):
        fake_AB = self.fake_AB_pool.query(torch.cat((self.real_A, self.fake_B), 1))
        pred_fake = self.netD(fake_AB.detach())
        self.loss_D_fake = self.criterionGAN(pred_fake, False)
        real_AB = torch.cat((self.real_A, self.real_B), 1)
        pred_real = self.netD(real_AB)
        self.loss_D_real = self.criterionGAN(pred_real, True)
        self.loss_D = (self.loss_D_fake + self.loss_D_real) * 0.5
        self.loss_D.backward()

    def backward_G(self):
        fake_AB = torch.cat((self.real_A, self.fake_B), 1)
        pred_fake = self.netD(fake_AB)
        self.loss_G_GAN = self.criterionGAN(pred_fake, True)
        self.loss_G_L1 = self.criterionL1(self.fake_B, self.real_B) * self.opt.lambda_L1
        self.loss_G = self.loss_G_GAN + self.loss_G_L1
        self.loss_G.backward()

    def optimize_parameters(self):
        self.forward()
        self.set_requires_grad(self.netD, True)
        self.optimizer_D.zero_grad()
        self.backward_D()
        self._buckets_D.all_reduce()
        self.optimizer_D.step()
        self.set_requires_grad(self.netD, False)
        self.optimizer_G.zero_grad()
        self.backward_G()
        self._buckets_G.all_reduce()
        self.optimizer_G.step()

    def get_current_losses(self):
        out = OrderedDict()
        for name in self.loss_names:
            v = getattr(self, 'loss_' + name)
            out[name] = float(v.detach()) if torch.is_tensor(v) else float(v)
        return out
