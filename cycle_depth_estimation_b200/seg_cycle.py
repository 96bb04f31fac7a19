"""The SegCycle training step of the reference (``models/seg_cycle.py:26-180``, the model ``train.py`` creates by
default — SURVEY 8(f) row f3) on the B200 networks: the CycleGAN of ``cycle_gan_model.py`` plus two U-Net task
networks (``encoder_decoder._UNetEncoder`` / ``_UNetDecoder``) trained with CrossEntropy(ignore 255) on the real and
the translated images of both domains (:102-108, :129-136).

Differences from ``CycleGANModel`` that follow the reference: ONE discriminator update per generator update
(:167-176, not the four of ``cycle_gan_model.py:151``); the generator optimizer also owns the four task networks
(:69-73); ``set_input`` reads ``lab_source`` / ``lab_target`` (:82-83); decoder A predicts 22 classes, decoder B 28
(:50-51).  Like ``CycleGANModel`` it does not inherit the reference's BaseModel (SURVEY B-12).
"""
import itertools

import torch

from . import losses
from .cycle_gan_model import CycleGANModel, FusedAdam, GradBuckets
from .encoder_decoder import _UNetDecoder, _UNetEncoder
from .graph_step import StepGraph


class SegCycle(CycleGANModel):
    D_ITERS = 1

    def initialize(self, opt):
        """opt: the CycleGAN fields (see CycleGANModel.initialize) plus optional ``seg_ngf`` (64), ``seg_classes_A``
        (22) and ``seg_classes_B`` (28); the reference hard-codes those three (:48-51)."""
        if not opt.isTrain:
            raise NotImplementedError("SegCycle is a training-time model (models/seg_cycle.py:38-41 loads only the "
                                      "generators at test time: use CycleGANModel)")
        # data parallelism: the CycleGAN part shards as in CycleGANModel; the task networks use BatchNorm, whose statistics
        # the tape all-reduces per layer (ops.bn_world), so a sharded step equals the single-process one
        # (tools/dp_bn_parity.py with DP_MODEL=segcycle)
        self._seg_ngf = int(getattr(opt, 'seg_ngf', 64))
        nc_a, nc_b = int(getattr(opt, 'seg_classes_A', 22)), int(getattr(opt, 'seg_classes_B', 28))
        dev = torch.device(getattr(opt, 'device', 'cuda'))
        self.net_encoderA = _UNetEncoder(input_nc=3, ngf=self._seg_ngf).to(dev)
        self.net_encoderB = _UNetEncoder(input_nc=3, ngf=self._seg_ngf).to(dev)
        self.net_decoderA = _UNetDecoder(output_nc=nc_a, ngf=self._seg_ngf).to(dev)
        self.net_decoderB = _UNetDecoder(output_nc=nc_b, ngf=self._seg_ngf).to(dev)
        self.criterionSeg = losses.CrossEntropyLoss(size_average=True, ignore_index=255)
        super().initialize(opt)
        self.loss_names = ['D_A', 'G_A', 'cycle_A', 'idt_A', 'D_B', 'G_B', 'cycle_B', 'idt_B', 'segAreal', 'segBreal',
                           'segAfake', 'segBfake']
        self.model_names = ['G_A', 'G_B', 'D_A', 'D_B', 'encoderA', 'encoderB', 'decoderA', 'decoderB']

    def _task_nets(self):
        return [self.net_encoderA, self.net_encoderB, self.net_decoderA, self.net_decoderB]

    def _g_params(self):
        return itertools.chain(self.netG_A.parameters(), self.netG_B.parameters(),
                               *[n.parameters() for n in self._task_nets()])

    def build_optimizers(self):
        opt = self.opt
        self._graph_mode = bool(getattr(opt, 'cuda_graph', False))
        if getattr(opt, 'fused_adam', True):
            adam = lambda ps, **kw: FusedAdam(ps, device_step=self._graph_mode, **kw)
        else:
            adam = torch.optim.Adam
        self.optimizer_G = adam(self._g_params(), lr=opt.lr, betas=(opt.beta1, 0.999))
        self.optimizer_D = adam(itertools.chain(self.netD_A.parameters(), self.netD_B.parameters()),
                                lr=opt.lr, betas=(opt.beta1, 0.999))
        self._step_graph, self._plan_host, self._plan_dev, self._plan_slot = StepGraph(), None, None, 0
        self.optimizers = [self.optimizer_G, self.optimizer_D]
        self._buckets_G = GradBuckets(self._g_params())
        self._buckets_D = GradBuckets(itertools.chain(self.netD_A.parameters(), self.netD_B.parameters()))

    def set_input(self, input):
        super().set_input(input)
        la, lb = input['lab_source'], input['lab_target']
        if getattr(self, '_graph_mode', False):
            if getattr(self, 'lab_A', None) is None or self.lab_A.shape != la.shape:
                self.lab_A = torch.empty(tuple(la.shape), dtype=torch.int64, device=self.device)
                self.lab_B = torch.empty(tuple(lb.shape), dtype=torch.int64, device=self.device)
            self.lab_A.copy_(la, non_blocking=True)
            self.lab_B.copy_(lb, non_blocking=True)
            return
        self.lab_A = la.to(self.device, non_blocking=True).long()
        self.lab_B = lb.to(self.device, non_blocking=True).long()

    def Seg_basic(self, encoder, decoder, input, gt):
        """models/seg_cycle.py:102-108: CrossEntropy on the full-resolution output of the decoder."""
        output = decoder(encoder(input))
        return self.criterionSeg(output[-1], gt.squeeze(1)), output

    def backward_G(self):
        loss_cycle_gan = super().backward_G()
        self.loss_segAreal, self.segAreal = self.Seg_basic(self.net_encoderA, self.net_decoderA, self.real_A, self.lab_A)
        self.loss_segAfake, self.segAfake = self.Seg_basic(self.net_encoderB, self.net_decoderA, self.fake_B, self.lab_A)
        self.loss_segBreal, self.segBreal = self.Seg_basic(self.net_encoderB, self.net_decoderB, self.real_B, self.lab_B)
        self.loss_segBfake, self.segBfake = self.Seg_basic(self.net_encoderA, self.net_decoderB, self.fake_A, self.lab_B)
        self.loss_G = (loss_cycle_gan + self.loss_segAfake + self.loss_segAreal + self.loss_segBfake
                       + self.loss_segBreal)
        return self.loss_G

    def _eager_step(self, train):
        gen_side = [self.netG_B, self.netG_A] + self._task_nets()
        self.forward()
        self.set_requires_grad([self.netD_A, self.netD_B], False)
        self.set_requires_grad(gen_side, True)
        self.optimizer_G.zero_grad()
        self.loss_G = self.backward_G()
        if train:
            self.loss_G.backward()
            self._buckets_G.all_reduce()
            self.optimizer_G.step()
        self.set_requires_grad([self.netD_A, self.netD_B], True)
        self.set_requires_grad(gen_side, False)
        self.optimizer_D.zero_grad()
        self.loss_D_A = self.backward_D_A()
        self.loss_D_B = self.backward_D_B()
        if train:
            self.loss_D_A.backward()
            self.loss_D_B.backward()
            self._buckets_D.all_reduce()
            self.optimizer_D.step()
