"""The joint segmentation + depth training step of the reference (``new_multi/model5.py:199-696``) on the B200
networks of ``networks5_ds.py``.

``Seg_Depth`` mirrors the reference class: ``initialize(opt)``, ``set_input(input, train_or_test)``,
``optimize_parameters(train_or_test)`` = ``backward_G_2`` (:585-638) -> step G_2, ``backward_G_1`` (:564-583) ->
step G_1, ``backward_R_D`` (:479-559) -> two R_D steps, ``backward_DISDEP`` (:415-474) -> three feature
discriminator steps; six Adam optimizers with lr/5, lr/3, lr/2, lr/4 (:250-275); CrossEntropy(ignore 255), L1,
LSGAN-MSE and BCEDepLoss (:280-285).  It does not inherit the reference's BaseModel and does not load the
authors' absolute-path checkpoints (:213-223): pass ``opt.g1_checkpoint`` / ``opt.g2_checkpoint`` to load
reference-layout state_dicts, otherwise the networks are initialised with ``init_net``.  The reference's prints
inside the step are dropped; its arithmetic quirks are kept (L1 between [B,1,H,W] and [B,H,W] broadcasts to
[B,B,H,W]; ``detach_list`` does not detach; G_2's features reach R_dep detached).

Data parallelism: one process per GPU; gradients of the network being stepped are averaged with the bucketed
all-reduce of ``cycle_gan_model.GradBuckets``.  BatchNorm statistics stay per rank, like under the reference's
nn.DataParallel (``networks5_ds.py:258-259``).
"""
from collections import OrderedDict

import torch

from . import losses, networks5_ds
from .cycle_gan_model import FusedAdam, GradBuckets
from .graph_step import StepGraph
from .image_pool import ImagePool
from .networks5_ds import get_masks, init_net


class Seg_Depth:
    def name(self):
        return 'Seg_Depth_Model'

    def initialize(self, opt):
        """opt: namespace with lr, beta1, pool_size (+ optional g1_checkpoint, g2_checkpoint, init_type, init_gain)."""
        self.opt = opt
        self.is_Train = True
        itype, igain = getattr(opt, 'init_type', 'normal'), getattr(opt, 'init_gain', 0.02)
        self.net_FD1 = init_net(networks5_ds._Discriminator(input_nc=512), itype, igain)
        self.net_FD2 = init_net(networks5_ds._Discriminator(input_nc=256), itype, igain)
        self.net_FD3 = init_net(networks5_ds._Discriminator(input_nc=128), itype, igain)
        self.net_G_1 = self._build(networks5_ds.G_1(), getattr(opt, 'g1_checkpoint', None), itype, igain)
        self.net_G_2 = self._build(networks5_ds.General_net(), getattr(opt, 'g2_checkpoint', None), itype, igain)
        self.net_R_D = init_net(networks5_ds.R_dep(), itype, igain)
        # opt.cuda_graph (single process): the whole step — eight optimizer updates included — replays as one graph
        self._graph_mode = bool(getattr(opt, 'cuda_graph', False))
        self._step_graph = StepGraph()
        if getattr(opt, 'fused_adam', True):
            adam = lambda ps, **kw: FusedAdam(ps, device_step=self._graph_mode, **kw)
        else:
            adam = torch.optim.Adam
        mk = lambda net, f: adam(net.parameters(), lr=opt.lr / f, betas=(opt.beta1, 0.999))
        self.optimizer_G_1, self.optimizer_G_2, self.optimizer_R_D = mk(self.net_G_1, 5), mk(self.net_G_2, 3), \
            mk(self.net_R_D, 2)
        self.optimizer_FD1, self.optimizer_FD2, self.optimizer_FD3 = mk(self.net_FD1, 4), mk(self.net_FD2, 4), \
            mk(self.net_FD3, 4)
        self._buckets = {n: GradBuckets(getattr(self, 'net_' + n).parameters())
                         for n in ('G_1', 'G_2', 'R_D', 'FD1', 'FD2', 'FD3')}
        self.syn_imgpool = ImagePool(opt.pool_size)
        self.real_imgpool = ImagePool(opt.pool_size)
        self.criterionGAN = networks5_ds.GANLoss(use_lsgan=True)
        self.criterionSeg = losses.CrossEntropyLoss(size_average=True, ignore_index=255)
        self.criterionDep = losses.L1Loss()
        self.criterionDep_bce = networks5_ds.BCEDepLoss()
        self.loss_names = ['G2', 'G1', 'RD_real', 'RD_syn', 'dep_ref', 'FD1', 'FD2', 'FD3']

    @staticmethod
    def _build(net, checkpoint, itype, igain):
        if checkpoint is None:
            return init_net(net, itype, igain)
        wrapped = networks5_ds._Replica(net.cuda())
        wrapped.load_state_dict(torch.load(checkpoint, map_location='cuda'), strict=True)
        return wrapped

    def set_input(self, input, train_or_test='train'):
        dev = torch.device('cuda')
        self.is_Train = train_or_test == 'train'
        fields = [('real_img', input['img_real']), ('syn_img', input['img_syn']),
                  ('syn_seg_l', input['seg_l_syn'].squeeze(1)), ('syn_dep_l', input['dep_l_syn'].squeeze(1)),
                  ('syn_dep_ls', input['depth_l_s'].float())]
        if self.is_Train:
            fields.append(('real_seg_l', input['seg_l_real'].squeeze(1)))
        for name, src in fields:
            if self._graph_mode:
                # static buffers: the captured step keeps reading the same device addresses
                cur = getattr(self, name, None)
                if cur is None or cur.shape != src.shape or cur.dtype != src.dtype:
                    if self._step_graph.graph is not None:
                        raise RuntimeError("cuda_graph mode: the input shape changed after the step was captured")
                    cur = torch.empty(tuple(src.shape), dtype=src.dtype, device=dev)
                    setattr(self, name, cur)
                cur.copy_(src, non_blocking=True)
            else:
                setattr(self, name, src.to(dev, non_blocking=True))

    def set_requires_grad(self, nets, requires_grad=False):
        if not isinstance(nets, list):
            nets = [nets]
        for net in nets:
            if net is not None:
                for param in net.parameters():
                    param.requires_grad = requires_grad

    @staticmethod
    def _sky_mask(seg_l):
        """model5.py:524-526: 0 where the label is 17 (sky), 1 elsewhere (written without boolean-mask
        assignment, which synchronises with the host and cannot be captured in a CUDA graph)."""
        return (seg_l != 17).to(seg_l.dtype)

    # ------------------------------------------------------------------ :585-638
    def backward_G_2(self):
        self.set_requires_grad([self.net_R_D, self.net_G_1], False)
        self.set_requires_grad([self.net_FD1, self.net_FD2], False)
        ss = self.net_G_1(self.syn_img)
        syn_features1, syn_Features = self.net_G_2(ss.detach(), 'S')
        feats, seg, (dep_4, dep_o) = self.net_R_D(syn_Features, syn_features1)
        self.syn_feats = feats
        sky_m = self._sky_mask(self.syn_seg_l)
        dep_loss = self.criterionDep(dep_o, sky_m.float() * self.syn_dep_l)
        s_seg_loss = self.criterionSeg(seg, self.syn_seg_l)
        D_syn_loss = dep_loss + s_seg_loss
        self.syn_features1 = syn_features1.detach()
        self.syn_Features = syn_Features
        del syn_features1, syn_Features, feats
        real_features1, real_Features = self.net_G_2(self.real_img, 'R')
        feats, seg, (dep_4, dep_o) = self.net_R_D(real_Features, real_features1)
        self.real_features1 = real_features1.detach()
        self.real_Features = real_Features
        seg_loss_real = self.criterionSeg(seg, self.real_seg_l) if self.is_Train else 0
        return D_syn_loss + 2 * seg_loss_real

    # ------------------------------------------------------------------ :564-583
    def backward_G_1(self):
        self.set_requires_grad([self.net_R_D, self.net_G_2], False)
        self.set_requires_grad([self.net_G_1], True)
        ss = self.net_G_1(self.syn_img)
        syn_features1, syn_Features = self.net_G_2(ss, 'S')
        s_feats, s_seg, (s_dep_4, s_dep_o) = self.net_R_D(syn_Features, syn_features1)
        loss_dep = self.criterionDep(s_dep_o, self.syn_dep_l)
        loss_seg_syn = self.criterionSeg(s_seg, self.syn_seg_l)
        return loss_seg_syn + loss_dep

    # ------------------------------------------------------------------ :479-559
    def backward_R_D(self, train_or_test):
        train = train_or_test == 'train'
        self.optimizer_R_D.zero_grad()
        feats, seg, (dep_4, dep_o) = self.net_R_D(self.real_Features, self.real_features1)
        seg_loss_real = self.criterionSeg(seg, self.real_seg_l) if self.is_Train else 0
        pred1, pred2, pred3 = self.net_FD1(feats[0]), self.net_FD2(feats[1]), self.net_FD3(feats[2])
        D_real_loss = (seg_loss_real + 0.2 * self.criterionGAN(pred1, False) + 0.2 * self.criterionGAN(pred2, False)
                       + 0.2 * self.criterionGAN(pred3, False))
        self.loss_RD_real = D_real_loss.detach()
        self.real_dep_ref = dep_o.squeeze(1).detach()
        if train:
            D_real_loss.backward()
            self._buckets['R_D'].all_reduce()
            self.optimizer_R_D.step()
        self.real_feats = feats
        self.optimizer_R_D.zero_grad()
        del seg, dep_4, dep_o
        feats, seg, (dep_4, dep_o) = self.net_R_D(self.syn_Features, self.syn_features1)
        sky_m = self._sky_mask(self.syn_seg_l)
        sky4 = torch.cat([sky_m.unsqueeze(1)] * 4, 1).float() * self.syn_dep_ls.clone()
        oms, zms = get_masks(sky4)
        dep_loss = self.criterionDep(dep_o, sky_m.float() * self.syn_dep_l)
        s_seg_loss = self.criterionSeg(seg, self.syn_seg_l)
        for s_Dep in dep_4:
            dep_loss = dep_loss + self.criterionDep_bce(sky_m.unsqueeze(1).float() * s_Dep, sky4, oms, zms)
        D_syn_loss = dep_loss + s_seg_loss
        if train:
            D_syn_loss.backward()
            self._buckets['R_D'].all_reduce()
            self.optimizer_R_D.step()
        self.loss_RD_syn = D_syn_loss.detach()
        self.loss_dep_ref = dep_loss.detach()
        self.syn_dep_ref = dep_o.squeeze(1).detach()
        self.syn_feats = feats
        return D_syn_loss

    # ------------------------------------------------------------------ :415-474
    def backward_DISDEP(self):
        for i, (net, optim, name) in enumerate(((self.net_FD1, self.optimizer_FD1, 'FD1'),
                                                (self.net_FD2, self.optimizer_FD2, 'FD2'),
                                                (self.net_FD3, self.optimizer_FD3, 'FD3'))):
            optim.zero_grad()
            D_real = net(self.real_feats[i].detach())
            D_fake = net(self.syn_feats[i].detach())
            loss = self.criterionGAN(D_real, True) + self.criterionGAN(D_fake, False)
            loss.backward()
            self._buckets[name].all_reduce()
            optim.step()
            setattr(self, 'loss_' + name, loss.detach())
        self.set_requires_grad([self.net_FD1, self.net_FD2, self.net_FD3], False)

    # ------------------------------------------------------------------ :640-696
    def optimize_parameters(self, train_or_test='train'):
        if self._graph_mode and train_or_test == 'train':
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                raise RuntimeError("opt.cuda_graph is a single-process mode (data-parallel steps run eagerly)")
            sg = self._step_graph
            if sg.graph is None and sg.calls >= StepGraph.WARMUP_STEPS:
                for o in (self.optimizer_G_1, self.optimizer_G_2, self.optimizer_R_D, self.optimizer_FD1,
                          self.optimizer_FD2, self.optimizer_FD3):
                    o.zero_grad()               # gradients must be (re)allocated inside the capture
            return sg.run(lambda: self._eager_step('train'),
                          (self.optimizer_G_1, self.optimizer_G_2, self.optimizer_R_D, self.optimizer_FD1,
                           self.optimizer_FD2, self.optimizer_FD3))
        return self._eager_step(train_or_test)

    def _eager_step(self, train_or_test='train'):
        train = train_or_test == 'train'
        self.set_requires_grad(self.net_G_2, True)
        self.optimizer_G_2.zero_grad()
        self.loss_G2 = self.backward_G_2()
        if train:
            self.loss_G2.backward()
            self._buckets['G_2'].all_reduce()
            self.optimizer_G_2.step()
        self.set_requires_grad([self.net_G_1], True)
        self.set_requires_grad([self.net_G_2], False)
        self.optimizer_G_1.zero_grad()
        self.loss_G1 = self.backward_G_1()
        if train:
            self.loss_G1.backward()
            self._buckets['G_1'].all_reduce()
            self.optimizer_G_1.step()
        self.set_requires_grad([self.net_G_1, self.net_G_2], False)
        self.set_requires_grad(self.net_R_D, True)
        self.backward_R_D(train_or_test)
        if train:
            self.set_requires_grad([self.net_G_1, self.net_G_2, self.net_R_D], False)
            self.set_requires_grad([self.net_FD1, self.net_FD2, self.net_FD3], True)
            self.backward_DISDEP()

    def get_current_losses(self):
        out = OrderedDict()
        for name in self.loss_names:
            v = getattr(self, 'loss_' + name, None)
            if v is not None:
                out[name] = float(v.detach()) if torch.is_tensor(v) else float(v)
        return out
