"""The CycleGAN training step of the reference (``models/cycle_gan_model.py:46-160``) on the B200
networks: G_A / G_B ResNet generators, D_A / D_B PatchGANs, LSGAN + L1 cycle + identity losses, two
ImagePools, Adam(lr, (beta1, 0.999)) for G and for D, and the fork's schedule of FOUR discriminator
updates per generator update (:151).

``CycleGANModel`` mirrors the reference class (``initialize(opt)``, ``set_input``, ``forward``,
``backward_G``, ``backward_D_A/B``, ``optimize_parameters(train_or_test)``, ``get_current_losses``) but
does not inherit the reference's BaseModel, whose glue is broken as shipped (SURVEY B-12).  With
``world_size > 1`` (one process per GPU) gradients are averaged with a bucketed NCCL all-reduce before
each optimizer step and the pools are replicated (all-gather of the fakes) so that every rank draws the
same ``random`` sequence over the global batch, as the single-process reference would.
"""
import itertools
from collections import OrderedDict

import torch
import torch.distributed as dist

from . import engine as engine_mod
from . import losses, networks, ops
from .graph_step import StepGraph
from .image_pool import ImagePool


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas, eps=1e-8) semantics on the multi-tensor cdb_adam_pack_multi kernel: ONE launch
    per optimizer step, which also writes the updated filters into the packed bf16 GEMM operands the convolution
    kernels read (SURVEY 8(f) f1 — no separate re-pack pass).

    A real ``torch.optim.Optimizer``: ``param_groups`` / ``state`` / ``state_dict`` / ``load_state_dict`` /
    ``zero_grad`` are the base class's, so ``networks.get_scheduler`` (models/networks.py:24-38) and
    ``update_learning_rate`` (models/base_model.py:91-95) work on it unchanged; the per-parameter state uses
    torch.optim.Adam's keys (``step``, ``exp_avg``, ``exp_avg_sq``).  With ``device_step`` (CUDA-graph replay of the
    training step) the step count AND the learning rate live in device memory: ``sync_lr()`` — called outside the
    graph before every replay — copies ``param_groups[i]['lr']`` to the device scalar the captured kernel reads, so
    a scheduler's change takes effect on the next replay."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, device_step=False):
        super(FusedAdam, self).__init__(params, dict(lr=lr, betas=betas, eps=eps))
        # device_step: step count and learning rate are read from device memory by the kernel, so that a CUDA graph
        # of the training step replays with the right bias corrections and the scheduler's current rate
        self.device_step = device_step
        self._step_dev = None
        self._lr_dev = {}        # group index -> (fp32 device scalar, last value written)

    def _device(self):
        return self.param_groups[0]['params'][0].device

    def sync_lr(self):
        """Copies every group's current learning rate to its device scalar (no-op when unchanged). Must run
        outside a CUDA-graph capture / before a replay."""
        if not self.device_step:
            return
        for gi, g in enumerate(self.param_groups):
            lr = float(g['lr'])
            cur = self._lr_dev.get(gi)
            if cur is None:
                self._lr_dev[gi] = [torch.full((1,), lr, dtype=torch.float32, device=self._device()), lr]
            elif cur[1] != lr:
                cur[0].fill_(lr)
                cur[1] = lr

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        capturing = torch.cuda.is_current_stream_capturing() if torch.cuda.is_available() else False
        if self.device_step:
            if self._step_dev is None:
                self._step_dev = torch.zeros((1,), dtype=torch.int32, device=self._device())
            self._step_dev.add_(1)
            if not capturing:
                self.sync_lr()
        for gi, g in enumerate(self.param_groups):
            b1, b2 = g['betas']
            groups = {}
            updated = []
            for p in g['params']:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                step = int(st['step']) + 1      # a loaded torch.optim.Adam state holds a tensor here
                st['step'] = step
                if p.grad.dtype != torch.float32 or not p.data.is_contiguous():
                    raise TypeError("FusedAdam: contiguous fp32 parameters and gradients expected")
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                packs = engine_mod._pack_cache.adam_targets(p)
                groups.setdefault(step, []).append((p.data, grad, st['exp_avg'], st['exp_avg_sq'], packs))
                updated.append((p, packs))
            # one multi-tensor launch per distinct step count (normally exactly one)
            lr_dev = self._lr_dev[gi][0] if (self.device_step and gi in self._lr_dev) else None
            for step, items in groups.items():
                ops.adam_pack_multi(items, g['lr'], b1, b2, g['eps'], step,
                                    self._step_dev if self.device_step else None, lr_dev)
            for p, packs in updated:
                # the kernel writes through raw pointers, which autograd's version counter cannot
                # see; the packed-weight cache also keys on this explicit counter
                p._cdb_version = getattr(p, '_cdb_version', 0) + 1
                engine_mod._pack_cache.stamp(p, packs)
            # cached layouts the kernel did not emit (a third layout of a filter, flipped / folded 7x7 forms):
            # one multi-tensor re-pack for the plain ones, the rest re-pack lazily at their next use
            engine_mod._pack_cache.refresh([p for p, _ in updated])
        return loss


class GradBuckets:
    """Gradient averaging over the data-parallel ranks.

    NCCL: ONE grouped collective per call — every gradient tensor is all-reduced in place (ncclGroupStart / End through
    torch's coalescing manager, ReduceOp.AVG), so nothing is flattened, concatenated or copied back (round 1 spent
    three extra passes over 91 MB + 4 x 22 MB per step on that).  The call is issued on the CURRENT stream: the step
    mirrors run it on a side stream so that the generators' all-reduce and Adam update overlap the discriminator
    phase.  Other backends (gloo in the CPU tests): ~25 MB flat buckets, SUM then divide."""

    def __init__(self, params, bucket_bytes=25 << 20):
        self.params = [p for p in params]
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)

    def all_reduce(self):
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
            return
        world = dist.get_world_size()
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads:
            return
        if dist.get_backend() == 'nccl' and hasattr(dist, '_coalescing_manager'):
            with dist._coalescing_manager(device=grads[0].device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.AVG)
            return
        works = []
        for bucket in self.buckets:
            bg = [p.grad for p in bucket if p.grad is not None]
            if not bg:
                continue
            flat = torch.cat([g.reshape(-1) for g in bg])
            works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, bg))
        for work, flat, bg in works:
            work.wait()
            flat.div_(world)
            off = 0
            for g in bg:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()


class CycleGANModel:
    D_ITERS = 4     # discriminator updates per generator update (the fork's schedule, models/cycle_gan_model.py:151)

    def name(self):
        return 'CycleGANModel'

    def initialize(self, opt):
        """opt: namespace with the fields of models/checkpoints/app2orange_cycle/opt.txt that the
        reference reads at models/cycle_gan_model.py:32-69 (input_nc, output_nc, ngf, ndf, netG, netD,
        n_layers_D, norm, no_dropout, init_type, init_gain, no_lsgan, pool_size, lr, beta1, lambda_A,
        lambda_B, lambda_identity, isTrain) plus ``device``."""
        self.opt = opt
        self.isTrain = opt.isTrain
        self.device = torch.device(getattr(opt, 'device', 'cuda'))
        gpu_ids = [self.device]
        self.loss_names = ['D_A', 'G_A', 'cycle_A', 'idt_A', 'D_B', 'G_B', 'cycle_B', 'idt_B']
        self.model_names = ['G_A', 'G_B', 'D_A', 'D_B'] if self.isTrain else ['G_A', 'G_B']
        self.netG_A = networks.define_G(opt.input_nc, opt.output_nc, opt.ngf, opt.netG, opt.norm, not opt.no_dropout,
                                        opt.init_type, opt.init_gain, gpu_ids)
        self.netG_B = networks.define_G(opt.output_nc, opt.input_nc, opt.ngf, opt.netG, opt.norm, not opt.no_dropout,
                                        opt.init_type, opt.init_gain, gpu_ids)
        if self.isTrain:
            use_sigmoid = opt.no_lsgan
            self.netD_A = networks.define_D(opt.output_nc, opt.ndf, opt.netD, opt.n_layers_D, opt.norm, use_sigmoid,
                                            opt.init_type, opt.init_gain, gpu_ids)
            self.netD_B = networks.define_D(opt.input_nc, opt.ndf, opt.netD, opt.n_layers_D, opt.norm, use_sigmoid,
                                            opt.init_type, opt.init_gain, gpu_ids)
            self.fake_A_pool = ImagePool(opt.pool_size)
            self.fake_B_pool = ImagePool(opt.pool_size)
            self.criterionGAN = networks.GANLoss(use_lsgan=not opt.no_lsgan).to(self.device)
            self.criterionCycle = losses.L1Loss()
            self.criterionIdt = losses.L1Loss()
            self.build_optimizers()

    def build_optimizers(self):
        opt = self.opt
        self._graph_mode = bool(getattr(opt, 'cuda_graph', False))
        if getattr(opt, 'fused_adam', True):
            adam = lambda ps, **kw: FusedAdam(ps, device_step=self._graph_mode, **kw)
        else:
            adam = torch.optim.Adam
        self.optimizer_G = adam(itertools.chain(self.netG_A.parameters(), self.netG_B.parameters()),
                                lr=opt.lr, betas=(opt.beta1, 0.999))
        self.optimizer_D = adam(itertools.chain(self.netD_A.parameters(), self.netD_B.parameters()),
                                lr=opt.lr, betas=(opt.beta1, 0.999))
        self._step_graph, self._plan_host, self._plan_dev, self._plan_slot = StepGraph(), None, None, 0
        self.optimizers = [self.optimizer_G, self.optimizer_D]
        self._buckets_G = GradBuckets(itertools.chain(self.netG_A.parameters(), self.netG_B.parameters()))
        self._buckets_D = GradBuckets(itertools.chain(self.netD_A.parameters(), self.netD_B.parameters()))

    def set_input(self, input):
        if getattr(self, '_graph_mode', False):
            # static input buffers: a captured step keeps reading the same device addresses
            a, b = input['img_source'], input['img_target']
            if getattr(self, 'real_A', None) is None or self.real_A.shape != a.shape:
                if self._graph is not None:
                    raise RuntimeError("cuda_graph mode: the input shape changed after the step was captured")
                if getattr(self, 'real_A', None) is None or self.real_A.shape != a.shape:
                    self.real_A = torch.empty(tuple(a.shape), dtype=torch.float32, device=self.device)
                    self.real_B = torch.empty(tuple(b.shape), dtype=torch.float32, device=self.device)
            self.real_A.copy_(a, non_blocking=True)
            self.real_B.copy_(b, non_blocking=True)
            return
        self.real_A = input['img_source'].to(self.device, non_blocking=True)
        self.real_B = input['img_target'].to(self.device, non_blocking=True)

    def set_requires_grad(self, nets, requires_grad=False):
        if not isinstance(nets, list):
            nets = [nets]
        for net in nets:
            if net is not None:
                for param in net.parameters():
                    param.requires_grad = requires_grad

    def _batched(self):
        """opt.batch_passes (default on): passes of one network over independent inputs run as ONE launch sequence on
        the concatenated batch.  InstanceNorm normalises per sample and the losses are means over equal-sized halves,
        so every per-sample result is the one the separate passes of models/cycle_gan_model.py:80-99,111-137 produce;
        what changes is the launch count (6 generator passes -> 3, 2 discriminator passes per update -> 1) and the
        tile occupancy of the small discriminator layers.  Networks with BatchNorm couple the samples: not batched."""
        return bool(getattr(self.opt, 'batch_passes', True)) and self.opt.norm == 'instance'

    def forward(self):
        b = self.real_A.shape[0]
        self._idt_A = self._idt_B = None
        if self._batched() and self.isTrain and self.real_A.shape == self.real_B.shape:
            want_idt = self.opt.lambda_identity > 0
            # G_A over [real_A | real_B]: fake_B and (if used) the identity output idt_A = G_A(real_B)
            out = self.netG_A(torch.cat([self.real_A, self.real_B], 0) if want_idt else self.real_A)
            self.fake_B = out[:b]
            if want_idt:
                self._idt_A = out[b:]
            # G_B over [real_B | fake_B | real_A]: fake_A, rec_A and idt_B = G_B(real_A)
            parts = [self.real_B, self.fake_B] + ([self.real_A] if want_idt else [])
            out = self.netG_B(torch.cat(parts, 0))
            self.fake_A, self.rec_A = out[:b], out[b:2 * b]
            if want_idt:
                self._idt_B = out[2 * b:]
            self.rec_B = self.netG_A(self.fake_A)
            return
        self.fake_B = self.netG_A(self.real_A)
        self.rec_A = self.netG_B(self.fake_B)
        self.fake_A = self.netG_B(self.real_B)
        self.rec_B = self.netG_A(self.fake_A)

    def _gather_fake(self, fake):
        """All ranks' fakes in rank-major order [world * B, ...].  The fakes do not change during the D_ITERS
        discriminator updates of a step, so each pool's all-gather runs once per step, not once per query."""
        cache = self.__dict__.setdefault('_gather_cache', {})
        hit = cache.get(id(fake))
        if hit is not None and hit[0] is fake:
            return hit[1]
        world = dist.get_world_size()
        gathered = torch.empty((world,) + tuple(fake.shape), dtype=fake.dtype, device=fake.device)
        dist.all_gather_into_tensor(gathered, fake.detach().contiguous())
        gathered = gathered.view((-1,) + tuple(fake.shape[1:]))
        if len(cache) > 4:
            cache.clear()
        cache[id(fake)] = (fake, gathered)
        return gathered

    def _pool_query(self, pool, fake):
        """Replicated pool under data parallelism: all ranks see the global batch in rank-major order
        and replay the identical random stream; each rank keeps its own slice of the result."""
        if getattr(self, '_plan_dev', None) is not None and getattr(self, '_in_graphed_step', False):
            slot = self._plan_slot
            self._plan_slot += 1
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                rank = dist.get_rank()
                out = pool.query_planned(self._gather_fake(fake), self._plan_dev[slot])
                b = fake.shape[0]
                return out[rank * b:(rank + 1) * b]
            return pool.query_planned(fake, self._plan_dev[slot])
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            rank = dist.get_rank()
            if fake.is_cuda:
                out = pool.query(self._gather_fake(fake))
            else:
                gathered = [torch.empty_like(fake) for _ in range(dist.get_world_size())]
                dist.all_gather(gathered, fake.detach().contiguous())
                out = pool.query(torch.cat(gathered, 0))
            b = fake.shape[0]
            return out[rank * b:(rank + 1) * b]
        return pool.query(fake)

    def backward_D_basic(self, netD, real, fake):
        if self._batched() and real.shape == fake.shape:
            pred = netD(torch.cat([real, fake], 0))
            b = real.shape[0]
            return (self.criterionGAN(pred[:b], True) + self.criterionGAN(pred[b:], False)) * 0.5
        loss_D_real = self.criterionGAN(netD(real), True)
        loss_D_fake = self.criterionGAN(netD(fake), False)
        return (loss_D_real + loss_D_fake) * 0.5

    def backward_D_A(self):
        fake_B = self._pool_query(self.fake_B_pool, self.fake_B)
        self.loss_D_A = self.backward_D_basic(self.netD_A, self.real_B, fake_B)
        return self.loss_D_A

    def backward_D_B(self):
        fake_A = self._pool_query(self.fake_A_pool, self.fake_A)
        self.loss_D_B = self.backward_D_basic(self.netD_B, self.real_A, fake_A)
        return self.loss_D_B

    # The two discriminators are independent networks on independent inputs, and their layers are small (a PatchGAN
    # layer at batch 16 is 10-35 us and does not fill 148 SMs): one update runs D_A (forward, loss, backward) and D_B
    # on two streams.  The pool queries stay on the calling stream and in the reference's order (B then A,
    # models/cycle_gan_model.py:101-109), so the random stream is consumed exactly as before.  Inside a CUDA-graph
    # capture the two streams become parallel branches of the graph.
    def _d_streams(self):
        st = getattr(self, '_d_side_streams', None)
        if st is None:
            st = (torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device))
            self._d_side_streams = st
        return st

    def _update_D_concurrently(self, train):
        fake_B = self._pool_query(self.fake_B_pool, self.fake_B)
        fake_A = self._pool_query(self.fake_A_pool, self.fake_A)
        cur = torch.cuda.current_stream()
        sa, sb = self._d_streams()
        for side, net, real, fake, name in ((sa, self.netD_A, self.real_B, fake_B, 'loss_D_A'),
                                            (sb, self.netD_B, self.real_A, fake_A, 'loss_D_B')):
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                loss = self.backward_D_basic(net, real, fake)
                setattr(self, name, loss)
                if train:
                    loss.backward()
        cur.wait_stream(sa)
        cur.wait_stream(sb)

    def backward_G(self):
        lambda_idt, lambda_A, lambda_B = self.opt.lambda_identity, self.opt.lambda_A, self.opt.lambda_B
        if lambda_idt > 0:
            self.idt_A = self._idt_A if getattr(self, '_idt_A', None) is not None else self.netG_A(self.real_B)
            self.loss_idt_A = self.criterionIdt(self.idt_A, self.real_B) * lambda_B * lambda_idt
            self.idt_B = self._idt_B if getattr(self, '_idt_B', None) is not None else self.netG_B(self.real_A)
            self.loss_idt_B = self.criterionIdt(self.idt_B, self.real_A) * lambda_A * lambda_idt
        else:
            self.loss_idt_A = 0
            self.loss_idt_B = 0
        if bool(getattr(self.opt, 'concurrent_D', True)) and self.fake_B.is_cuda:
            # the two (frozen) discriminator passes of the generator objective on the two discriminator streams: their
            # small layers, and later their data gradients, overlap each other and the generators' backward
            cur = torch.cuda.current_stream()
            sa, sb = self._d_streams()
            for side, net, fake, name in ((sa, self.netD_A, self.fake_B, 'loss_G_A'), (sb, self.netD_B, self.fake_A, 'loss_G_B')):
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    setattr(self, name, self.criterionGAN(net(fake), True))
            cur.wait_stream(sa)
            cur.wait_stream(sb)
        else:
            self.loss_G_A = self.criterionGAN(self.netD_A(self.fake_B), True)
            self.loss_G_B = self.criterionGAN(self.netD_B(self.fake_A), True)
        self.loss_cycle_A = self.criterionCycle(self.rec_A, self.real_A) * lambda_A
        self.loss_cycle_B = self.criterionCycle(self.rec_B, self.real_B) * lambda_B
        self.loss_G = (self.loss_G_A + self.loss_G_B + self.loss_cycle_A + self.loss_cycle_B + self.loss_idt_A
                       + self.loss_idt_B)
        return self.loss_G

    # ---------------------------------------------------------------------------------------------
    # CUDA-graph replay of the whole step (opt.cuda_graph, single process): ~3200 kernel launches per step
    # become one graph launch. Host-side decisions are kept out of the graph: the ImagePool's random draws
    # are made up-front (ImagePool.plan, same ``random`` calls in the same order as the eager step) and
    # reach the device through a pinned table copied by the graph's first node; Adam reads its step count
    # from device memory.
    # ---------------------------------------------------------------------------------------------
    GRAPH_WARMUP_STEPS = StepGraph.WARMUP_STEPS

    def _draw_pool_plans(self):
        b = self.real_A.shape[0]
        if dist.is_available() and dist.is_initialized():
            b *= dist.get_world_size()          # replicated pools see the global batch in rank-major order
        if self._plan_host is None:
            self._plan_host = torch.empty((2 * self.D_ITERS, b, 2), dtype=torch.int32).pin_memory()
            self._plan_dev = torch.empty((2 * self.D_ITERS, b, 2), dtype=torch.int32, device=self.device)
        self._step_graph.wait_previous()        # the previous replay has consumed the pinned table
        rows = []
        for _ in range(self.D_ITERS):           # the order of backward_D_A / backward_D_B in the step
            rows.append(self.fake_B_pool.plan(b))
            rows.append(self.fake_A_pool.plan(b))
        self._plan_host.copy_(torch.tensor(rows, dtype=torch.int32))
        self._plan_slot = 0

    @property
    def _graph(self):
        return self._step_graph.graph

    @property
    def _graph_launches(self):
        return self._step_graph.launches

    def _graphed_step(self):
        self._draw_pool_plans()
        if self._step_graph.graph is None and self._step_graph.calls >= StepGraph.WARMUP_STEPS:
            self.optimizer_G.zero_grad()        # gradients must be (re)allocated inside the capture
            self.optimizer_D.zero_grad()

        def step():
            self._plan_dev.copy_(self._plan_host, non_blocking=True)
            self._eager_step(True)
        # the planned (device-table) pool path is taken only here: a validation step ('test') after the capture
        # goes through the ordinary host-side pool.query (models/train.py:35-41 runs 'test' steps between epochs)
        self._in_graphed_step = True
        try:
            self._step_graph.run(step, self.optimizers)
        finally:
            self._in_graphed_step = False

    def optimize_parameters(self, train_or_test='train'):
        if getattr(self, '_graph_mode', False) and train_or_test == 'train':
            return self._graphed_step()
        return self._eager_step(train_or_test == 'train')

    def _eager_step(self, train):
        self.forward()
        self.set_requires_grad([self.netD_A, self.netD_B], False)
        self.optimizer_G.zero_grad()
        self.loss_G = self.backward_G()
        g_stream = None
        if train:
            self.loss_G.backward()
            if self.real_A.is_cuda and bool(getattr(self.opt, 'overlap_G_update', True)):
                # the discriminator phase below reads the fakes and the D weights only: the generators' gradient
                # all-reduce and Adam update run on a side stream underneath it and are joined at the end of the step
                g_stream = self.__dict__.get('_g_update_stream')
                if g_stream is None:
                    g_stream = self.__dict__['_g_update_stream'] = torch.cuda.Stream(device=self.device)
                g_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(g_stream):
                    self._buckets_G.all_reduce()
                    self.optimizer_G.step()
            else:
                self._buckets_G.all_reduce()
                self.optimizer_G.step()
        concurrent = bool(getattr(self.opt, 'concurrent_D', True)) and self.real_A.is_cuda
        for _ in range(self.D_ITERS):
            self.set_requires_grad([self.netD_A, self.netD_B], True)
            self.optimizer_D.zero_grad()
            if concurrent:
                self._update_D_concurrently(train)
            else:
                self.loss_D_A = self.backward_D_A()
                self.loss_D_B = self.backward_D_B()
                if train:
                    self.loss_D_A.backward()
                    self.loss_D_B.backward()
            if train:
                self._buckets_D.all_reduce()
                self.optimizer_D.step()
        if g_stream is not None:
            torch.cuda.current_stream().wait_stream(g_stream)
        self.__dict__.pop('_gather_cache', None)

    def get_current_losses(self):
        out = OrderedDict()
        for name in self.loss_names:
            v = getattr(self, 'loss_' + name)
            out[name] = float(v.detach()) if torch.is_tensor(v) else float(v)
        return out
