"""Drop-in for the reference's ``util/image_pool.py``.

Same decisions, in the same order, from the same Python ``random`` stream as the reference
(util/image_pool.py:12-32): fill the pool first, then with probability 1/2 return a random stored image
and store the new one.  The images live in ONE preallocated device tensor [pool_size, C, H, W] and the
returned batch is written in place, so a query is index bookkeeping on the host plus slice copies on the
device — no per-image unsqueeze / clone / cat.  ``trace`` records the most recent decisions for the parity
tests (bounded: a long training run does not grow host memory).
"""
import random

import torch


class _Trace(list):
    """List of the most recent decisions: the older half is dropped whenever MAX entries are reached."""
    MAX = 1 << 16

    def append(self, item):
        if len(self) >= self.MAX:
            del self[:self.MAX // 2]
        list.append(self, item)


class ImagePool():

    def __init__(self, pool_size):
        self.pool_size = pool_size
        self.trace = _Trace()
        if self.pool_size > 0:
            self.num_imgs = 0
            self.images = None      # [pool_size, C, H, W], allocated at the first query

    def query(self, images):
        if self.pool_size == 0:
            return images
        images = images.detach()
        if self.images is None:
            self.images = torch.empty((self.pool_size,) + tuple(images.shape[1:]), dtype=images.dtype,
                                      device=images.device)
        out = torch.empty_like(images)
        for i in range(images.shape[0]):
            if self.num_imgs < self.pool_size:
                self.images[self.num_imgs].copy_(images[i])
                out[i].copy_(images[i])
                self.trace.append(('fill', self.num_imgs))
                self.num_imgs = self.num_imgs + 1
            else:
                p = random.uniform(0, 1)
                if p > 0.5:
                    random_id = random.randint(0, self.pool_size - 1)  # inclusive, as in the reference
                    out[i].copy_(self.images[random_id])
                    self.images[random_id].copy_(images[i])
                    self.trace.append(('swap', random_id))
                else:
                    out[i].copy_(images[i])
                    self.trace.append(('pass', -1))
        return out

    # ------------------------------------------------------------------------------------------------
    # graph-replay form: the same decisions, drawn up-front on the host, applied by one device kernel
    # ------------------------------------------------------------------------------------------------
    def plan(self, batch):
        """Draws the decisions of ONE query of `batch` images from Python's ``random`` exactly as ``query`` would
        (same calls, same order) and returns [(return_from, store_to)] with -1 = the incoming image / no store."""
        if self.pool_size == 0:          # query() returns the batch untouched and draws nothing
            return [(-1, -1)] * batch
        out = []
        for _ in range(batch):
            if self.num_imgs < self.pool_size:
                out.append((-1, self.num_imgs))
                self.trace.append(('fill', self.num_imgs))
                self.num_imgs = self.num_imgs + 1
            else:
                p = random.uniform(0, 1)
                if p > 0.5:
                    random_id = random.randint(0, self.pool_size - 1)
                    out.append((random_id, random_id))
                    self.trace.append(('swap', random_id))
                else:
                    out.append((-1, -1))
                    self.trace.append(('pass', -1))
        return out

    def query_planned(self, images, plan_dev):
        """``query`` with the decisions read from the int32 device table plan_dev [batch, 2] (filled from
        ``plan``); capturable in a CUDA graph. Returns the detached batch the discriminator sees."""
        from . import ops
        if self.pool_size == 0:
            return images
        images = images.detach().contiguous()
        if self.images is None:
            self.images = torch.empty((self.pool_size,) + tuple(images.shape[1:]), dtype=images.dtype,
                                      device=images.device)
        out = torch.empty_like(images)
        ops.image_pool_apply(images, self.images, plan_dev, out)
        return out

