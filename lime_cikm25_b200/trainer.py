"""Host-side mirror of the reference's training step (trainer.py:21-148) on top of the differentiable
B200 path: same loss, gradient clipping and optimizer (the reference's own torch.optim.Adam and
clip_grad_norm_ — optimizer state is host-side bookkeeping, not part of the model plugin), plus a
working data-parallel variant (the reference's distributed_train, trainer.py:246-426, is stale: it
unpacks 21 of the dataset's 25 tensors and crashes on the first batch — SURVEY.md section 2.1).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.optim as optim


def negative_log_softmax(logits):
    """trainer.py:71-73: the positive candidate is column 0."""
    return (-torch.log_softmax(logits, dim=1).select(dim=1, index=0)).mean()


def remaining_lifetime(config, news_category, news_freshness, news_user_topic_lifetime):
    """trainer.py:121-129 / util.py:98-106."""
    if config.lifetime_type == "fixed":
        return config.fixed_lifetime - news_freshness
    if config.lifetime_type == "topic_wise":
        return config.category_lifetime_map[news_category] - news_freshness
    if config.lifetime_type == "user_topic":
        return news_user_topic_lifetime - news_freshness
    raise ValueError("Invalid lifetime_type")


def allreduce_gradients(params, group=None, flat=None):
    """Data-parallel gradient averaging: ONE all-reduce over a flat fp32 buffer of every gradient (NCCL over NVLink on
    the box, gloo in the CPU tests).  Identical to what DistributedDataParallel computes (mean over ranks), done after
    backward so that the clip-by-global-norm of trainer.py:147 sees the reduced gradients, as with DDP +
    clip_grad_norm_ (trainer.py:334-337).  ``flat``: the buffer the parameters' ``.grad`` tensors are views of
    (``Trainer`` keeps one, like DDP's gradient_as_bucket_view): reduced in place, no gather / scatter copies.  Without
    it the gradients are concatenated, reduced and copied back."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    if flat is not None:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        return flat.numel() * 4
    # every rank must contribute the same layout: parameters without a gradient on this rank send zeros
    plist = [p for p in params if p.requires_grad]
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in plist])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    off = 0
    for p in plist:
        n = p.numel()
        if p.grad is not None:
            p.grad.copy_(flat[off:off + n].view_as(p))
        off += n
    return flat.numel() * 4


class Trainer:
    """One optimisation step exactly as trainer.py:89-148 does it (loss + auxiliary losses, zero_grad,
    backward, clip_grad_norm_(gradient_clip_norm), Adam(lr, weight_decay))."""

    def __init__(self, model, config, group=None):
        self.model, self.config, self.group = model, config, group
        params = [p for p in model.parameters() if p.requires_grad]
        # the reference's Adam (trainer.py:33); on the device torch's fused implementation of the same update (one multi-tensor
        # kernel per step instead of a dozen foreach passes over the 18 M parameters)
        fused = bool(params) and all(p.is_cuda and p.dtype == torch.float32 for p in params)
        self.optimizer = optim.Adam(params, lr=config.lr, weight_decay=config.weight_decay, **({"fused": True} if fused else {}))
        self.gradient_clip_norm = config.gradient_clip_norm
        # every .grad is a view of ONE flat buffer (zeroed per step instead of zero_grad's set-to-None): autograd accumulates
        # into the views in place, the data-parallel all-reduce runs on the buffer itself
        # (only with weight_decay == 0, the reference default: a parameter that never receives a gradient -- the dead ISAB /
        # MAB blocks -- then sees a zero gradient, which Adam ignores exactly like the reference's ``grad is None``)
        plist = [p for p in model.parameters() if p.requires_grad]
        self._flat = None
        if plist and config.weight_decay == 0:
            self._flat = torch.zeros(sum(p.numel() for p in plist), dtype=torch.float32, device=plist[0].device)
            off = 0
            for p in plist:
                p.grad = self._flat[off:off + p.numel()].view_as(p)
                off += p.numel()

    def step(self, batch):
        """batch: the 25 tensors of Train_Dataset.__getitem__ (dataset.py:105-141), already on the device."""
        cfg, model = self.config, self.model
        news_category, news_freshness, news_life = batch[15], batch[23], batch[24]
        rem = remaining_lifetime(cfg, news_category, news_freshness, news_life)
        logits = model(*batch, rem)
        loss = negative_log_softmax(logits)
        if model.news_encoder.auxiliary_loss is not None:                    # trainer.py:137-139
            loss = loss + model.news_encoder.auxiliary_loss.mean()
        if model.user_encoder.auxiliary_loss is not None:                    # trainer.py:140-142
            loss = loss + model.user_encoder.auxiliary_loss.mean()
        if self._flat is not None:
            self._flat.zero_()                                              # == optimizer.zero_grad(), keeping the views
        else:
            self.optimizer.zero_grad()
        loss.backward()
        self.allreduce_bytes = allreduce_gradients(model.parameters(), self.group, flat=self._flat)
        if self.gradient_clip_norm > 0:
            if self._flat is not None:
                # clip_grad_norm_ (trainer.py:147) on the flat buffer every .grad is a view of: the global 2-norm is the norm of
                # the concatenation, the same clamp(max_norm / (norm + 1e-6), max=1) factor -- 2 kernels instead of foreach passes
                total = torch.linalg.vector_norm(self._flat)
                self._flat.mul_((self.gradient_clip_norm / (total + 1e-6)).clamp(max=1.0))
            else:
                nn.utils.clip_grad_norm_(model.parameters(), self.gradient_clip_norm)
        self.optimizer.step()
        return loss.detach()
