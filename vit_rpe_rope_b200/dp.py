"""Batch data-parallel training for the ViT hot path: one process per GPU, one NCCL communicator,
gradients all-reduced in flat buckets that overlap with the rest of backward.

The reference has no multi-GPU code; images are independent units, so forward/backward need no
exchange and the single collective is the gradient sum before ``optimizer.step``.  Design:

* parameters are grouped into flat fp32 buckets (~``bucket_mb`` each) in *reverse registration
  order* - the order backward produces their gradients; parameters whose gradient only completes
  at the very end of backward (the shared PE parameter that every block feeds, the cls token,
  the patch projection, an absolute table) go into the last bucket;
* ``param.grad`` is a view into its bucket, so the weight-gradient kernels accumulate straight
  into bucket memory - no gather copy before the collective;
* a post-accumulate hook counts a bucket's gradients; when the last one lands the bucket is
  pre-scaled by 1/world and all-reduced asynchronously (NCCL runs it on its own stream, behind an
  event on the compute stream), while backward keeps computing earlier layers;
* ``sync()`` (called before the optimizer step) makes the compute stream wait for the collectives.

Works unchanged on CPU tensors with the ``gloo`` backend (tests/test_dp_gloo.py, world_size 2).
"""
from typing import List

import torch
import torch.distributed as dist


class _Bucket:
    def __init__(self, params, device, dtype):
        self.params = params
        self.numel = sum(p.numel() for p in params)
        self.flat = torch.zeros(self.numel, device=device, dtype=dtype)
        self.pending = len(params)
        self.work = None
        off = 0
        for p in params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()


class BucketedDataParallel(torch.nn.Module):
    def __init__(self, module: torch.nn.Module, bucket_mb: float = 32.0, process_group=None,
                 broadcast_from_rank0: bool = True):
        super().__init__()
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("module has no trainable parameters")
        if broadcast_from_rank0 and self.world > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)
        late_ids = self._late_parameter_ids(module)
        early = [p for p in reversed(params) if id(p) not in late_ids]
        late = [p for p in reversed(params) if id(p) in late_ids]
        cap = int(bucket_mb * (1 << 20) / 4)
        self.buckets: List[_Bucket] = []
        for group in self._split(early, cap) + ([late] if late else []):
            self.buckets.append(_Bucket(group, group[0].device, group[0].dtype))
        self._bucket_of = {}
        for b in self.buckets:
            for p in b.params:
                self._bucket_of[id(p)] = b
                p.register_post_accumulate_grad_hook(self._on_grad)
        self.allreduce_bytes = sum(b.numel for b in self.buckets) * 4

    @staticmethod
    def _late_parameter_ids(module):
        """Parameters whose gradient is complete only when backward reaches the stem."""
        late = set()
        for name, p in module.named_parameters():
            if name.startswith(("pos_embed.", "patch_embed.")) or name == "cls_token":
                late.add(id(p))
        return late

    @staticmethod
    def _split(params, cap):
        groups, cur, size = [], [], 0
        for p in params:
            if cur and (size + p.numel() > cap or p.dtype != cur[0].dtype):
                groups.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += p.numel()
        if cur:
            groups.append(cur)
        return groups

    def _on_grad(self, p):
        b = self._bucket_of[id(p)]
        b.pending -= 1
        if b.pending == 0 and self.world > 1:
            b.flat.mul_(1.0 / self.world)
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def zero_grad(self, set_to_none: bool = False):  # noqa: ARG002 - grads must stay views of the buckets
        for b in self.buckets:
            b.flat.zero_()
            b.pending = len(b.params)
            b.work = None

    def sync(self):
        """Wait for every bucket's all-reduce (call after ``backward``, before ``optimizer.step``)."""
        for b in self.buckets:
            if b.pending != 0 and self.world > 1:
                raise RuntimeError("a bucket never completed: some parameter received no gradient")
            if b.work is not None:
                b.work.wait()
                b.work = None
