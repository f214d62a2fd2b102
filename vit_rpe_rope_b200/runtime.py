"""Training-step runner: the whole step (zero_grad -> forward -> loss -> backward -> all-reduce ->
AdamW) captured once into a CUDA graph and replayed.

The reference's step (train.py:108-116) is ~900 kernel launches for ViT-B/16; issued eagerly from
Python they leave the GPU idle ~12 % of the step.  B200-first: no tracing compiler - the eager step,
which already runs the hand-written kernels through the C ABI on the current stream, is captured with
``torch.cuda.graph`` (kernel parameters incl. the TMA tensor maps are baked into the graph; every
buffer lives in the graph's private pool, the gradient buckets of ``BucketedDataParallel`` are
persistent) and replayed with one launch.  Inputs are copied into static device buffers before each
replay; the loss is read from a static tensor.

Single-process only: capturing the NCCL all-reduce of ``BucketedDataParallel`` in the graph hung at
2, 4 and 8 ranks in round-1 testing (undiagnosed), so callers pass ``use_graph=False`` when the world
size is > 1 and the step is issued eagerly.
"""
import torch
import torch.nn.functional as F

from . import _lib
from .dp import BucketedDataParallel


class GraphedTrainStep:
    def __init__(self, dp: BucketedDataParallel, optimizer: torch.optim.Optimizer, images_shape, num_classes,
                 device, bf16: bool = True, use_graph: bool = True, warmup: int = 3):
        self.dp, self.opt, self.bf16, self.device = dp, optimizer, bf16, device
        self.static_images = torch.zeros(images_shape, device=device, dtype=torch.float32)
        self.static_labels = torch.zeros(images_shape[0], device=device, dtype=torch.long)
        self.static_loss = torch.zeros((), device=device, dtype=torch.float32)
        self.num_classes = num_classes
        self.graph = None
        self.launches_per_step = 0
        self._warmup = warmup
        self._use_graph = use_graph

    def _eager_step(self):
        self.dp.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.bf16):
            logits = self.dp(self.static_images)
            loss = F.cross_entropy(logits.float(), self.static_labels)
        loss.backward()
        self.dp.sync()
        self.opt.step()
        self.static_loss.copy_(loss.detach())

    def prepare(self, images, labels):
        """Warm up (allocator, cuBLAS workspaces, one-time kernel attributes) and capture."""
        self.load(images, labels)
        if not self._use_graph:
            # eager mode (always the case for world size > 1): warm up on the current stream, exactly
            # the call sequence of a plain training loop
            for _ in range(self._warmup):
                before = _lib.launch_count()
                self._eager_step()
                self.launches_per_step = _lib.launch_count() - before
            torch.cuda.synchronize()
            return
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                before = _lib.launch_count()
                self._eager_step()
                self.launches_per_step = _lib.launch_count() - before
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if self._use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._eager_step()
            torch.cuda.synchronize()

    def load(self, images, labels):
        self.static_images.copy_(images, non_blocking=True)
        self.static_labels.copy_(labels, non_blocking=True)

    def run(self):
        """One training step on whatever is in the static input buffers; returns the (device) loss."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._eager_step()
        return self.static_loss
