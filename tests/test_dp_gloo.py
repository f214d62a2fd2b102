"""world_size-2 CPU (gloo) test of the bucketed data-parallel wrapper: after one
backward + sync every rank holds the mean gradient of the two ranks' batches, identical to a
single-process run on the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vit_rpe_rope_b200.dp import BucketedDataParallel


class TinyNet(torch.nn.Module):
    """Parameter names mimic the ViT so the 'late bucket' rule (pos_embed / cls_token / patch_embed) is exercised."""

    def __init__(self):
        super().__init__()
        self.patch_embed = torch.nn.Linear(12, 16)
        self.cls_token = torch.nn.Parameter(torch.zeros(1, 16))
        self.pos_embed = torch.nn.Linear(16, 16, bias=False)  # shared by both blocks, like the shared PE module
        self.blocks = torch.nn.ModuleList([torch.nn.Linear(16, 16) for _ in range(3)])
        self.head = torch.nn.Linear(16, 4)

    def forward(self, x):
        x = self.patch_embed(x) + self.cls_token
        for blk in self.blocks:
            x = torch.tanh(blk(x) + self.pos_embed(x))
        return self.head(x)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, style="dp_zero_grad"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)  # different initial weights per rank: rank 0's must win
        net = TinyNet()
        dp = BucketedDataParallel(net, bucket_mb=0.001)  # tiny cap -> several buckets
        assert len(dp.buckets) >= 3
        late = dp.buckets[-1].params
        assert any(p is net.cls_token for p in late) and any(p is net.pos_embed.weight for p in late)
        g = torch.Generator().manual_seed(7)
        x = torch.randn(8, 12, generator=g)
        y = torch.randint(0, 4, (8,), generator=g)
        shard = slice(rank * 4, rank * 4 + 4)
        opt = torch.optim.SGD(net.parameters(), lr=0.0)  # lr 0: weights stay put, only the zeroing style matters
        for step in range(3):  # later steps check zeroing / re-arming of the buckets
            if style == "dp_zero_grad":
                dp.zero_grad()           # default: gradients dropped, autograd's tensors adopted, one gather per bucket
            elif style == "dp_zero_in_place":
                dp.zero_grad(set_to_none=False)  # gradients stay views of the zeroed buckets
            else:  # the reference loop (train.py:111): set_to_none=True unbinds every p.grad from its bucket
                opt.zero_grad()
            if style == "accumulate":  # two half-shards: the first backward only accumulates locally
                a, b_ = slice(rank * 4, rank * 4 + 2), slice(rank * 4 + 2, rank * 4 + 4)
                with dp.no_sync():
                    (torch.nn.functional.cross_entropy(dp(x[a]), y[a]) * 0.5).backward()
                (torch.nn.functional.cross_entropy(dp(x[b_]), y[b_]) * 0.5).backward()
            else:
                loss = torch.nn.functional.cross_entropy(dp(x[shard]), y[shard])
                loss.backward()
            dp.sync()
            opt.step()
        for b in dp.buckets:  # whatever the zeroing style, the gradients live in the flat buckets again
            for p in b.params:
                assert p.grad.data_ptr() == b.views[id(p)].data_ptr()
        for p in net.parameters():
            assert p.grad is not None and p.grad.data_ptr() != 0
        out[rank] = ({k: v.detach().clone() for k, v in net.state_dict().items()},
                     {n: p.grad.detach().clone() for n, p in net.named_parameters()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("style", ["dp_zero_grad", "dp_zero_in_place", "optimizer_zero_grad", "accumulate"])
def test_bucketed_dp_matches_single_process(style):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out, style), nprocs=world, join=True)
    (sd0, g0), (sd1, g1) = out[0], out[1]
    for k in sd0:  # broadcast from rank 0 at construction
        assert torch.equal(sd0[k], sd1[k]), k
    for n in g0:   # all-reduced: identical on both ranks
        assert torch.allclose(g0[n], g1[n], atol=1e-7), n
    ref = TinyNet()
    ref.load_state_dict(sd0)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 12, generator=g)
    y = torch.randint(0, 4, (8,), generator=g)
    torch.nn.functional.cross_entropy(ref(x), y).backward()
    for n, p in ref.named_parameters():
        assert torch.allclose(g0[n], p.grad, atol=1e-6), n


def test_single_process_wrapper_is_transparent():
    net = TinyNet()
    dp = BucketedDataParallel(net, bucket_mb=1.0)
    x = torch.randn(4, 12)
    dp.zero_grad()
    dp(x).sum().backward()
    dp.sync()
    ref = TinyNet()
    ref.load_state_dict(net.state_dict())
    ref(x).sum().backward()
    for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-6), n
    with pytest.raises(ValueError):
        BucketedDataParallel(torch.nn.ReLU())


def test_single_process_optimizer_zero_grad_rebinds():
    """optimizer.zero_grad() (set_to_none=True) between steps: gradients are re-bound to the buckets and
    match a plain module, step after step."""
    net = TinyNet()
    ref = TinyNet()
    ref.load_state_dict(net.state_dict())
    dp = BucketedDataParallel(net, bucket_mb=0.001)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    for step in range(3):
        x = torch.randn(4, 12, generator=torch.Generator().manual_seed(step))
        opt.zero_grad()
        ropt.zero_grad()
        dp(x).sum().backward()
        dp.sync()
        ref(x).sum().backward()
        for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            assert torch.allclose(p.grad, q.grad, atol=1e-6), (step, n)
        opt.step()
        ropt.step()
