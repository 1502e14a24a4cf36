"""CPU tests of the joint data-parallel step (world_size 2, gloo) and of the joint loss against the oracle."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_fbms_joint_loss_matches_oracle():
    from awesome_b200 import measures as M
    from oracle import prior_oracle as O
    g = torch.Generator().manual_seed(0)
    for clip_case in (0.05, 0.9):            # penalty below / above the segmentation loss
        seg = torch.rand(2, 1, 24, 32, generator=g) * 0.9 + 0.05
        pri = (seg + clip_case * torch.randn(2, 1, 24, 32, generator=g)).clamp(0.01, 0.99)
        tgt = torch.full((2, 1, 24, 32), 2.0)
        tgt[torch.rand(2, 1, 24, 32, generator=g) < 0.1] = 0.0
        tgt[torch.rand(2, 1, 24, 32, generator=g) < 0.3] = 1.0
        seg.requires_grad_(True)
        pri.requires_grad_(True)
        ours = M.FBMSJointLoss()(torch.cat([seg, pri], 1), tgt)
        gs, gp = torch.autograd.grad(ours, [seg, pri])
        s2, p2 = seg.detach().clone().requires_grad_(True), pri.detach().clone().requires_grad_(True)
        ref = O.loss_fbms_joint(s2, p2, tgt)
        rs, rp = torch.autograd.grad(ref, [s2, p2])
        torch.testing.assert_close(ours, ref)
        torch.testing.assert_close(gs, rs)
        torch.testing.assert_close(gp, rp)


class _Seg(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.c = torch.nn.Conv2d(3, 1, 3, padding=1)

    def forward(self, x):
        return self.c(x)


class _PriorStandIn(torch.nn.Module):
    """CPU stand-in with the prior interface (the native prior is CUDA-only): per-pixel MLP on the grid."""

    def __init__(self):
        super().__init__()
        self.a = torch.nn.Conv2d(3, 8, 1)
        self.b = torch.nn.Conv2d(8, 1, 1)

    def forward(self, g):
        return self.b(torch.relu(self.a(g)))

    def enforce_convexity(self):
        with torch.no_grad():
            self.b.weight.clamp_(min=0)


def _mse_joint(out, labels):
    half = out.shape[1] // 2
    return ((out[:, :half] - labels) ** 2).mean() + ((out[:, half:] - out[:, :half]) ** 2).mean()


def _make(seed=0):
    torch.manual_seed(seed)
    return _Seg(), _PriorStandIn()


def _data(n, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n, 3, 12, 16, generator=g), torch.rand(n, 3, 12, 16, generator=g),
            (torch.rand(n, 1, 12, 16, generator=g) > 0.5).float())


def _worker(rank, world, port, ret, comm="overlap"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import awesome_b200 as A
        seg, pri = _make(seed=100 + rank)           # replicas start different: broadcast must fix that
        tr = A.JointTrainer(seg, pri, _mse_joint, optimizer_cls=torch.optim.Adam, optimizer_args=dict(lr=1e-2), n_buckets=3, comm=comm)
        tr.broadcast_parameters(0)
        img, grid, lab = _data(4, seed=7)
        sl = slice(rank * 2, rank * 2 + 2)           # 2 frames per rank
        losses = [float(tr.step(img[sl], grid[sl], lab[sl])) for _ in range(3)]
        flat = torch.cat([p.detach().reshape(-1) for p in list(seg.parameters()) + list(pri.parameters())])
        ret[rank] = (flat, losses, tr.bucket.nbytes, tr.bucket.n_buckets)
    finally:
        dist.destroy_process_group()


def test_joint_step_two_ranks_equals_single_process_on_concatenated_batch():
    import awesome_b200 as A
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    flat0, _, nbytes, nb = ret[0]
    flat1 = ret[1][0]
    assert nb == 3                                    # bucketed: three collectives launched from the backward hooks
    assert torch.equal(flat0, flat1), "replicas diverged"
    # single process, whole batch, same initial weights as rank 0
    seg, pri = _make(seed=100)
    tr = A.JointTrainer(seg, pri, _mse_joint, optimizer_cls=torch.optim.Adam, optimizer_args=dict(lr=1e-2))
    img, grid, lab = _data(4, seed=7)
    for _ in range(3):
        tr.step(img, grid, lab)
    ref = torch.cat([p.detach().reshape(-1) for p in list(seg.parameters()) + list(pri.parameters())])
    torch.testing.assert_close(flat0, ref, rtol=1e-5, atol=1e-6)
    assert nbytes == 4 * ref.numel()
    assert float(pri.b.weight.min()) >= 0.0          # clamp applied for non-fused optimizers


def test_grad_bucket_views_survive_zero_grad():
    import awesome_b200 as A
    seg, pri = _make()
    b = A.GradBucket(list(seg.parameters()) + list(pri.parameters()))
    img, grid, lab = _data(2, 1)
    _mse_joint(torch.cat([torch.sigmoid(seg(img)), torch.sigmoid(pri(grid))], 1), lab).backward()
    assert float(b.flat.abs().sum()) > 0
    for p in seg.parameters():
        p.grad = None                                 # e.g. zero_grad(set_to_none=True) by foreign code
    b.zero()
    assert float(b.flat.abs().sum()) == 0 and all(p.grad is not None for p in seg.parameters())


def test_overlapped_buckets_equal_one_allreduce_after_backward():
    """comm="overlap" (bucket collectives launched from autograd hooks during backward) and comm="after" (one collective
    after backward) are the same arithmetic: bit-identical replicas and bit-identical results between the two modes."""
    world = 2
    mgr = mp.Manager()
    res = {}
    for i, comm in enumerate(("overlap", "after")):
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, 29700 + (os.getpid() % 1000) + 7 * i, ret, comm), nprocs=world, join=True)
        assert torch.equal(ret[0][0], ret[1][0])
        res[comm] = ret[0]
    assert torch.equal(res["overlap"][0], res["after"][0]) and res["overlap"][1] == res["after"][1]
    assert res["overlap"][3] == 3 and res["after"][3] == 1
