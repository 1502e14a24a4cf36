"""Multi-GPU check of the joint UNet + prior step (BASELINE config 5): run under torchrun, one rank per GPU.
Asserts that the replicas stay bit-identical after the NCCL gradient all-reduce and prints the step time.
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/joint_dp_check.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import awesome_b200 as A
from awesome_b200 import measures as M

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
H, W, B, T = 480, 640, 2, 200                       # 2 frames per GPU per step
torch.manual_seed(42)


def block(i, o):
    return torch.nn.Sequential(torch.nn.Conv2d(i, o, 3, padding=1), torch.nn.BatchNorm2d(o), torch.nn.ReLU(),
                               torch.nn.Conv2d(o, o, 3, padding=1), torch.nn.BatchNorm2d(o), torch.nn.ReLU())


class MiniUNet(torch.nn.Module):
    """Stock-PyTorch stand-in for the reference UNet(in_chn=4) (out of scope): 3 levels."""

    def __init__(self, cin=4, w=32):
        super().__init__()
        self.d1, self.d2, self.d3 = block(cin, w), block(w, 2 * w), block(2 * w, 4 * w)
        self.u2, self.u1 = block(6 * w, 2 * w), block(3 * w, w)
        self.out = torch.nn.Conv2d(w, 1, 1)

    def forward(self, x):
        a = self.d1(x)
        b = self.d2(torch.nn.functional.max_pool2d(a, 2))
        c = self.d3(torch.nn.functional.max_pool2d(b, 2))
        b = self.u2(torch.cat([torch.nn.functional.interpolate(c, scale_factor=2.0), b], 1))
        a = self.u1(torch.cat([torch.nn.functional.interpolate(b, scale_factor=2.0), a], 1))
        return self.out(a)


seg = MiniUNet().to(dev)
pri = A.real_nvp_path_connected_net(channels=3, hidden_units=32, flow_n_flows=18, flow_output_fn="tanh",
                                    convex_net_hidden_layers=2, precision=os.environ.get("AWB_PRECISION", "f16")).to(dev)
g = torch.Generator().manual_seed(1000 + rank)
img = torch.randn(B, 4, H, W, generator=g).to(dev)
f0 = rank * B
grid = A.GridSpecHost("linspace", B, H, W, t0=f0 / (T - 1), t_step=1.0 / (T - 1)).materialize(3, dev)
lab = torch.full((B, 1, H, W), 2.0)
lab[torch.rand(B, 1, H, W, generator=g) < 0.05] = 0.0
lab[torch.rand(B, 1, H, W, generator=g) < 0.1] = 1.0
lab = lab.to(dev)
pri(grid)                                            # ActNorm init (rank-local data), then replicate rank 0
tr = A.JointTrainer(seg, pri, M.FBMSJointLoss(), optimizer_args=dict(lr=1e-4))
tr.broadcast_parameters(0)
for _ in range(3):
    tr.step(img, grid, lab)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
steps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = tr.step(img, grid, lab)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
flat = torch.cat([p.detach().reshape(-1) for p in list(seg.parameters()) + list(pri.parameters())])
ok = True
if world > 1:
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([float(torch.equal(ref, flat))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    ok = bool(same.item() > 0)
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
if rank == 0:
    print(f"[joint_dp_check] world={world} frames/step={world * B} ms/step={ms:.2f} frames/s={world * B / ms * 1e3:.1f} "
          f"bucket={tr.bucket.nbytes / 1e6:.2f} MB replicas_identical={ok} loss={float(loss):.5f}", flush=True)
assert ok, "replicas diverged"
if world > 1:
    dist.destroy_process_group()
