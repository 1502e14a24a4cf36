"""Per-kernel-class CUDA-event timing of one fit configuration (GPU box).
    python scripts/prof_classes.py [icnn|flow|flow3|diffeo] [precision]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import awesome_b200 as A
from awesome_b200 import _lib

kind = sys.argv[1] if len(sys.argv) > 1 else "flow"
prec = sys.argv[2] if len(sys.argv) > 2 else "f16"
H, W = 480, 640
torch.manual_seed(0)
yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
un = torch.sigmoid((torch.sqrt(((xx - 0.52) / 0.27) ** 2 + ((yy - 0.47) / 0.31) ** 2) - 1) / 0.08).cuda()
if kind == "icnn":
    m = A.ConvexNextNet(n_hidden_layers=2, precision=prec).cuda()
    grid = A.GridSpecHost("linspace", 1, H, W)
elif kind == "flow":
    m = A.real_nvp_path_connected_net(channels=2, hidden_units=32, flow_n_flows=12, flow_output_fn="tanh", precision=prec).cuda()
    grid = A.GridSpecHost("linspace", 1, H, W)
elif kind == "flow3":
    m = A.real_nvp_path_connected_net(channels=3, hidden_units=32, flow_n_flows=18, flow_output_fn="tanh", precision=prec).cuda()
    grid = A.GridSpecHost("linspace", 1, H, W, t0=0.3)
else:
    m = A.ConvexDiffeomorphismNet(n_hidden_layers=1, precision=prec).cuda()
    grid = A.GridSpecHost("linspace", 1, H, W)
f = m.make_fitter(grid, un, A.LossConfig("mse"), A.OptimConfig("adamax", lr=1e-3, weight_decay=[1e-5, 0, 0, 0]), use_graph=False)
f.run(5)
lib = _lib.load()
n = lib.awb_profile_classes()
lib.awb_profile_enable(1)
f.run(6)
tot, cnt = (C.c_double * n)(), (C.c_int32 * n)()
_lib.check(lib.awb_profile_read(tot, cnt))
lib.awb_profile_enable(0)
for i in range(n):
    if cnt[i]:
        print(f"{lib.awb_profile_class_name(i).decode():14s} {tot[i] / 6:8.4f} ms/step  ({cnt[i] / 6:.1f} launches/step)")
