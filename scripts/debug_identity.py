"""Debug helper (GPU box): one learn_flow_identity step, CUDA vs oracle, per-key max difference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import awesome_b200 as A
from oracle import prior_oracle as O

g = torch.load("tests/golden/pcn_c3.pt", weights_only=False)
m = A.real_nvp_path_connected_net(channels=2, hidden_units=32, flow_n_flows=12, flow_output_fn="tanh",
                                  convex_net_hidden_units=130, convex_net_hidden_layers=2)
m.load_state_dict(g["init"]); m = m.to("cuda")
grid = g["grid"].cuda()
m.actnorm_init(A.GridSpecHost.from_tensor(grid), use_linear=False)
sd0 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
# oracle: same init
p = O.clone_params(g["init"])
mn, mx, nmn, nmx = (p["flow_net.norm." + k] for k in ("min", "max", "new_min", "new_max"))
B, C, H, W = g["grid"].shape
fkeys = [k for k in p if k.startswith("flow_net.") and p[k].dtype.is_floating_point
         and not k.endswith(("data_dep_init_done", ".b")) and ".norm." not in k]
for k in fkeys: p[k].requires_grad_(True)
z = O.flow_forward(p, O.pixelize(O.minmax(g["grid"], mn, mx, nmn, nmx)), O.FLOW_PREFIX, actnorm_init=True)
out = O.minmax(O.unpixelize(z, B, H, W), nmn, nmx, mn, mx)
loss = ((g["grid"] - out) ** 2).mean()
grads = dict(zip(fkeys, torch.autograd.grad(loss, [p[k] for k in fkeys])))
print("actnorm init diff:")
for k in fkeys:
    if k.endswith((".s", ".t")) and "net" not in k.split("flows.")[1]:
        d = (sd0[k] - p[k].detach()).abs().max()
        if d > 1e-5: print("  ", k, float(d), sd0[k].tolist(), p[k].detach().tolist())
print("oracle loss0", float(loss))
fit = A.FlowIdentityFitter(m._prior_for(grid.device), m._ensure_flat(), A.GridSpecHost.from_tensor(grid),
                           A.OptimConfig("adamax", lr=1e-2, weight_decay=[1e-5, 0, 0, 0]), use_graph=False)
h = fit.run(1); print("cuda loss0", float(h.cpu()))
sd1 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
ms = {k: torch.zeros_like(p[k]) for k in fkeys}; us = {k: torch.zeros_like(p[k]) for k in fkeys}
for k in fkeys:
    p[k].requires_grad_(False)
    O.adamax_step(p[k], grads[k], ms[k], us[k], 1, 1e-2, weight_decay=1e-5)
bad = 0
for k in fkeys:
    d = (sd1[k] - p[k]).abs()
    if float(d.max()) > 1e-4:
        bad += 1
        i = int(d.reshape(-1).argmax())
        print(f"  {k}: maxdiff {float(d.max()):.4g} at {i}; ours step {float((sd1[k]-sd0[k]).reshape(-1)[i]):.4g} "
              f"oracle grad {float(grads[k].reshape(-1)[i]):.4g} n_bad {(d>1e-4).sum().item()}/{d.numel()}")
print("keys with mismatch:", bad, "of", len(fkeys))
for k in sd1:
    if not k.startswith("flow_net.net") and not torch.equal(sd1[k], sd0[k]): print("  changed non-flow key:", k)
