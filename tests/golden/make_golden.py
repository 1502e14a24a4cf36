"""Generate the golden fixtures under ``tests/golden/`` from the REFERENCE's own modules.

Runs only in the build container (needs ``/root/reference``):

    python tests/golden/make_golden.py

The reference is imported through ``oracle/ref_shim.py`` (stubs for missing
plotting/serialisation packages, restated ``normflows`` subset).  Every tensor
written here is produced by reference code (``awesome.model.*``,
``awesome.measures.*``, ``torch.optim``), never by the oracle; the oracle and the
CUDA path are then checked against these files.  Sizes are kept tiny so the
fixtures stay small; full-size parity is covered by oracle-vs-CUDA tests.
"""
from __future__ import annotations

import copy
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_shim  # noqa: E402

ref_shim.install()

from awesome.dataset.transformator import Transformator  # noqa: E402
from awesome.measures.fbms_joint_loss import FBMSJointLoss  # noqa: E402
from awesome.measures.miou import MIOU  # noqa: E402
from awesome.measures.se import SE  # noqa: E402
from awesome.measures.unaries_weighted_loss import UnariesWeightedLoss  # noqa: E402
from awesome.measures.weighted_loss import WeightedLoss  # noqa: E402
from awesome.model.convex_diffeomorphism_net import ConvexDiffeomorphismNet  # noqa: E402
from awesome.model.convex_net import ConvexNet, ConvexNextNet  # noqa: E402
from awesome.model.net_factory import real_nvp_path_connected_net  # noqa: E402
from awesome.model.path_connected_net import PathConnectedNet  # noqa: E402
from awesome.run.runner import seed_all  # noqa: E402

torch.set_num_threads(1)          # deterministic reductions


def sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def grads(m):
    return {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
            for k, p in m.named_parameters()}


def notebook_grid(H, W):
    # notebooks/how_to/convexity.ipynb cell 7
    x = torch.arange(0, W)
    y = torch.arange(0, H)
    xx, yy = torch.meshgrid(x, y, indexing="xy")
    grid = torch.stack((xx, yy), dim=0)
    return grid.unsqueeze(0).float() / torch.tensor([W, H]).float().unsqueeze(-1).unsqueeze(-1)


def blob_unaries(H, W, soft, cx=0.55, cy=0.45, rx=0.25, ry=0.3, seed=0):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    sdf = torch.sqrt(((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2) - 1
    if soft:
        return torch.sigmoid((sdf + 0.05 * torch.randn(H, W, generator=g)) / 0.1)
    mask = (sdf > 0).float()
    flip = torch.rand(H, W, generator=g) < 0.05
    return torch.where(flip, 1 - mask, mask)


def gen_icnn_c1():
    """Config 1: convexity how-to (ConvexNextNet L=1, notebook grid, fg/bg-weighted SE, Adam 2e-3)."""
    H, W = 24, 32
    seed_all(0)
    model = ConvexNextNet(n_hidden_layers=1)
    out = {"H": H, "W": W, "init": sd(model)}
    x = notebook_grid(H, W)
    unaries = blob_unaries(H, W, soft=False)
    out["grid"], out["unaries"] = x, unaries
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    crit = SE(reduction="none")
    bu = unaries[None, None]
    bg = bu == 1.0
    fg = ~bg
    fw = torch.tensor(0.4)
    hist = []
    for step in range(6):
        logits = model(x)
        o = torch.sigmoid(logits)
        loss = (1 - fw) * (crit(o[bg], bu[bg]).sum() / bg.sum()) + fw * (crit(o[fg], bu[fg]).sum() / fg.sum())
        opt.zero_grad()
        loss.backward()
        if step == 0:
            out["logits0"], out["loss0"], out["grads0"] = logits.detach().clone(), loss.detach().clone(), grads(model)
        opt.step()
        model.enforce_convexity()
        hist.append(float(loss))
        if step == 0:
            out["after1"] = sd(model)
    out["after6"] = sd(model)
    out["loss_hist"] = torch.tensor(hist)
    out["logits6"] = model(x).detach().clone()
    return out


def gen_icnn_c2():
    """Config 2: ConvexNextNet L=2 on the FBMS grid, UnariesWeightedLoss(SE) (all modes), Adam 1e-3,
    and the Adamax + ReduceLROnPlateau variant of the pretrain loop (path_connected_net.py:929-953)."""
    H, W = 30, 40
    seed_all(42)
    model = ConvexNextNet(n_hidden=130, in_features=2, n_hidden_layers=2)
    out = {"H": H, "W": W, "init": sd(model)}
    x = Transformator.get_positional_matrices(W, H)[None]
    unaries = blob_unaries(H, W, soft=True, seed=1)[None, None]
    out["grid"], out["unaries"] = x, unaries
    for mode in ("none", "sssdms", "ratio", "equal"):
        crit = UnariesWeightedLoss(SE("mean"), mode=mode) if mode != "ratio" else \
            UnariesWeightedLoss(SE("mean"), mode=mode, ratio=0.5)
        model.zero_grad()
        logits = model(x)
        loss = crit(torch.sigmoid(logits), unaries)
        loss.backward()
        out[f"loss_{mode}"], out[f"grads_{mode}"] = loss.detach().clone(), grads(model)
    out["logits0"] = logits.detach().clone()
    # Adam trajectory, MSE
    m2 = copy.deepcopy(model)
    opt = torch.optim.Adam(m2.parameters(), lr=1e-3)
    crit = UnariesWeightedLoss(SE("mean"))
    hist = []
    for step in range(4):
        opt.zero_grad()
        loss = crit(torch.sigmoid(m2(x)), unaries)
        loss.backward()
        opt.step()
        m2.enforce_convexity()
        hist.append(float(loss))
    out["adam_after4"], out["adam_hist"] = sd(m2), torch.tensor(hist)
    # Adamax + plateau (patience shrunk so a reduction actually happens inside the fixture)
    m3 = copy.deepcopy(model)
    opt = torch.optim.Adamax(m3.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, patience=2, factor=0.5, threshold=0.5)
    hist, lrs = [], []
    for step in range(8):
        opt.zero_grad()
        loss = crit(torch.sigmoid(m3(x)), unaries)
        loss.backward()
        opt.step()
        m3.enforce_convexity()
        sched.step(loss)
        hist.append(float(loss))
        lrs.append(opt.param_groups[0]["lr"])
    out["adamax_after8"], out["adamax_hist"], out["adamax_lrs"] = sd(m3), torch.tensor(hist), torch.tensor(lrs)
    out["plateau_args"] = dict(patience=2, factor=0.5, threshold=0.5)
    # older ConvexNet naming
    seed_all(3)
    cn = ConvexNet()
    out["convexnet_init"] = sd(cn)
    out["convexnet_logits"] = cn(x.permute(0, 2, 3, 1).reshape(-1, 2)).detach().clone()
    return out


def gen_pcn(channels, n_flows, T=None):
    """Configs 3/5: RealNVP + ConvexNextNet(L=2) PathConnectedNet from the reference factory."""
    H, W = 20, 28
    seed_all(42)
    model = real_nvp_path_connected_net(channels=channels, hidden_units=32, flow_n_flows=n_flows,
                                        flow_output_fn="tanh", norm="minmax",
                                        convex_net_hidden_units=130, convex_net_hidden_layers=2)
    out = {"H": H, "W": W, "channels": channels, "n_flows": n_flows, "init": sd(model)}
    if channels == 2:
        x = Transformator.get_positional_matrices(W, H)[None]
        unaries = blob_unaries(H, W, soft=True, seed=2)[None, None]
    else:
        x = torch.stack([Transformator.get_positional_matrices(W, H, t=t, t_max=T - 1) for t in range(T)])
        unaries = torch.stack([blob_unaries(H, W, soft=True, seed=2 + t, cx=0.4 + 0.1 * t)[None] for t in range(T)])
    out["grid"], out["unaries"] = x, unaries
    # learn_flow_identity: ActNorm data-dependent init happens on its first forward (train mode)
    seed_all(7)
    hist = model.learn_flow_identity(x, lr=1e-2, weight_decay=1e-5, max_iter=3, device=torch.device("cpu"),
                                     use_progress_bar=False, batch_size=x.shape[0])
    out["identity_hist"], out["after_identity"] = hist.detach().clone(), sd(model)
    model.zero_grad()
    logits = model(x)
    out["logits"] = logits.detach().clone()
    out["deformation"] = model.get_deformation(x).detach().clone()
    out["flow_inverse"] = model.flow_net.inverse(out["deformation"]).detach().clone()
    crit = UnariesWeightedLoss(SE("mean"))
    loss = crit(torch.sigmoid(logits), unaries)
    loss.backward()
    out["loss"], out["grads"] = loss.detach().clone(), grads(model)
    # main fit loop: Adamax groups + plateau (path_connected_net.py:923-953)
    groups = [dict(params=model.flow_net.parameters(), weight_decay=1e-5),
              dict(params=model.convex_net.parameters()), dict(params=model.linear.parameters())]
    opt = torch.optim.Adamax(groups, lr=1e-3)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, patience=200, factor=0.5)
    hist = []
    for step in range(3):
        opt.zero_grad()
        loss = crit(torch.sigmoid(model(x)), unaries)
        loss.backward()
        opt.step()
        model.enforce_convexity()
        sched.step(loss)
        hist.append(float(loss))
    out["fit_after3"], out["fit_hist"] = sd(model), torch.tensor(hist)
    return out


def gen_variants():
    """Round-2 rows: meanstd norm + output_scale, circle prefit, shuffled-batch flow-identity prefit and the
    spatio-temporal main loop with its once-per-epoch scheduler -- all run by the reference's modules."""
    H, W, T = 14, 18, 5
    out = {"H": H, "W": W, "T": T}
    # (1) meanstd + output_scale through the reference factory
    seed_all(21)
    m = real_nvp_path_connected_net(channels=2, hidden_units=16, flow_n_flows=4, flow_output_fn="tanh", flow_output_scale=0.5,
                                    norm="meanstd", spatial_shape=(H, W), convex_net_hidden_units=130, convex_net_hidden_layers=1)
    x = Transformator.get_positional_matrices(W, H)[None]
    un = blob_unaries(H, W, soft=True, seed=4)[None, None]
    g = torch.Generator().manual_seed(3)
    m.train()
    m(x)                                              # ActNorm data-dependent init
    with torch.no_grad():                             # move off the zero-initialised couplings, or scale / tanh are invisible
        for k_, p_ in m.flow_net.named_parameters():
            p_.add_(0.2 * torch.randn(p_.shape, generator=g))
    m.zero_grad()
    y = m(x)
    loss = UnariesWeightedLoss(SE("mean"))(torch.sigmoid(y), un)
    loss.backward()
    out["ms"] = {"state": sd(m), "grid": x, "unaries": un, "logits": y.detach().clone(), "loss": loss.detach().clone(),
                 "grads": grads(m), "deformation": m.get_deformation(x).detach().clone()}
    # (2) circle approximation + learn_convex_net(mode="circle")
    seed_all(22)
    m2 = real_nvp_path_connected_net(channels=2, hidden_units=16, flow_n_flows=4, flow_output_fn="tanh", norm="minmax",
                                     convex_net_hidden_units=130, convex_net_hidden_layers=1)
    hard = blob_unaries(H, W, soft=False, seed=5, cx=0.4, cy=0.6, rx=0.2, ry=0.3)[None, None]
    m2.train()
    m2(x)
    out["circle"] = {"state": sd(m2), "unaries": hard, "circle": m2.get_unary_circle_approximation(1 - hard[0]).clone()}
    hist = m2.learn_convex_net(x, hard, mode="circle", use_deformed_grid=True, lr=1e-3, max_iter=3, device=torch.device("cpu"),
                               use_progress_bar=False)
    out["circle"]["hist"], out["circle"]["after"] = hist.detach().clone(), sd(m2)
    # (3) spatio-temporal: flow identity prefit in shuffled batches, then the main loop (path_connected_net.py:633-719)
    seed_all(23)
    m3 = real_nvp_path_connected_net(channels=3, hidden_units=16, flow_n_flows=6, flow_output_fn="tanh", norm="minmax",
                                     convex_net_hidden_units=130, convex_net_hidden_layers=1)
    xs = torch.stack([Transformator.get_positional_matrices(W, H, t=t, t_max=T - 1) for t in range(T)])
    uns = torch.stack([blob_unaries(H, W, soft=True, seed=30 + t, cx=0.35 + 0.07 * t)[None] for t in range(T)])
    out["st"] = {"init": sd(m3), "grid": xs, "unaries": uns}
    seed_all(24)                                       # the DataLoader's shuffle draws from the global RNG
    ih = m3.learn_flow_identity(PathConnectedNet.create_normalized_grid((T, H, W)), lr=1e-2, weight_decay=1e-5, max_iter=3,
                                device=torch.device("cpu"), use_progress_bar=False, batch_size=2)
    out["st"]["identity_hist"], out["st"]["after_identity"] = ih.detach().clone(), sd(m3)
    groups = [dict(params=m3.flow_net.parameters(), weight_decay=1e-5), dict(params=m3.convex_net.parameters()),
              dict(params=m3.linear.parameters())]
    opt = torch.optim.Adamax(groups, lr=1e-3)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, patience=1, factor=0.5, threshold=0.9)
    crit = UnariesWeightedLoss(SE("mean"))
    bs, step_hist, lrs = 2, [], []
    m3.train()
    for epoch in range(5):
        ep, n_steps = torch.tensor(0.), (T + bs - 1) // bs
        for b0 in range(0, T, bs):
            opt.zero_grad()
            o = torch.sigmoid(m3(xs[b0:b0 + bs]))
            l = crit(o, uns[b0:b0 + bs])
            l.backward()
            opt.step()
            m3.enforce_convexity()
            step_hist.append(float(l))
            ep += (1 / n_steps) * l.item()
        sched.step(ep)
        lrs.append(opt.param_groups[0]["lr"])
    out["st"].update(step_hist=torch.tensor(step_hist), lrs=torch.tensor(lrs), final=sd(m3),
                     plateau_args=dict(patience=1, factor=0.5, threshold=0.9))
    return out


def gen_diffeo():
    """a6: ConvexDiffeomorphismNet (NormalizingFlow1D + ConvexNextNet)."""
    H, W = 16, 20
    seed_all(5)
    model = ConvexDiffeomorphismNet(n_hidden=130, n_hidden_layers=1, nf_layers=4, nf_hidden=70)
    x = Transformator.get_positional_matrices(W, H)[None]
    rows = x.permute(0, 2, 3, 1).reshape(-1, 2)
    with torch.no_grad():
        lin = rows @ model.linear.weight.T + model.linear.bias
        xd = model.diffeo_net(lin)
        y = model(x)
    out = {"H": H, "W": W, "init": sd(model), "grid": x, "lin": lin, "deformed": xd, "logits": y}
    # a trained-like state (every parameter perturbed, in particular the WNScale bias, which is 0 at init):
    # forward, MSE(sigmoid(y), unaries) and all parameter gradients incl. the weight-norm (g, v) pairs
    g = torch.Generator().manual_seed(17)
    with torch.no_grad():
        for p_ in model.parameters():
            p_.add_(0.15 * torch.randn(p_.shape, generator=g))
        model.enforce_convexity()
    un = torch.rand(1, 1, H, W, generator=g)
    y2 = model(x)
    loss = ((torch.sigmoid(y2) - un) ** 2).mean()
    loss.backward()
    with torch.no_grad():
        rows2 = x.permute(0, 2, 3, 1).reshape(-1, 2)
        xd2 = model.diffeo_net(rows2 @ model.linear.weight.T + model.linear.bias)
    out.update({"pert": sd(model), "pert_unaries": un, "pert_deformed": xd2, "pert_logits": y2.detach().clone(),
                "pert_loss": loss.detach().clone(), "pert_grads": grads(model)})
    return out


def gen_star():
    """a16: the star-shape prior class is defined only in a notebook cell; exec that cell's
    source straight from the reference tree (nothing is copied into this repo)."""
    nb = json.load(open(os.path.join(ref_shim.REFERENCE_ROOT,
                                     "notebooks/icml_teaser_code/star_shaped/star.ipynb")))
    src = "".join(nb["cells"][2]["source"])
    ns = {}
    exec("import torch\nimport torch.nn as nn\nimport torch.nn.functional as F\n" + src, ns)
    seed_all(11)
    net = ns["myNet"](150)
    with torch.no_grad():
        net.offset.copy_(torch.tensor([[0.03, -0.02]]))
    g = torch.Generator().manual_seed(3)
    x = torch.rand(257, 2, generator=g) - 0.5
    t = (torch.rand(257, generator=g) > 0.5).float()
    net.offset.requires_grad = True
    y = net(x)
    loss = torch.nn.functional.mse_loss(torch.sigmoid(y).squeeze(), t)
    loss.backward()
    return {"init": sd(net), "x": x, "t": t, "logits": y.detach().clone(), "loss": loss.detach().clone(),
            "grads": grads(net)}


def gen_losses():
    """a10/a15: FBMSJointLoss, WeightedLoss(BCE, sssdms, noneclass=2), MIOU."""
    g = torch.Generator().manual_seed(9)
    H, W = 18, 22
    seg = torch.rand(1, 1, H, W, generator=g).clamp(0.02, 0.98).requires_grad_(True)
    pri = torch.rand(1, 1, H, W, generator=g).clamp(0.02, 0.98).requires_grad_(True)
    target = torch.full((1, 1, H, W), 2.0)
    r = torch.rand(1, 1, H, W, generator=g)
    target[r < 0.08] = 0.0
    target[r > 0.55] = 1.0
    out = {"seg": seg.detach().clone(), "prior": pri.detach().clone(), "target": target}
    for name, scale in (("joint", 1.0), ("joint_clipped", 0.02)):
        crit = FBMSJointLoss(criterion=WeightedLoss(torch.nn.BCELoss(), mode="sssdms", noneclass=2.),
                             penalty_criterion=SE(reduction="mean"), alpha=scale, beta=1.0)
        crit.log = lambda *a, **k: None
        seg.grad = pri.grad = None
        loss = crit(torch.cat([seg, pri], dim=1), target)
        loss.backward()
        out[name] = loss.detach().clone()
        out[name + "_dseg"], out[name + "_dprior"] = seg.grad.clone(), pri.grad.clone()
    miou = MIOU(average="binary", invert=True)
    a = (torch.rand(H, W, generator=g) > 0.4).float()
    b = (torch.rand(H, W, generator=g) > 0.5).float()
    out["miou_a"], out["miou_b"] = a, b
    out["miou_ab"] = miou(a, b).clone()
    out["miou_a_nofg"] = miou(a, torch.ones(H, W)).clone()
    out["miou_aa"] = miou(a, a).clone()
    return out


def main():
    jobs = {
        "icnn_c1.pt": gen_icnn_c1,
        "icnn_c2.pt": gen_icnn_c2,
        "pcn_c3.pt": lambda: gen_pcn(2, 12),
        "pcn_c5.pt": lambda: gen_pcn(3, 18, T=2),
        "variants.pt": gen_variants,
        "diffeo.pt": gen_diffeo,
        "star.pt": gen_star,
        "losses.pt": gen_losses,
    }
    only = set(sys.argv[1:])
    for name, fn in jobs.items():
        if only and name not in only:
            continue
        data = fn()
        path = os.path.join(HERE, name)
        torch.save(data, path)
        print(f"wrote {name}: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
