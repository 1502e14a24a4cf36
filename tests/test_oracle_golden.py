"""Pins oracle/prior_oracle.py against fixtures generated from the reference's own modules
(tests/golden/make_golden.py).  CPU only."""
import copy

import pytest
import torch

from oracle import prior_oracle as O

torch.set_num_threads(1)
TOL = dict(rtol=2e-5, atol=2e-6)


def close(a, b, **kw):
    tol = dict(TOL)
    tol.update(kw)
    torch.testing.assert_close(a, b, **tol)


def rows(grid):
    return O.pixelize(grid)


def test_grid_notebook_and_linspace(golden):
    g1 = golden("icnn_c1.pt")
    close(O.grid_index(g1["H"], g1["W"]), g1["grid"], rtol=0, atol=0)
    g2 = golden("icnn_c2.pt")
    close(O.grid_linspace(g2["H"], g2["W"])[None], g2["grid"], rtol=0, atol=0)
    g5 = golden("pcn_c5.pt")
    T = g5["grid"].shape[0]
    g = torch.stack([O.grid_linspace(g5["H"], g5["W"], t=t, t_max=T - 1) for t in range(T)])
    close(g, g5["grid"], rtol=0, atol=0)


def test_icnn_c1_forward_loss_grads(golden):
    g = golden("icnn_c1.pt")
    p = O.clone_params(g["init"], requires_grad=True)
    y = O.icnn_forward(p, rows(g["grid"]))
    close(O.unpixelize(y, 1, g["H"], g["W"]), g["logits0"])
    loss = O.loss_fgbg_se(y, g["unaries"], 0.4)
    close(loss, g["loss0"])
    keys = O.icnn_param_keys(p)
    grads = torch.autograd.grad(loss, [p[k] for k in keys])
    for k, gr in zip(keys, grads):
        close(gr, g["grads0"][k], atol=1e-7)


def test_icnn_c1_fit_trajectory(golden):
    g = golden("icnn_c1.pt")
    p = O.clone_params(g["init"])
    rec = []
    O.fit_icnn(p, rows(g["grid"]), g["unaries"], steps=1, loss=O.LOSS_FGBG_SE, optimizer="adam", lr=2e-3,
               fg_weight=0.4, record=rec)
    for k in g["after1"]:
        close(p[k], g["after1"][k], atol=1e-6)
    p = O.clone_params(g["init"])
    rec = []
    O.fit_icnn(p, rows(g["grid"]), g["unaries"], steps=6, loss=O.LOSS_FGBG_SE, optimizer="adam", lr=2e-3,
               fg_weight=0.4, record=rec)
    close(torch.tensor(rec), g["loss_hist"], rtol=1e-4)
    for k in g["after6"]:
        close(p[k], g["after6"][k], rtol=1e-3, atol=2e-5)
    close(O.unpixelize(O.icnn_forward(p, rows(g["grid"])), 1, g["H"], g["W"]), g["logits6"], rtol=1e-3, atol=1e-4)
    # clamp really applied
    for k in O.icnn_clamp_keys(p):
        assert float(p[k].min()) >= 0.0


@pytest.mark.parametrize("mode", ["none", "sssdms", "ratio", "equal"])
def test_icnn_c2_weighted_losses(golden, mode):
    g = golden("icnn_c2.pt")
    p = O.clone_params(g["init"], requires_grad=True)
    y = O.icnn_forward(p, rows(g["grid"]))
    close(O.unpixelize(y, 1, g["H"], g["W"]), g["logits0"])
    loss = O.loss_unaries_weighted_se(y, g["unaries"], mode, ratio=0.5)
    close(loss, g[f"loss_{mode}"])
    keys = O.icnn_param_keys(p)
    grads = torch.autograd.grad(loss, [p[k] for k in keys])
    for k, gr in zip(keys, grads):
        close(gr, g[f"grads_{mode}"][k], atol=1e-7)


def test_icnn_c2_adam_and_adamax_plateau(golden):
    g = golden("icnn_c2.pt")
    p = O.clone_params(g["init"])
    rec = []
    O.fit_icnn(p, rows(g["grid"]), g["unaries"], steps=4, optimizer="adam", lr=1e-3, record=rec)
    close(torch.tensor(rec), g["adam_hist"], rtol=1e-4)
    for k in g["adam_after4"]:
        close(p[k], g["adam_after4"][k], rtol=1e-3, atol=2e-5)
    # Adamax + plateau with the fixture's shrunk patience
    p = O.clone_params(g["init"])
    keys = O.icnn_param_keys(p)
    ms = {k: torch.zeros_like(p[k]) for k in keys}
    us = {k: torch.zeros_like(p[k]) for k in keys}
    sched = O.Plateau([1e-3], **g["plateau_args"])
    hist, lrs = [], []
    for step in range(1, 9):
        for k in keys:
            p[k].requires_grad_(True)
        loss = O.loss_unaries_weighted_se(O.icnn_forward(p, rows(g["grid"])), g["unaries"])
        grads = torch.autograd.grad(loss, [p[k] for k in keys])
        for k, gr in zip(keys, grads):
            p[k].requires_grad_(False)
            O.adamax_step(p[k], gr, ms[k], us[k], step, sched.lrs[0])
        O.icnn_enforce_convexity(p)
        sched.step(float(loss))
        hist.append(float(loss))
        lrs.append(sched.lrs[0])
    close(torch.tensor(hist), g["adamax_hist"], rtol=1e-4)
    close(torch.tensor(lrs), g["adamax_lrs"].float(), rtol=1e-6)
    assert lrs[-1] < 1e-3, "fixture must exercise a plateau reduction"
    for k in g["adamax_after8"]:
        close(p[k], g["adamax_after8"][k], rtol=1e-3, atol=2e-5)


def test_convexnet_old_keys(golden):
    g = golden("icnn_c2.pt")
    close(O.convexnet_forward(g["convexnet_init"], rows(g["grid"])), g["convexnet_logits"])


@pytest.mark.parametrize("name", ["pcn_c3.pt", "pcn_c5.pt"])
def test_pathconnected_masks_forward_inverse_grads(golden, name):
    g = golden(name)
    C, F = g["channels"], g["n_flows"]
    masks = O.realnvp_masks(C, F)
    for f in range(F):
        assert torch.equal(g["init"][f"{O.FLOW_PREFIX}flows.{2 * f}.b"].reshape(-1), masks[f])
    p = O.clone_params(g["after_identity"], requires_grad=True)
    B, _, H, W = g["grid"].shape
    xd = O.pathconnected_deformation(p, g["grid"])
    close(O.unpixelize(xd, B, H, W), g["deformation"], atol=1e-5)
    y = O.pathconnected_forward(p, g["grid"])
    close(y, g["logits"], atol=1e-5)
    # inverse of the NormNet(flow): minmax -> inverse flow -> inverse minmax
    mn, mx, nmn, nmx = (p["flow_net.norm." + k] for k in ("min", "max", "new_min", "new_max"))
    with torch.no_grad():
        zi = O.flow_inverse(p, O.pixelize(O.minmax(g["deformation"], mn, mx, nmn, nmx)), O.FLOW_PREFIX)
        zi = O.minmax(O.unpixelize(zi, B, H, W), nmn, nmx, mn, mx)
    close(zi, g["flow_inverse"], atol=1e-5)
    loss = O.loss_unaries_weighted_se(O.pixelize(y), g["unaries"])
    close(loss, g["loss"])
    keys = [k for k in g["grads"]]
    grads = torch.autograd.grad(loss, [p[k] for k in keys], allow_unused=True)
    for k, gr in zip(keys, grads):
        gr = torch.zeros_like(p[k]) if gr is None else gr
        close(gr, g["grads"][k], rtol=1e-3, atol=2e-7)


def test_pathconnected_actnorm_init_and_identity_fit(golden):
    """First forward in learn_flow_identity performs the ActNorm data-dependent init; three Adamax
    steps with SE against the input grid follow (path_connected_net.py:155-250)."""
    g = golden("pcn_c3.pt")
    p = O.clone_params(g["init"])
    fkeys = [k for k in p if k.startswith("flow_net.") and p[k].dtype.is_floating_point
             and not k.endswith(("data_dep_init_done", ".b")) and ".norm." not in k]
    ms = {k: torch.zeros_like(p[k]) for k in fkeys}
    us = {k: torch.zeros_like(p[k]) for k in fkeys}
    B, C, H, W = g["grid"].shape
    mn, mx, nmn, nmx = (p["flow_net.norm." + k] for k in ("min", "max", "new_min", "new_max"))
    hist = []
    for step in range(1, 4):
        for k in fkeys:
            p[k].requires_grad_(True)
        z = O.flow_forward(p, O.pixelize(O.minmax(g["grid"], mn, mx, nmn, nmx)), O.FLOW_PREFIX,
                           actnorm_init=True)
        out = O.minmax(O.unpixelize(z, B, H, W), nmn, nmx, mn, mx)
        loss = ((g["grid"] - out) ** 2).mean()
        grads = torch.autograd.grad(loss, [p[k] for k in fkeys])
        for k, gr in zip(fkeys, grads):
            p[k].requires_grad_(False)
            O.adamax_step(p[k], gr, ms[k], us[k], step, 1e-2, weight_decay=1e-5)
        hist.append(float(loss))
    close(torch.tensor(hist), g["identity_hist"], rtol=1e-4, atol=1e-7)
    for k in fkeys:
        close(p[k], g["after_identity"][k], rtol=1e-3, atol=2e-5)
    for f in range(g["n_flows"]):
        assert float(p[f"{O.FLOW_PREFIX}flows.{2 * f + 1}.data_dep_init_done"]) == 1.0


def test_diffeo_flow1d(golden):
    g = golden("diffeo.pt")
    p = g["init"]
    rows_ = rows(g["grid"])
    lin = rows_ @ p["linear.weight"].T + p["linear.bias"]
    close(lin, g["lin"])
    xd = O.flow1d_forward(p, lin)
    close(xd, g["deformed"], atol=1e-5)
    y = O.icnn_forward(p, xd, "convex_net.")
    close(O.unpixelize(y, 1, g["H"], g["W"]), g["logits"], atol=1e-5)
    # perturbed ("trained-like") state: non-zero WNScale bias, weight-norm gradients
    p = O.clone_params(g["pert"], requires_grad=True)
    lin = rows_ @ p["linear.weight"].T + p["linear.bias"]
    xd = O.flow1d_forward(p, lin)
    close(xd.detach(), g["pert_deformed"], atol=1e-5)
    y = O.unpixelize(O.icnn_forward(p, xd, "convex_net."), 1, g["H"], g["W"])
    close(y.detach(), g["pert_logits"], atol=2e-5)
    loss = ((torch.sigmoid(y) - g["pert_unaries"]) ** 2).mean()
    close(loss.detach(), g["pert_loss"])
    keys = list(g["pert_grads"])
    grads = torch.autograd.grad(loss, [p[k] for k in keys], allow_unused=True)
    for k, gr in zip(keys, grads):
        gr = torch.zeros_like(p[k]) if gr is None else gr
        close(gr, g["pert_grads"][k], rtol=2e-3, atol=1e-7)


def test_star(golden):
    g = golden("star.pt")
    p = O.clone_params(g["init"], requires_grad=True)
    y = O.star_forward(p, g["x"])
    close(y, g["logits"], atol=1e-5)
    loss = torch.nn.functional.mse_loss(torch.sigmoid(y).squeeze(), g["t"])
    close(loss, g["loss"])
    keys = list(g["grads"])
    grads = torch.autograd.grad(loss, [p[k] for k in keys])
    for k, gr in zip(keys, grads):
        close(gr, g["grads"][k], rtol=1e-3, atol=1e-7)


def test_joint_loss_and_miou(golden):
    g = golden("losses.pt")
    for name, alpha in (("joint", 1.0), ("joint_clipped", 0.02)):
        seg = g["seg"].clone().requires_grad_(True)
        pri = g["prior"].clone().requires_grad_(True)
        loss = O.loss_fbms_joint(seg, pri, g["target"], alpha=alpha, beta=1.0)
        close(loss, g[name])
        ds, dp = torch.autograd.grad(loss, [seg, pri])
        close(ds, g[name + "_dseg"], atol=1e-8)
        close(dp, g[name + "_dprior"], atol=1e-8)
    assert abs(O.miou_binary_inverted(g["miou_a"], g["miou_b"]) - float(g["miou_ab"])) < 1e-6
    assert O.miou_binary_inverted(g["miou_a"], torch.ones_like(g["miou_a"])) == float(g["miou_a_nofg"]) == 0.0
    assert abs(O.miou_binary_inverted(g["miou_a"], g["miou_a"]) - float(g["miou_aa"])) < 1e-6
