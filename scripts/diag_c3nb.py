"""Diagnostic: c3nb full-schedule fit on the fp32 and the tensor path, loss curves and mask agreement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import awesome_b200 as A
from awesome_b200 import synth
g = torch.load("tests/golden/full_c3nb.pt", weights_only=False)
H, W = g["H"], g["W"]
un = synth.c3_unaries(H, W, seed=7, hard=True)
ref_fg = synth.unpack_mask(g["mask_fg_packed"], H, W)
spec = A.GridSpecHost("index", 1, H, W)
def model(prec):
    fl = g["schedule"]["flow"]
    flow = A.init_realnvp(channels=2, n_flows=fl["n_flows"], hidden_units=fl["hidden_units"], height=H, width=W, output_fn="tanh")
    norm = A.get_norm("minmax", dim=(0, 2, 3)); norm.fit(spec.materialize(2, "cpu"))
    m = A.PathConnectedNet(convex_net=A.ConvexNextNet(n_hidden_layers=2, precision=prec), flow_net=A.NormNet(net=A.PixelizeNet(flow), norm=norm))
    m.load_state_dict(g["after_identity"]); return m.cuda()
idx = [0, 1, 2, 10, 100, 300, 600, 1000, 1500, 1900, 1999]
print("ref ", [f"{float(g['loss_hist'][i]):.5f}" for i in idx])
for prec in ("fp32", "f16"):
    m = model(prec)
    opt = A.OptimConfig("adamax", lr=[2e-3, 1e-3, 1e-3, 1e-3], weight_decay=[1e-5, 0.0, 0.0, 0.0])
    f = m.make_fitter(spec, un.cuda(), A.LossConfig("fgbg_bce_logits", fg_weight=0.3), opt)
    hist = f.run(2000).reshape(-1).cpu()
    print(prec, [f"{float(hist[i]):.5f}" for i in idx])
    with torch.no_grad():
        y = m(spec.materialize(2, "cuda")).reshape(H, W).cpu()
    fg = y < 0
    print(prec, "IoU vs target", synth.fg_iou(fg, un < 0.5), "vs ref mask", synth.fg_iou(fg, ref_fg), "fg px", int(fg.sum()), "ref fg px", int(ref_fg.sum()), "target fg", int((un < 0.5).sum()))
    # exact loss of the final weights, evaluated in fp32 on the CPU-side formula
    t = un.reshape(-1); yl = y.reshape(-1)
    bce = torch.nn.functional.binary_cross_entropy_with_logits(yl, t, reduction="none")
    fgm = t != 1.0
    print(prec, "exact final loss", float(0.7 * bce[~fgm].sum() / (~fgm).sum() + 0.3 * bce[fgm].sum() / fgm.sum()))
yr = g["logits_f16"].float().reshape(-1); t = un.reshape(-1); fgm = t != 1.0
bce = torch.nn.functional.binary_cross_entropy_with_logits(yr, t, reduction="none")
print("ref exact final loss", float(0.7 * bce[~fgm].sum() / (~fgm).sum() + 0.3 * bce[fgm].sum() / fgm.sum()))
print("---- IoU trace over the last 200 steps (every 10)")
for prec in ("fp32", "f16"):
    m = model(prec)
    opt = A.OptimConfig("adamax", lr=[2e-3, 1e-3, 1e-3, 1e-3], weight_decay=[1e-5, 0.0, 0.0, 0.0])
    f = m.make_fitter(spec, un.cuda(), A.LossConfig("fgbg_bce_logits", fg_weight=0.3), opt, steps_per_graph=10)
    f.run(1800)
    tr = []
    for k in range(21):
        with torch.no_grad():
            y = m(spec.materialize(2, "cuda")).reshape(H, W).cpu()
        tr.append(synth.fg_iou(y < 0, un < 0.5))
        if k < 20: f.run(10)
    t = torch.tensor(tr)
    print(prec, "mean %.4f std %.4f min %.4f max %.4f" % (t.mean(), t.std(), t.min(), t.max()), [f"{x:.4f}" for x in tr])
