"""Debug helper (GPU box): tensor-path vs fp32-path reduced gradients of one fit step, per tensor."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import awesome_b200 as A

L = int(sys.argv[1]) if len(sys.argv) > 1 else 2
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (96, 128)
torch.manual_seed(0)
m32 = A.ConvexNextNet(n_hidden_layers=L).cuda()
m16 = A.ConvexNextNet(n_hidden_layers=L, precision="f16"); m16.load_state_dict(m32.state_dict()); m16 = m16.cuda()
yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
un = torch.sigmoid((torch.sqrt(((xx - 0.52) / 0.27) ** 2 + ((yy - 0.47) / 0.31) ** 2) - 1) / 0.08).cuda()
grid = A.GridSpecHost("linspace", 1, H, W)
P = m32._arena.numel()
gs = {}
for name, m in (("fp32", m32), ("f16", m16)):
    f = m.make_fitter(grid, un, A.LossConfig("mse"), A.OptimConfig("adam", lr=1e-3), use_graph=False)
    f.run(1); torch.cuda.synchronize()
    gs[name] = f.opt_state[:4 * P].view(torch.float32).clone() * 10.0
    print(name, "loss", f.scalars().last_loss)
off = 0
for name, p in m32.named_parameters():
    n = p.numel()
    a, b = gs["f16"][off:off + n].reshape(p.shape), gs["fp32"][off:off + n].reshape(p.shape)
    rel = float((a - b).norm() / b.norm().clamp(min=1e-20))
    print(f"{name:22s} rel {rel:9.3e}  |ref| {float(b.norm()):.3e} |ours| {float(a.norm()):.3e}")
    if rel > 1e-2:
        if a.dim() == 2 and a.shape[0] > 8:
            rows = (a - b).norm(dim=1) / b.norm(dim=1).clamp(min=1e-20)
            badr = (rows > 1e-2).nonzero().reshape(-1).tolist()
            print("    bad rows:", badr[:20], "... n=", len(badr))
            if a.shape[1] > 8:
                cols = (a - b).norm(dim=0) / b.norm(dim=0).clamp(min=1e-20)
                badc = (cols > 1e-2).nonzero().reshape(-1).tolist()
                print("    bad cols:", badc[:20], "... n=", len(badc))
        print("    ours", a.reshape(-1)[:6].tolist()); print("    ref ", b.reshape(-1)[:6].tolist())
        ratio = (a.reshape(-1)[:6] / b.reshape(-1)[:6]).tolist(); print("    ratio", ratio)
    off += n
