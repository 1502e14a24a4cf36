"""GPU parity of the older diffeomorphism prior (nn.Linear -> NormalizingFlow1D -> ConvexNextNet) against the
fixture generated from the reference's ConvexDiffeomorphismNet (tests/golden/diffeo.pt)."""
import pytest
import torch

import __graft_entry__ as entry

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def built():
    entry.build()


@pytest.fixture(scope="module")
def A():
    import awesome_b200
    return awesome_b200


def make(A, g, state, precision="fp32"):
    m = A.ConvexDiffeomorphismNet(n_hidden=130, n_hidden_layers=1, nf_layers=4, nf_hidden=70, precision=precision)
    assert list(m.state_dict().keys()) == list(g[state].keys())
    m.load_state_dict(g[state])
    return m.to(DEV)


def test_init_state_forward(A, golden):
    g = golden("diffeo.pt")
    m = make(A, g, "init")
    x = g["grid"].to(DEV)
    B, C, H, W = x.shape
    xd = m.get_deformation(x).permute(0, 2, 3, 1).reshape(-1, 2).cpu()
    torch.testing.assert_close(xd, g["deformed"], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(m(x).cpu(), g["logits"], rtol=1e-4, atol=2e-5)


def test_trained_like_state_forward_and_all_gradients(A, golden):
    """Non-zero WNScale bias, weight-norm (g, v) gradients, the full 2x2 linear: loss and every gradient vs the reference."""
    g = golden("diffeo.pt")
    m = make(A, g, "pert")
    x = g["grid"].to(DEV)
    xd = m.get_deformation(x).permute(0, 2, 3, 1).reshape(-1, 2).cpu()
    torch.testing.assert_close(xd, g["pert_deformed"], rtol=1e-4, atol=2e-5)
    y = m(x)
    torch.testing.assert_close(y.detach().cpu(), g["pert_logits"], rtol=1e-4, atol=5e-5)
    loss = ((torch.sigmoid(y) - g["pert_unaries"].to(DEV)) ** 2).mean()
    loss.backward()
    torch.testing.assert_close(loss.detach().cpu(), g["pert_loss"], rtol=1e-5, atol=1e-7)
    for k, p in m.named_parameters():
        torch.testing.assert_close(p.grad.cpu(), g["pert_grads"][k], rtol=5e-3, atol=5e-7, msg=lambda s: f"{k}: {s}")


def test_fused_fit_step_fp32_and_tensor_path(A, golden):
    """One fused fit step (Adam) == autograd step with torch.optim.Adam on the same module; the tensor path agrees
    with the fp32 path within the fp16-operand tolerance."""
    g = golden("diffeo.pt")
    H, W = 48, 64
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    un = torch.sigmoid((torch.sqrt(((xx - 0.5) / 0.3) ** 2 + ((yy - 0.45) / 0.25) ** 2) - 1) / 0.1).to(DEV)
    grid = A.GridSpecHost("linspace", 1, H, W)
    m_ref = make(A, g, "pert")
    opt = torch.optim.Adam(m_ref.parameters(), lr=1e-3)
    y = m_ref(grid.materialize(2, DEV))
    loss_ref = ((torch.sigmoid(y)[0, 0] - un) ** 2).mean()
    loss_ref.backward()
    opt.step()
    m_ref.convex_net.enforce_convexity()
    res = {}
    for prec in ("fp32", "f16"):
        m = make(A, g, "pert", precision=prec)
        f = m.make_fitter(grid, un, A.LossConfig("mse"), A.OptimConfig("adam", lr=1e-3), use_graph=False)
        res[prec] = (float(f.run(1)[0, 0]), {k: v.detach().clone() for k, v in m.state_dict().items()})
    assert abs(res["fp32"][0] - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
    for k, v in m_ref.state_dict().items():
        torch.testing.assert_close(res["fp32"][1][k], v, rtol=1e-4, atol=2e-6, msg=lambda s: f"{k}: {s}")
    assert abs(res["f16"][0] - res["fp32"][0]) <= 2e-3 * abs(res["fp32"][0])
    # Adam's first step is lr * sign(g): parameters of both paths move the same way wherever the gradient is not noise
    moved32 = torch.cat([(res["fp32"][1][k] - g["pert"][k].to(DEV)).reshape(-1) for k in g["pert"]])
    moved16 = torch.cat([(res["f16"][1][k] - g["pert"][k].to(DEV)).reshape(-1) for k in g["pert"]])
    big = moved32.abs() > 5e-4
    assert float((torch.sign(moved16[big]) == torch.sign(moved32[big])).float().mean()) > 0.995


def test_plateau_reduces_every_group_including_weight_g(A, golden):
    """ReduceLROnPlateau halves EVERY param group, also the weight-norm gains' own group 3
    (convex_diffeomorphism_net.py:399-400; ADVICE r1: the scheduler loop used to stop at group 2)."""
    g = golden("diffeo.pt")
    H, W = 24, 32
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    un = torch.sigmoid((torch.sqrt(((xx - 0.5) / 0.3) ** 2 + ((yy - 0.45) / 0.25) ** 2) - 1) / 0.1).to(DEV)
    m = make(A, g, "pert")
    # lr = 0: the loss cannot improve, so the scheduler (patience 2) must fire after 4 steps
    opt = A.OptimConfig("adam", lr=[1e-9, 2e-9, 3e-9, 4e-9], weight_decay=[0, 0, 0, 5e-5], plateau=True, patience=2, factor=0.5,
                        threshold=0.5, plateau_eps=0.0)
    f = m.make_fitter(A.GridSpecHost("linspace", 1, H, W), un, A.LossConfig("mse"), opt, use_graph=False)
    f.run(3)
    before = list(f.scalars(0).lr)
    f.run(3)
    after = list(f.scalars(0).lr)
    assert before == pytest.approx([1e-9, 2e-9, 3e-9, 4e-9])
    assert after == pytest.approx([0.5e-9, 1e-9, 1.5e-9, 2e-9]), after
