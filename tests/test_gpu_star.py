"""GPU parity of the star-shape prior against the fixture generated from the reference notebook's
own class (tests/golden/star.pt) and against the CPU oracle loop."""
import pytest
import torch

import __graft_entry__ as entry
from oracle import prior_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def built():
    entry.build()


@pytest.fixture(scope="module")
def A():
    import awesome_b200
    return awesome_b200


def test_star_forward_loss_and_gradients_vs_notebook_class(A, golden):
    g = golden("star.pt")
    m = A.StarShapedNet(150)
    assert list(m.state_dict().keys()) == list(g["init"].keys())
    m.load_state_dict(g["init"])
    m = m.to(DEV)
    y = m(g["x"].to(DEV)).cpu()
    torch.testing.assert_close(y, g["logits"], rtol=1e-4, atol=2e-5)
    # one fused step with lr = 0 leaves the weights alone and exposes loss and gradients (Adam: exp_avg = 0.1 * g)
    f = m.make_fitter(optim=A.OptimConfig("adam", lr=0.0))
    f.train_offset = True
    loss = f.step(g["x"], g["t"])
    torch.testing.assert_close(loss.cpu()[0], g["loss"], rtol=1e-5, atol=1e-8)
    P = m._arena.numel()
    grads = f.opt_state[:4 * P].view(torch.float32).clone().cpu() * 10.0
    off = 0
    for k, p in m.named_parameters():
        n = p.numel()
        torch.testing.assert_close(grads[off:off + n].reshape(p.shape), g["grads"][k], rtol=2e-3, atol=2e-7,
                                   msg=lambda s: f"{k}: {s}")
        off += n


def test_star_fit_steps_vs_oracle_adam_with_late_offset(A, golden):
    """5 steps of cell 3 (Adam lr 1e-2, clamp of W2_r.weight); the offset joins at step 2 with its own Adam step count
    (torch keeps the step per parameter)."""
    g = golden("star.pt")
    m = A.StarShapedNet(150)
    m.load_state_dict(g["init"])
    m = m.to(DEV)
    f = m.make_fitter(optim=A.OptimConfig("adam", lr=1e-2))
    p = O.clone_params(g["init"], requires_grad=True)
    keys = list(g["grads"])
    opt_keys = [k for k in keys if k != "offset"]
    opt = torch.optim.Adam([p[k] for k in opt_keys], lr=1e-2)
    opt_off = torch.optim.Adam([p["offset"]], lr=1e-2)
    ours, ref = [], []
    for step in range(5):
        f.train_offset = step >= 2
        ours.append(float(f.step(g["x"], g["t"])))
        opt.zero_grad(); opt_off.zero_grad()
        loss = torch.nn.functional.mse_loss(torch.sigmoid(O.star_forward(p, g["x"])).squeeze(), g["t"])
        loss.backward()
        opt.step()
        if step >= 2:
            opt_off.step()
        with torch.no_grad():
            p["W2_r.weight"].clamp_(min=0)
        ref.append(float(loss))
    torch.testing.assert_close(torch.tensor(ours), torch.tensor(ref), rtol=2e-4, atol=1e-7)
    sd = m.state_dict()
    for k in keys:
        torch.testing.assert_close(sd[k].cpu(), p[k].detach(), rtol=2e-3, atol=2e-5, msg=lambda s: f"{k}: {s}")
    assert float(sd["W2_r.weight"].min()) >= 0.0


def test_star_full_image_forward_and_ragged_sizes(A):
    torch.manual_seed(0)
    m = A.StarShapedNet(150).to(DEV)
    p = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    for n in (1, 5, 1000, 64 * 48):
        x = torch.rand(n, 2) - 0.5
        torch.testing.assert_close(m(x.to(DEV)).cpu(), O.star_forward(p, x), rtol=1e-4, atol=2e-5)
