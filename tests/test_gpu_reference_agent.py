"""The drop-in under the REAL reference agent: the reference's ``TorchAgent`` / ``WrapperModule`` / ``PriorDataset`` /
``PriorCache`` (installed in ``baseline/_ref`` by ``baseline/install_reference.py``; non-numeric packages missing offline
are stubbed by ``oracle/ref_shim.py``) drive ``awesome_b200.real_nvp_path_connected_net`` through ``TorchAgent._pretrain``
(``awesome/agent/torch_agent.py:553-627``) on three synthetic frames; the saved ``pretrain_state_path`` is reloaded with
the reference's own ``PriorCache``; and a ``PriorCache`` file written by the reference loads into ``DevicePriorCache``."""
import os

import pytest
import torch

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
DEV = "cuda:0"


def _reference():
    if not os.path.isdir(os.path.join(REF, "awesome")):
        pytest.skip("reference package not installed (python baseline/install_reference.py)")
    from oracle import ref_shim
    if not ref_shim._installed:
        ref_shim.REFERENCE_ROOT = REF
    ref_shim.install()


def blob(H, W, cx=0.5, cy=0.5, rx=0.22, ry=0.27, tau=0.08):
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    return torch.sigmoid((torch.sqrt(((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2) - 1) / tau)


def test_reference_prior_cache_file_loads_into_device_cache(golden):
    """A file written by the reference's ``PriorCache.save`` (fixture ``prior_cache_ref.pth``, ``make_golden_full.py``) is the
    on-disk format ``DevicePriorCache`` reads and writes (``awesome/util/prior_cache.py:61-91``)."""
    import awesome_b200 as A
    state = golden("prior_cache_ref.pth")
    assert set(state) == {"model_type", "model_args", "store_device", "cache"} and set(state["cache"]) == {"0", "5"}
    cache = A.DevicePriorCache(A.ConvexNextNet, dict(n_hidden=130, in_features=2, n_hidden_layers=2), store_device=torch.device("cpu"))
    saved_type = state["model_type"]
    st = dict(state)
    st["model_type"] = "awesome_b200.ConvexNextNet"          # the YAML switch of INTEGRATION.md; keys and tensors untouched
    cache.set_state(st)
    assert 0 in cache and 5 in cache and 1 not in cache
    for key in (0, 5):
        got = cache[key]
        assert list(got.keys()) == list(state["cache"][str(key)].keys())
        for k, v in state["cache"][str(key)].items():
            assert torch.equal(got[k].cpu(), v), (key, k)
    back = cache.get_state()
    assert set(back) == set(state) and saved_type.endswith("ConvexNextNet")
    for k, v in state["cache"]["5"].items():
        assert torch.equal(back["cache"]["5"][k].cpu(), v)


@pytest.mark.gpu
def test_pretrain_under_the_reference_torch_agent(tmp_path):
    entry.build()
    _reference()
    import awesome_b200 as A
    from awesome.agent.torch_agent import TorchAgent
    from awesome.dataset.prior_dataset import PriorDataset, prior
    from awesome.dataset.torch_datasource import TorchDataSource
    from awesome.model.pretrainable_module import PretrainableModule
    from awesome.model.wrapper_module import WrapperModule
    from awesome.util.prior_cache import PriorCache

    assert A.integrate_with_reference(force=True)
    H, W, T = 40, 56, 3
    prior_args = dict(channels=2, hidden_units=32, flow_n_flows=6, flow_output_fn="tanh", norm="minmax",
                      convex_net_hidden_units=130, convex_net_hidden_layers=2, precision="f16")

    class SynthFrames(PriorDataset, TorchDataSource):
        """Three synthetic frames in the item format of ``AwesomeDataset`` (image mode, ``param_clean_grid``): inputs =
        (image [4,H,W], feature grid, clean coordinate grid [2,H,W]), label map."""

        def __init__(self, **kw):
            super().__init__(returns_index=False, **kw)
            self.frames = [blob(H, W, cx=0.42 + 0.05 * i) for i in range(T)]
            self.grid = A.GridSpecHost("linspace", 1, H, W).materialize(2, "cpu")[0]

        def __len__(self):
            return T

        @prior()
        def __getitem__(self, i):
            u = self.frames[i]
            image = torch.stack([u, u * 0.5, 1 - u, torch.zeros_like(u)])
            return (image, self.grid.clone(), self.grid.clone()), (u > 0.5).float()[None]

    class TinySeg(torch.nn.Module):
        """Frozen "UNet": its logit is a fixed function of the image's first channel (so that sigmoid gives back the blob)."""

        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(4, 1, 1)
            with torch.no_grad():
                self.conv.weight.zero_()
                self.conv.weight[0, 0] = 8.0
                self.conv.bias.fill_(-4.0)

        def forward(self, image, *args, **kwargs):
            return self.conv(image)

    ds = SynthFrames(prior_model_type=A.real_nvp_path_connected_net, prior_model_args=prior_args)
    assert ds.has_prior and isinstance(ds.__prior_cache__, PriorCache)
    prior_module = A.real_nvp_path_connected_net(**prior_args)
    assert isinstance(prior_module, PretrainableModule)              # the reference's gate (wrapper_module.py:325-340)
    state_path = str(tmp_path / "pretrain_state.pth")
    agent = TorchAgent(
        name="awb_integration", model_type=WrapperModule,
        model_args=dict(segmentation_module=TinySeg(), prior_module=prior_module, mode="multi", input_mode="image",
                        prior_arg_mode="param_clean_grid", use_segmentation_sigmoid=True),
        optimizer_type=A.FusedAdam, optimizer_args=dict(lr=1e-3), loss=torch.nn.MSELoss(), training_dataset=ds,
        agent_directory=str(tmp_path / "agent"), runs_directory=str(tmp_path / "runs"), do_pretraining=True,
        pretrain_args=dict(num_epochs=250, reuse_state_epochs=60, lr=3e-3, prefit_flow_net_identity=True,
                           prefit_flow_net_identity_num_epochs=15, prefit_convex_net=True, prefit_convex_net_num_epochs=25,
                           do_pretrain_checkpoints=True, pretrain_checkpoint_dir=str(tmp_path / "ck")),
        pretrain_state_path=state_path, force_pretrain=True, pretrain_only=True, device=DEV)
    model = agent._get_prepared_model()
    assert isinstance(model, WrapperModule) and next(model.parameters()).is_cuda
    saved = {}
    agent.save = lambda *a, **k: saved.setdefault("called", True)      # checkpoint serialisation of the agent needs jsonpickle
    agent._pretrain(model, ds, ds, use_progress_bar=False)
    assert os.path.exists(state_path) and saved.get("called")
    assert sorted(os.listdir(tmp_path / "ck")) == [f"pretrain_checkpoint_{i}.pth" for i in range(T)]
    # the saved state is the reference's PriorCache format; the reference's own class reloads it
    state = torch.load(state_path, map_location="cpu", weights_only=False)
    assert set(state) == {"model_type", "model_args", "store_device", "cache"} and set(state["cache"]) == {"0", "1", "2"}
    assert state["model_type"].endswith("real_nvp_path_connected_net")
    pc = PriorCache(None, None)
    pc.set_state(state)
    assert pc.model_type is A.real_nvp_path_connected_net
    # each frame's cached prior, applied by the reference's PriorManager path, reproduces that frame's mask
    from awesome.dataset.prior_dataset import PriorManager
    grid = ds.grid[None].to(DEV)
    for i in range(T):
        with PriorManager(model, prior_state=(i, pc[i]), prior_cache=pc, model_device=torch.device(DEV)):
            with torch.no_grad():
                prob = torch.sigmoid(model.prior_module(grid))
        assert A.mask_iou(prob.reshape(1, -1), ds.frames[i].to(DEV).reshape(1, -1)) > 0.85, i
    # and pretrain_load_state through the agent: a second agent finds the state file and does not fit again
    ds2 = SynthFrames(prior_model_type=A.real_nvp_path_connected_net, prior_model_args=prior_args)
    agent2 = TorchAgent(
        name="awb_integration2", model_type=WrapperModule,
        model_args=dict(segmentation_module=TinySeg(), prior_module=A.real_nvp_path_connected_net(**prior_args), mode="multi",
                        input_mode="image", prior_arg_mode="param_clean_grid"),
        optimizer_type=A.FusedAdam, optimizer_args=dict(lr=1e-3), loss=torch.nn.MSELoss(), training_dataset=ds2,
        agent_directory=str(tmp_path / "agent2"), runs_directory=str(tmp_path / "runs"), do_pretraining=True,
        pretrain_args=dict(num_epochs=250), pretrain_state_path=state_path, force_pretrain=False, device=DEV)
    agent2.save = lambda *a, **k: None
    m2 = agent2._get_prepared_model()
    fits = []
    orig = m2.prior_module.pretrain
    m2.prior_module.pretrain = lambda *a, **k: fits.append(1) or orig(*a, **k)
    agent2._pretrain(m2, ds2, ds2, use_progress_bar=False)
    assert not fits and 1 in ds2.__prior_cache__
    for k, v in state["cache"]["1"].items():
        assert torch.equal(ds2.__prior_cache__[1][k].cpu(), v.cpu()), k
