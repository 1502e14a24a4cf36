"""GPU test of the grouped per-frame pretrain loop (awesome_b200.fit_frames_grouped): G frames per fused launch."""
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


def blob(H, W, cx=0.5, cy=0.5, rx=0.27, ry=0.31, tau=0.08):
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    return torch.sigmoid((torch.sqrt(((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2) - 1) / tau)


def test_fit_frames_grouped_matches_independent_frame_fits(A):
    """fit_frames_grouped: G frames per fused launch, every frame of a group starting from the group's entry state --
    the same masks as fitting each frame on its own from that state; skip, short last group, warm second group, retry."""
    H, W, G = 60, 80, 3
    torch.manual_seed(2)
    multi = A.NumberBasedMultiPriorModule(prior_type=A.ConvexNextNet,
                                          prior_args=dict(n_hidden_layers=2, precision="f16"), min_priors=G).to(DEV)
    entry = {k: v.clone() for k, v in multi.priors[0].state_dict().items()}
    frames = [blob(H, W, cx=0.40 + 0.04 * i, cy=0.5, rx=0.2, ry=0.25) for i in range(5)]
    frames[1] = torch.ones(H, W)                      # background only -> skipped, does not occupy a slot
    grid = A.GridSpecHost("linspace", 1, H, W)
    sched = A.FitSchedule(num_epochs=600, reuse_state_epochs=150, optimizer="adam", plateau=False, lr=2e-3,
                          proper_prior_fit_threshold=0.5)
    seen = []
    res = A.fit_frames_grouped(multi, grid, frames, sched, on_frame=lambda r: seen.append(r.index))
    assert [r.index for r in res] == [0, 1, 2, 3, 4] and sorted(seen) == [0, 1, 2, 3, 4]
    assert [r.skipped for r in res] == [False, True, False, False, False]
    assert [r.steps for r in res] == [600, 0, 600, 600, 150]           # second group (frame 4 alone) is warm
    for r in res:
        if not r.skipped:
            assert r.proper_fit and r.iou > 0.9 and r.state is not None, (r.index, r.iou)
    # frame 2 on its own from the same entry state: same mask to within a handful of boundary pixels
    single = A.ConvexNextNet(n_hidden_layers=2, precision="f16").to(DEV)
    single.load_state_dict(entry)
    alone = A.fit_frames(single, [grid], [frames[2]], A.FitSchedule(num_epochs=600, optimizer="adam", plateau=False, lr=2e-3,
                                                                   reuse_state=False))[0]
    assert abs(alone.iou - res[2].iou) < 1e-3 and alone.final_loss == pytest.approx(res[2].final_loss, rel=1e-3)
    # an impossible threshold sends every frame through the reference's reset + refit retry exactly once
    hard = A.FitSchedule(num_epochs=60, optimizer="adam", plateau=False, lr=2e-3, reuse_state=False,
                         proper_prior_fit_threshold=1.01, proper_prior_fit_retrys=1)
    res2 = A.fit_frames_grouped(multi, grid, frames[2:4], hard)
    assert [r.retries for r in res2] == [1, 1] and not any(r.proper_fit for r in res2)
    assert [r.steps for r in res2] == [120, 120]


@pytest.mark.parametrize("group", [1, 2])
def test_sharded_sequence_fit_is_independent_of_the_number_of_ranks(A, group):
    """fit_sequence_sharded: the segmentation is fixed, so the per-frame results of a 1-rank run and of a 2-rank run
    (simulated: rank 0 and rank 1 one after the other, results merged) are bit-identical -- states, masks, IoUs."""
    from awesome_b200 import synth
    from awesome_b200.sharding import merge_by_unit
    T, H, W = 10, 60, 80
    frames = [blob(H, W, cx=0.40 + 0.02 * i, cy=0.5, rx=0.2, ry=0.25) for i in range(T)]
    frames[6] = torch.ones(H, W)                      # one frame without foreground -> skipped
    grid = A.GridSpecHost("linspace", 1, H, W)
    sched = A.FitSchedule(num_epochs=300, reuse_state_epochs=80, optimizer="adam", plateau=False, lr=2e-3, steps_per_graph=20)
    args = dict(n_hidden_layers=2, precision="f16")
    kw = dict(n_segments=4, group=group, device=DEV, seed=7)
    one = A.fit_sequence_sharded(A.ConvexNextNet, args, grid, frames, T, sched, rank=0, world=1, **kw)
    two = merge_by_unit([A.fit_sequence_sharded(A.ConvexNextNet, args, grid, frames, T, sched, rank=r, world=2, gather=False, **kw)
                         for r in range(2)])
    assert sorted(one) == sorted(two) == list(range(T))
    assert A.plan_segments(T, 4) == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]
    for i in range(T):
        a, b = one[i], two[i]
        assert a["skipped"] == b["skipped"] == (i == 6)
        assert a["segment"] == b["segment"] and b["rank"] == a["segment"] % 2
        if a["skipped"]:
            continue
        assert torch.equal(a["state"], b["state"]) and torch.equal(a["mask_fg_packed"], b["mask_fg_packed"]), i
        assert a["iou"] == b["iou"] and a["steps"] == b["steps"]
        assert a["proper_fit"] and a["iou"] > 0.9
        fg = synth.unpack_mask(a["mask_fg_packed"], H, W)
        assert abs(synth.fg_iou(fg, frames[i] < 0.5) - a["iou"]) < 2e-3
    # first group of a segment is cold, the following groups warm (group = 1, the default: the reference's frame-by-frame chain)
    steps = [one[i]["steps"] for i in (0, 1, 2, 3, 7, 8, 9)]
    assert steps == ([300, 80, 80, 300, 300, 300, 80] if group == 1 else [300, 300, 80, 300, 300, 300, 300])
