"""Host logic of awesome_b200.fit_frames_grouped without a GPU: the native pieces (pixel counts, IoU counts, the fused
fitter, the one-frame retry path) are replaced by small stand-ins, so that skip handling, grouping, padding of a short
last group, cold / warm step counts, the entry-state broadcast, the retry bookkeeping and the result order are checked
on the CPU."""
import types

import pytest
import torch

import awesome_b200.pretrain as PT


class FakePrior(torch.nn.Module):
    in_features = 2

    def __init__(self, row):
        super().__init__()
        self.row = row
        self.resets = 0

    def reset_parameters(self):
        self.resets += 1
        self.row.fill_(-1.0)

    def _ensure_flat(self):
        return self.row


class FakeFitter:
    def __init__(self, multi, target):
        self.multi, self.target, self.runs, self.resets = multi, target, [], 0

    def set_target(self, target, loss=None):
        self.target = target

    def reset_optimizer(self):
        self.resets += 1

    def run(self, steps):
        self.runs.append((steps, self.multi.big.clone(), self.target.clone()))
        self.multi.big += 1.0                              # "training" moves every prior's row
        return torch.full((max(steps, 1), len(self.multi.priors)), 0.25)

    def raise_if_nonfinite(self):
        pass


class FakeMulti:
    def __init__(self, G, P=5):
        self.big = torch.arange(G * P, dtype=torch.float32).reshape(G, P)
        self.priors = [FakePrior(self.big[k]) for k in range(G)]
        self._arena_all = self.big
        self.fitters = []

    def _group_arena(self):
        return self.big

    def make_fitter(self, grid, target, loss, optim, **kw):
        self.fitters.append(FakeFitter(self, target))
        return self.fitters[-1]

    def __call__(self, x, num_priors=None):
        return torch.zeros(1, len(self.priors), 1, 2, 3)


@pytest.fixture
def patched(monkeypatch):
    calls = {"retry": []}

    def target_counts(un, rule, n_objects=1):             # [fg, bg] pixel counts of the unaries
        fg = int((un < 0.5).sum())
        return torch.tensor([[fg, un.numel() - fg]])

    def iou_counts(pred, target, pred_is_logit, n_objects=1):
        # IoU of frame k = its first target value (lets the test script proper / improper fits)
        out = torch.zeros(n_objects, 4, dtype=torch.int64)
        for k in range(n_objects):
            q = int(round(float(target.reshape(n_objects, -1)[k, 0]) * 100))
            out[k] = torch.tensor([q, 100, 100, 0]) if q < 100 else torch.tensor([100, 100, 100, 0])
            out[k, 0] = min(q, 100)
            out[k, 1], out[k, 2] = 100, out[k, 0]          # pf + tf - inter = 100 -> iou = inter / 100
        return out

    def fit_frames(model, grids, unaries, schedule, frame_indices=None, **kw):
        calls["retry"].append((frame_indices[0], schedule.proper_prior_fit_retrys, schedule.reuse_state))
        return [PT.FrameResult(index=frame_indices[0], iou=0.11, proper_fit=False, retries=0, steps=schedule.num_epochs,
                               final_loss=0.5)]
    monkeypatch.setattr(PT, "target_counts", target_counts)
    monkeypatch.setattr(PT, "iou_counts", iou_counts)
    monkeypatch.setattr(PT, "fit_frames", fit_frames)
    monkeypatch.setattr(PT, "_as_grid", lambda g, dev: types.SimpleNamespace(materialize=lambda c, d: torch.zeros(1, c, 2, 3)))
    return calls


def frame(first, fg=True, n=6):
    u = torch.full((n,), 0.9)
    if fg:
        u[1:3] = 0.1
    u[0] = first                                           # first value scripts the IoU (see iou_counts above)
    return u


def test_grouping_skip_padding_warm_start_and_order(patched):
    multi = FakeMulti(G=3)
    entry = multi.big[0].clone()
    frames = [frame(0.95), frame(0.9, fg=False), frame(0.93), frame(0.97), frame(0.99)]
    frames[1][:] = 0.9                                     # no foreground at all -> skipped
    seen = []
    sched = PT.FitSchedule(num_epochs=40, reuse_state_epochs=7, proper_prior_fit_threshold=0.5, proper_prior_fit_retrys=1)
    res = PT.fit_frames_grouped(multi, object(), frames, sched, on_frame=lambda r: seen.append(r.index),
                                frame_indices=[10, 11, 12, 13, 14])
    assert [r.index for r in res] == [10, 11, 12, 13, 14]
    assert [r.skipped for r in res] == [False, True, False, False, False]
    assert [r.steps for r in res] == [40, 0, 40, 40, 7]            # first group cold, second group warm
    assert sorted(seen) == [10, 11, 12, 13, 14] and not patched["retry"]
    f = multi.fitters[0]
    assert len(multi.fitters) == 1 and f.resets == 2 and [r[0] for r in f.runs] == [40, 7]
    # group 1: every prior starts from the entry state; targets are frames 10, 12, 13
    assert torch.equal(f.runs[0][1], entry.expand(3, -1))
    assert torch.equal(f.runs[0][2], torch.stack([frames[0], frames[2], frames[3]]))
    # group 2: the short group repeats its last frame, and starts from the last proper state of group 1 (frame 13's)
    assert torch.equal(f.runs[1][2], torch.stack([frames[4]] * 3))
    assert torch.equal(f.runs[1][1], res[3].state.expand(3, -1))
    assert all(r.proper_fit and r.iou == pytest.approx(v) for r, v in zip((res[0], res[2], res[3], res[4]), (0.95, 0.93, 0.97, 0.99)))
    assert res[4].final_loss == 0.25 and res[1].state is None and res[0].state is not None


def test_improper_fit_goes_through_the_reference_retry_once(patched):
    multi = FakeMulti(G=2)
    frames = [frame(0.2), frame(0.8)]                      # IoU 0.2 < threshold -> retry; 0.8 proper
    sched = PT.FitSchedule(num_epochs=30, reuse_state=False, proper_prior_fit_threshold=0.5, proper_prior_fit_retrys=1)
    res = PT.fit_frames_grouped(multi, object(), frames, sched)
    assert patched["retry"] == [(0, 0, False)]             # one-frame path, no further retries, no chaining
    assert multi.priors[0].resets == 1 and multi.priors[1].resets == 0
    assert (res[0].retries, res[0].steps, res[0].proper_fit, res[0].iou) == (1, 60, False, 0.11)
    assert (res[1].retries, res[1].steps, res[1].proper_fit) == (0, 30, True)
    # with retries disabled the improper frame is reported as is
    patched["retry"].clear()
    res = PT.fit_frames_grouped(FakeMulti(G=2), object(), frames, PT.FitSchedule(num_epochs=5, reuse_state=False,
                                                                                   proper_prior_fit_retrys=0))
    assert not patched["retry"] and not res[0].proper_fit and res[0].iou == pytest.approx(0.2)


def test_flow_priors_are_refused(patched):
    multi = FakeMulti(G=2)
    multi.priors[1].flow_net = object()
    with pytest.raises(NotImplementedError):
        PT.fit_frames_grouped(multi, object(), [frame(0.9)], PT.FitSchedule())
