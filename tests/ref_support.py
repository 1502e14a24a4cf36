"""Synthetic stand-ins the REFERENCE's runner can resolve by dotted name (``AwesomeConfig.dataset_type`` /
``segmentation_model_type``) in ``tests/test_gpu_reference_agent.py``.  Importable only after ``oracle.ref_shim.install()``
(the classes derive from the reference's ``PriorDataset`` / ``TorchDataSource``)."""
import torch
from awesome.dataset.prior_dataset import PriorDataset, prior
from awesome.dataset.torch_datasource import TorchDataSource

H, W, T = 40, 56, 3


def blob(H, W, cx=0.5, cy=0.5, rx=0.22, ry=0.27, tau=0.08):
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    return torch.sigmoid((torch.sqrt(((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2) - 1) / tau)


class SynthFrames(PriorDataset, TorchDataSource):
    """Three synthetic frames in the item format of ``AwesomeDataset`` (image mode, ``param_clean_grid``): inputs =
    (image [4,H,W], feature grid, clean coordinate grid [2,H,W]), label map; the ``@prior()`` decorator of the reference
    prepends the frame's prior state."""

    def __init__(self, prior_model_type=None, prior_model_args=None, **kw):
        keep = {k: v for k, v in kw.items() if k in ("split_seed", "split_ratio", "batch_size", "shuffle_in_dataloader")}
        super().__init__(prior_model_type=prior_model_type, prior_model_args=prior_model_args, returns_index=False, **keep)
        self.frames = [blob(H, W, cx=0.42 + 0.05 * i) for i in range(T)]
        yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
        self.grid = torch.stack([xx, yy])                  # Transformator.get_positional_matrices order: (x, y)

    def __len__(self):
        return T

    def get_number_of_classes(self):
        return 2

    @prior()
    def __getitem__(self, i):
        u = self.frames[i]
        image = torch.stack([u, u * 0.5, 1 - u, torch.zeros_like(u)])
        return (image, self.grid.clone(), self.grid.clone()), (u > 0.5).float()[None]


class TinySeg(torch.nn.Module):
    """Frozen "UNet": its logit is a fixed function of the image's first channel (sigmoid gives back the blob)."""

    def __init__(self, in_chn=4, out_chn=1, **kw):
        super().__init__()
        self.conv = torch.nn.Conv2d(4, 1, 1)
        with torch.no_grad():
            self.conv.weight.zero_()
            self.conv.weight[0, 0] = 8.0
            self.conv.bias.fill_(-4.0)

    def forward(self, image, *args, **kwargs):
        return self.conv(image)
