"""GPU parity of the image preprocessing kernels (awb_image.cu, through the C-ABI) against the oracle and the OpenCV
golden outputs: bit-exact, including degenerate sizes and a full FBMS-shaped frame."""
import os

import numpy as np
import pytest
import torch

import __graft_entry__ as entry
from oracle import image_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "image_golden.npz")


@pytest.fixture(scope="module", autouse=True)
def built():
    entry.build()


@pytest.fixture(scope="module")
def A():
    import awesome_b200
    return awesome_b200


def test_kernels_match_opencv_golden_bit_for_bit(A):
    g = np.load(GOLD)
    names = [k[3:] for k in g.files if k.startswith("in_")]
    assert len(names) >= 9
    for name in names:
        img = torch.from_numpy(g[f"in_{name}"]).cuda()
        assert np.array_equal(A.image.process_image(img).cpu().numpy(), g[f"proc_{name}"]), name
        assert np.array_equal(A.image.process_image(img, image_channel_format="bgr").cpu().numpy(), g[f"procbgr_{name}"]), name
        assert np.array_equal(A.image.create_edge_map(img).cpu().numpy(), g[f"edge_{name}"]), name
        assert torch.equal(A.image.process_image(img, do_image_blurring=False), img)


@pytest.mark.parametrize("H,W", [(480, 640), (481, 643), (31, 33), (17, 129), (16, 32), (7, 3),
                                 (8, 128), (37, 260), (130, 132), (200, 1000)])      # streaming kernels: W % 4 == 0, W >= 128, H >= 8
def test_kernels_match_oracle_on_random_frames(A, H, W):
    rng = np.random.default_rng(H * 1000 + W)
    img = rng.random((3, H, W), dtype=np.float32)
    img[:, : H // 2, : W // 3] = (img[:, : H // 2, : W // 3] > 0.5)          # a block of hard edges
    t = torch.from_numpy(img).cuda()
    assert np.array_equal(A.image.process_image(t).cpu().numpy(), O.process_image(img))
    assert np.array_equal(A.image.create_edge_map(t).cpu().numpy(), O.edge_map(img))


def test_properties_and_errors(A):
    t = torch.full((3, 40, 50), 0.3, device="cuda")
    assert not A.image.create_edge_map(t).any()                               # constant frame: exactly zero edges
    e = A.image.create_edge_map(torch.rand(3, 64, 64, device="cuda"))
    assert e.shape == (1, 64, 64) and float(e.min()) >= 0.0 and float(e.max()) <= 1.0
    batch = torch.rand(5, 3, 33, 47, device="cuda")                           # a batch == its frames one by one
    assert torch.equal(A.image.process_image(batch), torch.stack([A.image.process_image(f) for f in batch]))
    assert torch.equal(A.image.create_edge_map(batch), torch.stack([A.image.create_edge_map(f) for f in batch]))
    wide = torch.rand(7, 3, 70, 256, device="cuda")                           # the same through the streaming kernels
    assert torch.equal(A.image.process_image(wide), torch.stack([A.image.process_image(f) for f in wide]))
    assert torch.equal(A.image.process_image(wide, image_channel_format="bgr"), A.image.process_image(wide).flip(1))
    assert torch.equal(A.image.create_edge_map(wide), torch.stack([A.image.create_edge_map(f) for f in wide]))
    assert torch.equal(A.image.create_edge_map(wide.flip(3)), A.image.create_edge_map(wide).flip(3))
    assert torch.equal(A.image.create_edge_map(wide.flip(2)), A.image.create_edge_map(wide).flip(2))
    flipped = A.image.process_image(t.flip(2))                                # reflect-101 borders are mirror symmetric
    assert torch.equal(flipped, A.image.process_image(t).flip(2))
    with pytest.raises(ValueError):
        A.image.process_image(torch.rand(1, 8, 8, device="cuda"))
    with pytest.raises(ValueError):
        A.image.create_edge_map(torch.rand(3, 8, 8))
