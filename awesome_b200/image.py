"""Device versions of the reference's per-frame image preprocessing, bit-identical to the OpenCV calls of
``awesome/dataset/image_sample.py``: ``process_image`` = ``ImageSample._process_image`` (:212-221),
``create_edge_map`` = ``ImageSample.create_edge_map`` (:260-275).  Both take the frame as the reference holds it
(``[3,H,W]`` float RGB in [0,1]; a ``[T,3,H,W]`` batch of frames goes through one launch) -- here on the device, so a sequence can be prepared without the host round trip
through numpy / cv2."""
from __future__ import annotations

import torch

from . import _lib as L


def _check(image: torch.Tensor) -> torch.Tensor:
    L.require_cuda()
    if image.dim() not in (3, 4) or image.shape[-3] != 3:
        raise ValueError(f"expected a [3,H,W] RGB image or a [T,3,H,W] batch of frames, got {tuple(image.shape)}")
    if not image.is_cuda:
        raise ValueError("awesome_b200.image works on CUDA tensors (no CPU fallback)")
    return image.detach().contiguous().float()


def process_image(image: torch.Tensor, do_image_blurring: bool = True, image_channel_format: str = "rgb") -> torch.Tensor:
    """``(image * 255).astype(uint8)`` -> ``cv2.GaussianBlur(.., (5, 5), 0)`` -> ``/ 255`` -> optional BGR order."""
    if image_channel_format not in ("rgb", "bgr"):
        raise ValueError(f"image_channel_format {image_channel_format!r} is not supported")
    img = _check(image)
    out = torch.empty_like(img)
    with torch.cuda.device(img.device):
        L.check(L.load().awb_image_process(img.data_ptr(), out.data_ptr(), img.shape[0] if img.dim() == 4 else 1,
                                           img.shape[-2], img.shape[-1], int(bool(do_image_blurring)),
                                           int(image_channel_format == "bgr"), L.stream_ptr()))
    return out


def create_edge_map(image: torch.Tensor) -> torch.Tensor:
    """Blurred Sobel magnitude ``[1,H,W]`` of a clean RGB frame (the reference caches it per frame as ``.pth``)."""
    img = _check(image)
    T = img.shape[0] if img.dim() == 4 else 1
    shape = (T, 1) + tuple(img.shape[-2:]) if img.dim() == 4 else (1,) + tuple(img.shape[-2:])
    out = torch.empty(shape, dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        L.check(L.load().awb_image_edge_map(img.data_ptr(), out.data_ptr(), T, img.shape[-2], img.shape[-1], L.stream_ptr()))
    return out
