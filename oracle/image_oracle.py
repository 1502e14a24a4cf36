"""CPU restatement (numpy, integer / fixed-point exact) of the reference's per-frame image preprocessing -- TEST
INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import it; the product path is
awesome_b200/csrc/awb_image.cu).

Follows ``awesome/dataset/image_sample.py``:
  * ``ImageSample._process_image`` (:212-221): ``(image * 255).astype(uint8)`` -> ``cv2.GaussianBlur(.., (5, 5), 0)`` ->
    ``/ 255`` in float32 -> optional BGR channel order;
  * ``ImageSample.create_edge_map`` (:260-275): uint8 -> ``GaussianBlur (3, 3)`` -> ``COLOR_RGB2GRAY`` -> ``Sobel`` x / y
    (``CV_16S``, ksize 3) -> ``convertScaleAbs`` -> ``addWeighted(0.5, 0.5)`` -> ``/ 255`` (float64) ->
    ``GaussianBlur (5, 5)`` -> float32 ``[1,H,W]``.

The algorithm itself lives in OpenCV (third-party; ``opencv-python`` 4.x, imported by the reference), so its published
semantics are restated here: sigma = 0 selects the fixed binomial kernels [1 2 1]/4 and [1 4 6 4 1]/16; 8-bit images are
filtered in fixed point and rounded half up once at the end; borders are BORDER_REFLECT_101; RGB2GRAY is
``(R*9798 + G*19235 + B*3735 + 16384) >> 15`` (the 15-bit coefficients of OpenCV 4.x); ``addWeighted`` on uint8 rounds half to even (``cvRound``); floating-point
images are filtered separably, rows first, in the symmetric form ``k0*x0 + k1*(x-1 + x+1) + k2*(x-2 + x+2)``.
Pinned against OpenCV itself by ``tests/golden/make_image_golden.py`` -> ``tests/golden/image_*.npz``."""
from __future__ import annotations

import numpy as np


def reflect101(i: np.ndarray, n: int) -> np.ndarray:
    """cv2.BORDER_REFLECT_101 index map (gfedcb|abcdefgh|gfedcba)."""
    i = np.asarray(i).copy()
    if n == 1:
        return np.zeros_like(i)
    while True:
        lo, hi = i < 0, i >= n
        if not (lo.any() or hi.any()):
            return i
        i = np.where(lo, -i, i)
        i = np.where(i >= n, 2 * n - 2 - i, i)


def _taps(a: np.ndarray, radius: int):
    """a[..., H, W] -> dict (dy, dx) -> shifted copy with reflect-101 borders."""
    H, W = a.shape[-2:]
    ys = [reflect101(np.arange(H) + d, H) for d in range(-radius, radius + 1)]
    xs = [reflect101(np.arange(W) + d, W) for d in range(-radius, radius + 1)]
    return ys, xs


def blur_u8(img: np.ndarray, ksize: int) -> np.ndarray:
    """cv2.GaussianBlur(img_u8, (ksize, ksize), 0) for ksize in {3, 5}; img [..., H, W] uint8."""
    k = {3: np.array([1, 2, 1], np.int64), 5: np.array([1, 4, 6, 4, 1], np.int64)}[ksize]
    r = ksize // 2
    ys, xs = _taps(img, r)
    a = img.astype(np.int64)
    rows = sum(k[j] * a[..., :, xs[j]] for j in range(ksize))
    full = sum(k[j] * rows[..., ys[j], :] for j in range(ksize))
    s = int(k.sum()) ** 2                       # 16 or 256
    return ((full + s // 2) // s).astype(np.uint8)


def to_u8(image: np.ndarray) -> np.ndarray:
    """``(image * 255).astype(np.uint8)``: float32 product, truncation toward zero (values are in [0, 1])."""
    return (image.astype(np.float32) * np.float32(255)).astype(np.uint8)


def process_image(image: np.ndarray, do_image_blurring: bool = True, bgr: bool = False) -> np.ndarray:
    """image [3,H,W] float32 in [0,1] -> float32 [3,H,W]  (image_sample.py:212-221)."""
    out = image.astype(np.float32)
    if do_image_blurring:
        out = blur_u8(to_u8(image), 5).astype(np.float32) / np.float32(255)
    if bgr:
        out = out[[2, 1, 0]]
    return out


def rgb2gray_u8(rgb: np.ndarray) -> np.ndarray:
    r, g, b = (rgb[c].astype(np.int64) for c in range(3))
    return ((r * 9798 + g * 19235 + b * 3735 + 16384) >> 15).astype(np.uint8)


def sobel_s16(gray: np.ndarray):
    ys, xs = _taps(gray, 1)
    a = gray.astype(np.int64)
    sm_y = a[ys[0], :] + 2 * a[ys[1], :] + a[ys[2], :]          # smooth along y
    gx = sm_y[:, xs[2]] - sm_y[:, xs[0]]
    sm_x = a[:, xs[0]] + 2 * a[:, xs[1]] + a[:, xs[2]]
    gy = sm_x[ys[2], :] - sm_x[ys[0], :]
    return gx, gy


def blur5_f64(a: np.ndarray) -> np.ndarray:
    """cv2.GaussianBlur(float64 image, (5, 5), 0): separable, rows first, symmetric summation order."""
    k0, k1, k2 = 0.375, 0.25, 0.0625
    ys, xs = _taps(a, 2)
    rows = k0 * a[:, xs[2]] + k1 * (a[:, xs[1]] + a[:, xs[3]]) + k2 * (a[:, xs[0]] + a[:, xs[4]])
    return k0 * rows[ys[2], :] + k1 * (rows[ys[1], :] + rows[ys[3], :]) + k2 * (rows[ys[0], :] + rows[ys[4], :])


def edge_map(image: np.ndarray) -> np.ndarray:
    """image [3,H,W] float32 RGB in [0,1] -> float32 [1,H,W]  (image_sample.py:260-275)."""
    src = blur_u8(to_u8(image), 3)
    gray = rgb2gray_u8(src)
    gx, gy = sobel_s16(gray)
    ax = np.minimum(np.abs(gx), 255)                              # convertScaleAbs: saturate_cast<uchar>(|x|)
    ay = np.minimum(np.abs(gy), 255)
    s = ax + ay                                                   # addWeighted(.5, .5): cvRound((ax + ay) / 2), half to even
    grad = s // 2 + ((s & 1) & ((s // 2) & 1))
    g = grad.astype(np.float64) / 255
    return blur5_f64(g).astype(np.float32)[None]
