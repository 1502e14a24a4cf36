"""Generates tests/golden/image_golden.npz with OpenCV itself (the library the reference calls in
awesome/dataset/image_sample.py:212-221,260-275): inputs and cv2 outputs of _process_image / create_edge_map for a few
small frames, including degenerate sizes.  Run in the build container:  python tests/golden/make_image_golden.py"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_process(image, bgr=False):                      # ImageSample._process_image
    im = (image.transpose(1, 2, 0) * 255).astype(np.uint8)
    im = cv2.GaussianBlur(im, (5, 5), 0)
    out = im.astype(np.float32).transpose(2, 0, 1) / np.float32(255)
    return out[[2, 1, 0]] if bgr else out


def ref_edge(image):                                    # ImageSample.create_edge_map
    im = (image.transpose(1, 2, 0) * 255).astype(np.uint8)
    src = cv2.GaussianBlur(im, (3, 3), 0)
    gray = cv2.cvtColor(src, cv2.COLOR_RGB2GRAY)
    gx = cv2.Sobel(gray, cv2.CV_16S, 1, 0, ksize=3, scale=1, delta=0, borderType=cv2.BORDER_DEFAULT)
    gy = cv2.Sobel(gray, cv2.CV_16S, 0, 1, ksize=3, scale=1, delta=0, borderType=cv2.BORDER_DEFAULT)
    g = cv2.addWeighted(cv2.convertScaleAbs(gx), 0.5, cv2.convertScaleAbs(gy), 0.5, 0)
    g = g / 255
    return cv2.GaussianBlur(g, (5, 5), 0).astype(np.float32)[None]


def frames():
    rng = np.random.default_rng(7)
    for k, (H, W) in enumerate([(48, 72), (37, 53), (33, 65), (5, 7), (3, 3), (2, 9), (1, 6), (6, 1), (1, 1)]):
        kind = k % 3
        if kind == 0:
            img = rng.random((3, H, W), dtype=np.float32)
        elif kind == 1:                                 # hard edges, exact 0 / 1 values
            img = (rng.random((3, H, W)) > 0.5).astype(np.float32)
        else:                                           # smooth ramps + a disc
            yy, xx = np.mgrid[0:H, 0:W]
            disc = (((xx - W / 2) ** 2 + (yy - H / 2) ** 2) < (min(H, W) / 3) ** 2).astype(np.float32)
            img = np.stack([xx / max(W, 1), yy / max(H, 1), disc]).astype(np.float32)
        yield f"{H}x{W}", img


if __name__ == "__main__":
    out = {"cv2_version": np.array(cv2.__version__)}
    for name, img in frames():
        out[f"in_{name}"] = img
        out[f"proc_{name}"] = ref_process(img)
        out[f"procbgr_{name}"] = ref_process(img, bgr=True)
        out[f"edge_{name}"] = ref_edge(img)
    np.savez_compressed(os.path.join(HERE, "image_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "image_golden.npz"), "with OpenCV", cv2.__version__)
