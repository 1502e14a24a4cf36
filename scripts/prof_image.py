"""ncu target: the two image preprocessing kernels on a 40-frame 640x480 batch (GPU box).
    ncu --set full -k regex:k_image_ python scripts/prof_image.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import awesome_b200 as A

pool = torch.rand((40, 3, 480, 640), device="cuda")
for _ in range(3):
    A.image.process_image(pool)
    A.image.create_edge_map(pool)
torch.cuda.synchronize()
