"""N4 at the reference's page size: time of rn_preprocess_pages per 16 pages of 2200 x 1712 (a train of calls between two
events); run it under `ncu --metrics gpu__time_duration.sum --csv` for the three kernels' durations.

    python profiles/time_preprocess.py [reps]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import retinanet_b200 as rn
import synthetic

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
H, W, B = 2200, 1712, 16
pages = np.stack([synthetic.document_page(900 + i, H, W) for i in range(4)])
src = torch.from_numpy(pages).cuda().repeat(B // 4, 1, 1, 1).contiguous()
outs = [torch.empty_like(src) for _ in range(2)]
for i in range(3):
    rn.preprocess.preprocess_pages(src, out=outs[i & 1])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    rn.preprocess.preprocess_pages(src, out=outs[i & 1])
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / reps
print("N4: %d pages of %dx%d in %.1f us = %.0f pages/s; 6 B/pixel -> %.0f GB/s algorithmic" % (B, H, W, us, B / (us * 1e-6), 6.0 * B * H * W / (us * 1e-6) / 1e9))
