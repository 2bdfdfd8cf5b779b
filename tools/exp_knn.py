# SPDX-License-Identifier: Apache-2.0
"""kNN timing at BASELINE config-5 shapes (bring-up only)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from warpconvnet_b200 import _ops
for n, nb in ((125000, 1), (1000000, 8), (1000000, 1)):
    pts = torch.rand(n, 3, device="cuda")
    offs = torch.arange(0, n + 1, n // nb, dtype=torch.int64)
    for _ in range(2):
        _ops.knn_search(pts, offs, pts, offs, 16)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        idx = _ops.knn_search(pts, offs, pts, offs, 16)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"knn k=16 n={n} batches={nb}: {ms:.3f} ms  ({n / ms / 1e3:.1f} M queries/s)")
    if n <= 125000:
        t0 = time.time()
        d = torch.cdist(pts[:4096], pts)
        ref = torch.topk(d, 16, dim=1, largest=False).indices
        torch.cuda.synchronize()
        print("   first 4096 rows equal to cdist+topk:", float((ref.sort(1).values == idx[:4096].sort(1).values).float().mean()))
