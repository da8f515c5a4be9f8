"""FC1-shaped GEMM (N=1024, K=12544) over M: exact waves (148 tiles = 4736 rows) vs the bench shapes, to separate the
per-wave time from the fixed cost and the tail.  CUDA events, L2 flushed, median of 15."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
N, K = 1024, 12544
W = (torch.randn(N, K, device=dev) * 0.01).to(torch.bfloat16)
b = torch.zeros(N, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for M in (128, 2368, 4736, 5000, 5400, 9472, 14208, 18944, 37888):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    out = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    for split in (1, 0):
        ts = []
        for i in range(20):
            flush.zero_()
            A.add_(0)                                     # A was "just written" (L2-warm as far as it fits), like in the step
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(100000)
            e0.record()
            ops.fc_gemm(A, W, b, relu=True, out=out, allow_split=bool(split))
            e1.record()
            torch.cuda.synchronize()
            if i >= 5:
                ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        med = ts[len(ts) // 2]
        tiles = (M + 127) // 128 * 4
        print(f"M={M:6d} tiles={tiles:5d} waves={tiles / 148:5.2f} split={split} {med:8.1f} us  {2.0 * M * N * K / med / 1e6:7.0f} TFLOP/s")
