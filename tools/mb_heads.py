"""Micro-benchmark of the small-head kernels (fc_reg: 4 outputs; fc_cls + fc_ins: 2 x 8) on the bench's hidden
activation shape, warm L2 (H was just written), median of 20."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M, D, C = 5400, 1024, 8
H = torch.randn(M, D, device=dev).relu().to(torch.bfloat16)
Wc, Wi, Wr = (torch.randn(C, D, device=dev) * 0.01 for _ in range(2)), None, torch.randn(4, D, device=dev) * 0.01
Wc, Wi = Wc
bc, bi, br = torch.zeros(C, device=dev), torch.zeros(C, device=dev), torch.zeros(4, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(fn, tag):
    ts = []
    for i in range(25):
        flush.zero_()
        H.add_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(100000)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{tag}: {ts[len(ts) // 2]:.1f} us")


run(lambda: ops.cls_ins_heads(H, Wc, bc, Wi, bi, M=M), "cls+ins heads (16 outputs)")
run(lambda: ops.cls_ins_heads(H, Wr, br, Wr, br, M=M), "two 4-output heads (8 outputs)")
run(lambda: torch.empty(1, device=dev), "empty (event overhead)")
