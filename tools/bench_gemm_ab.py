#!/usr/bin/env python
"""A/B of the single-CTA vs paired (cta_group::2) FC GEMM on wave-exact shapes.  PTB200_GEMM_PAIR=0|1."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import ops
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for M, N, K in [(18944, 1024, 12544), (5000, 1024, 12544), (5400, 1024, 12544), (96000, 1024, 12544)]:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
    bias = torch.zeros(N, device=dev)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        ops.fc_gemm(A, B, bias, relu=True, out=out)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.fc_gemm(A, B, bias, relu=True, out=out); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = ts[len(ts) // 2] * 1e-3
    res[f"{M}x{N}x{K}"] = {"us": t * 1e6, "TFLOPs": 2.0 * M * N * K / t / 1e12}
    del A, out
print(os.environ.get("PTB200_GEMM_PAIR", "1"), json.dumps(res))
