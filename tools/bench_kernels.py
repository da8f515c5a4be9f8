#!/usr/bin/env python
"""Per-kernel microbenchmarks on the config-#1 shapes (CUDA events, median of N, optional L2 flush).
Usage: python tools/bench_kernels.py [roi|gemm|heads|all] [--iters 20] [--no-flush]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import ops, synth  # noqa: E402


def timeit(fn, iters, flush):
    buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush else None
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3   # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-flush", action="store_true")
    a = ap.parse_args()
    flush = not a.no_flush
    dev = torch.device("cuda")
    d = synth.hbb_batch(seed=0)
    res = {}
    if a.what in ("roi", "all"):
        from point_teacher_b200.proposals import boxes_to_rois, fine_proposals_from_cfg
        props, _ = fine_proposals_from_cfg([b[:100].to(dev) for b in d["pseudo_boxes"]], synth.HBB_EXT_CFG[0], d["img_metas"])
        rois = boxes_to_rois(props)
        negs = boxes_to_rois([n.to(dev) for n in d["neg_boxes"][0]])
        rois_all = torch.cat([rois, negs]).contiguous()
        x = d["feat"].to(dev)
        f32, b16 = ops.nchw_to_nhwc(x), ops.nchw_to_nhwc(x, torch.bfloat16)
        res["nchw_to_nhwc_f32_us"] = timeit(lambda: ops.nchw_to_nhwc(x), a.iters, flush)
        K = rois.shape[0]
        outb = torch.empty((rois_all.shape[0], 12544), dtype=torch.bfloat16, device=dev)
        outf = torch.empty((rois_all.shape[0], 256, 7, 7), dtype=torch.float32, device=dev)
        for name, feat, mode, o, r in [("roi_f32in_bf16out_pos5000", f32, 0, outb, rois),
                                       ("roi_bf16in_bf16out_pos5000", b16, 0, outb, rois),
                                       ("roi_f32in_f32nchw_pos5000", f32, 1, outf, rois),
                                       ("roi_f32in_bf16out_pos+neg5400", f32, 0, outb, rois_all),
                                       ("roi_f32in_bf16out_neg400", f32, 0, outb, negs)]:
            t = timeit(lambda: ops.roi_align_forward(feat, r, mode, 0.125, out=o), a.iters, flush)
            nb = r.shape[0] * 12544 * (4 if mode == 1 else 2) + feat.numel() * feat.element_size() + r.numel() * 4
            res[name] = {"us": t, "GBps": nb / t / 1e3}
        # stress config #4 (one image): 1500 GT x 64 instances = 96 000 RoIs, 2.4 GB bf16 out
        ds = synth.hbb_batch(seed=1, batch=1, gt_range=(1500, 1500))
        props, _ = hbb.fine_proposals(ds["pseudo_boxes"], synth.stress_ext_cfg(8)[0], ds["img_metas"])
        rs = hbb.bbox2roi(props).to(dev)
        xs = ds["feat"].to(dev)
        outs = torch.empty((rs.shape[0], 12544), dtype=torch.bfloat16, device=dev)
        for name, feat in [("roi_stress96k_f32in", ops.nchw_to_nhwc(xs)), ("roi_stress96k_bf16in", ops.nchw_to_nhwc(xs, torch.bfloat16))]:
            t = timeit(lambda: ops.roi_align_forward(feat, rs, 0, 0.125, out=outs), max(a.iters // 4, 3), flush)
            nb = rs.shape[0] * 12544 * 2 + feat.numel() * feat.element_size() + rs.numel() * 4
            res[name] = {"us": t, "GBps": nb / t / 1e3, "K": rs.shape[0]}
        del outs
    if a.what in ("gemm", "all"):
        for M, N, K in [(5000, 1024, 12544), (5400, 1024, 12544), (5000, 1024, 1024), (4736, 1024, 12544)]:
            A = torch.randn(M, K, device=dev).to(torch.bfloat16)
            Bm = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
            bias = torch.zeros(N, device=dev)
            out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            for split in (True, False):
                t = timeit(lambda: ops.fc_gemm(A, Bm, bias, relu=True, out=out, allow_split=split), a.iters, flush)
                res[f"gemm_{M}x{N}x{K}_split{int(split)}"] = {"us": t, "TFLOPs": 2.0 * M * N * K / t / 1e6}
    if a.what in ("heads", "all"):
        M = 5400
        H = torch.randn(M, 1024, device=dev).to(torch.bfloat16)
        Wc, Wi = torch.randn(8, 1024, device=dev) * 0.01, torch.randn(8, 1024, device=dev) * 0.01
        bz = torch.zeros(8, device=dev)
        res["cls_ins_us"] = timeit(lambda: ops.cls_ins_heads(H, Wc, bz, Wi, bz), a.iters, flush)
        w = torch.randn(1024, 12544, device=dev) * 0.01
        res["prep_fc1_us"] = timeit(lambda: ops.prep_fc1_weight(w, 256), a.iters, flush)
        w2 = torch.randn(1024, 1024, device=dev) * 0.01
        res["cast_w_us"] = timeit(lambda: ops.cast_weight(w2), a.iters, flush)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
