"""In-step device time of the training step per C-ABI entry point (eager step behind a device-side sleep, CUDA events
on the launching stream): python tools/trace_train.py [obb]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import _lib, synth  # noqa: E402
from point_teacher_b200.mil_head import MILHead, RotatedMILHead  # noqa: E402
from point_teacher_b200.train import Phase2Trainer  # noqa: E402

rot = len(sys.argv) > 1 and sys.argv[1] == "obb"
dev = torch.device("cuda:0")
d = synth.obb_batch(seed=0) if rot else synth.hbb_batch(seed=0)
torch.manual_seed(0)
head = (RotatedMILHead(num_classes=9, num_stages=1, top_k=3) if rot else MILHead(num_classes=8, num_stages=1, top_k=1)).to(dev)
to = lambda l: [t.to(dev) for t in l]  # noqa: E731
fine, ext = (synth.OBB_FINE_CFG, synth.OBB_EXT_CFG) if rot else (synth.HBB_FINE_CFG, synth.HBB_EXT_CFG)
tr = Phase2Trainer(head, fine, ext, num_stages=1, cap=100)
x = d["feat"].to(dev).requires_grad_(True)
args = (d["img_metas"], to(d["pseudo_boxes"]), to(d["pseudo_points"]), to(d["pseudo_labels"]), to(d["gt_boxes"]))
negs = [to(d["neg_boxes"][0])]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    x.grad = None
    tr.step((x,), *args, neg_boxes=negs, reduce_logs=False)
torch.cuda.synchronize()
_lib.TRACE["on"], _lib.TRACE["events"] = True, []
n = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for _ in range(n):
    x.grad = None
    flush.zero_()
    torch.cuda._sleep(int(12e-3 * 1.9e9))
    e0.record()
    tr.step((x,), *args, neg_boxes=negs, reduce_logs=False)
    e1.record()
    torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
_lib.TRACE["on"] = False
agg, cnt = {}, {}
for name, a, b in _lib.TRACE["events"]:
    agg[name] = agg.get(name, 0.0) + a.elapsed_time(b) * 1e3 / n
    cnt[name] = cnt.get(name, 0) + 1
print(f"step (eager, events around the whole step incl. torch glue): {tot / n * 1e3:.0f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"{k:36s} {cnt[k] // n:3d} launches {v:8.1f} us")
print(f"sum {sum(agg.values()):.0f} us")
