import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import synth
from point_teacher_b200.mil_head import MILHead
from point_teacher_b200.refine import Phase2Pipeline, CapturedPhase2
prec, mode = sys.argv[1], sys.argv[2]
dev = torch.device("cuda")
d = synth.hbb_batch(seed=0, batch=2, img_hw=(256, 256), gt_range=(6, 10), n_neg=20)
head = MILHead(num_classes=8, num_stages=1, top_k=1, precision=prec).to(dev)
to = lambda l: [t.to(dev) for t in l]
def mk():
    return dict(feat=d["feat"].to(dev), pseudo_boxes=to(d["pseudo_boxes"]), pseudo_points=to(d["pseudo_points"]),
              pseudo_labels=to(d["pseudo_labels"]), gt_boxes=to(d["gt_boxes"]), neg_boxes=[to(d["neg_boxes"][0])])
caps = []
n = {"one": 1, "two": 2, "three": 3}[mode]
for i in range(n):
    caps.append(CapturedPhase2(head, mk(), d["img_metas"], synth.HBB_FINE_CFG, synth.HBB_EXT_CFG, num_stages=1, cap=100))
    torch.cuda.synchronize(); print("captured", i, flush=True)
for r in range(4):
    for i, c in enumerate(caps):
        c.replay(); torch.cuda.synchronize(); print("replay", r, i, "ok", flush=True)
