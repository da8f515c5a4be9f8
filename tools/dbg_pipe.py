import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import synth
from point_teacher_b200.mil_head import MILHead
from point_teacher_b200.refine import Phase2Pipeline, CapturedPhase2
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
small = len(sys.argv) > 2
dev = torch.device("cuda")
d = synth.hbb_batch(seed=0, **(dict(batch=2, img_hw=(256, 256), gt_range=(6, 10), n_neg=20) if small else {}))
head = MILHead(num_classes=8, num_stages=1, top_k=1, precision=prec).to(dev)
to = lambda l: [t.to(dev) for t in l]
pin = lambda l: [t.pin_memory() for t in l]
host = dict(feat=d["feat"].pin_memory(), pseudo_boxes=pin(d["pseudo_boxes"]), pseudo_points=pin(d["pseudo_points"]),
            pseudo_labels=pin(d["pseudo_labels"]), gt_boxes=pin(d["gt_boxes"]), neg_boxes=[pin(d["neg_boxes"][0])])
inputs = dict(feat=d["feat"].to(dev), pseudo_boxes=to(d["pseudo_boxes"]), pseudo_points=to(d["pseudo_points"]),
              pseudo_labels=to(d["pseudo_labels"]), gt_boxes=to(d["gt_boxes"]), neg_boxes=[to(d["neg_boxes"][0])])
cap = CapturedPhase2(head, inputs, d["img_metas"], synth.HBB_FINE_CFG, synth.HBB_EXT_CFG, num_stages=1, cap=100)
for _ in range(3):
    cap.replay()
torch.cuda.synchronize(); print("cap ok")
pipe = Phase2Pipeline(head, inputs, d["img_metas"], synth.HBB_FINE_CFG, synth.HBB_EXT_CFG, num_stages=1, cap=100)
torch.cuda.synchronize(); print("pipe built")
for i in range(6):
    t = pipe.submit(host)
    torch.cuda.synchronize(); print("submit", i, "ok")
print(pipe.result(t)[2])
