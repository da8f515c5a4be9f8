import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hbb
from point_teacher_b200 import synth
from point_teacher_b200.mil_head import MILHead
from point_teacher_b200.refine import phase2_refine
SMALL = dict(batch=2, img_hw=(256, 256), gt_range=(6, 10), n_neg=20)
cuda = torch.device("cuda")
seed, alpha = 0, (1.0, 1.0)
d = synth.hbb_batch(seed=seed, **SMALL)
P = hbb.MilHeadParams(num_stages=1, seed=seed).requires_grad_(True)
feat = d["feat"].clone().requires_grad_(True)
ob, op, ol, aux = hbb.phase2_refine(P, (feat,), [d["stride"]], d["img_metas"], d["pseudo_boxes"], d["pseudo_points"],
                                    d["pseudo_labels"], d["gt_boxes"], synth.HBB_FINE_CFG, synth.HBB_EXT_CFG,
                                    num_stages=1, alpha=alpha, topk=1, injected_negs=d["neg_boxes"])
(ol["stage0_loss_mil_bbox"] + ol["stage0_loss_mil_bags"]).backward()
ref = {k: v.grad for k, v in P.state_dict().items()}
head = MILHead(num_classes=8, num_stages=1, top_k=1, precision="bf16").to(cuda)
head.load_state_dict({k: v.detach() for k, v in P.state_dict().items()}, strict=False)
x = d["feat"].to(cuda).requires_grad_(True)
to = lambda l: [t.to(cuda) for t in l]
gb, gp, gl = phase2_refine(head, (x,), d["img_metas"], to(d["pseudo_boxes"]), to(d["pseudo_points"]),
                           to(d["pseudo_labels"]), to(d["gt_boxes"]), synth.HBB_FINE_CFG, synth.HBB_EXT_CFG,
                           num_stages=1, alpha=alpha, neg_boxes=[to(d["neg_boxes"][0])], train=True)
(gl["stage0_loss_mil_bbox"] + gl["stage0_loss_mil_bags"]).backward()
got = dict(head.named_parameters())
def rep(name, a, b):
    a, b = a.double().cpu(), b.double()
    print(f"{name:28s} max-rel {((a-b).abs().max()/b.abs().max()).item():.4f}  fro-rel {((a-b).norm()/b.norm()).item():.4f}  cos {torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), 0).item():.5f}  |ref|max {b.abs().max().item():.3e}")
for k, r in ref.items():
    rep(k, got[k].grad, r)
rep("feat", x.grad, feat.grad)
