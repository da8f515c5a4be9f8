"""TEST INFRASTRUCTURE ONLY (dev container, needs /root/reference) -- pins the standalone
restatement in oracle/ against the reference's OWN files executed under oracle/ref_shim.py
on identical seeded inputs.  Run: python -m oracle.check_oracle_vs_ref"""
import sys
import time

import torch

from oracle import hbb, ref_shim
from point_teacher_b200 import synth


def _eq(a, b, name, tol=0.0):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if a.dtype == torch.bool or not a.dtype.is_floating_point:
        ok = torch.equal(a, b)
        err = 0 if ok else 1
    else:
        err = (a.double() - b.double()).abs().max().item() if a.numel() else 0.0
        ok = err <= tol
    print(f"  {'OK ' if ok else 'BAD'} {name}: max|diff|={err:.3g} (tol {tol})")
    return ok


def check_hbb(seed=0, ext_cfg=None, fine_cfg=None, topk=1, stages=1):
    ns = ref_shim.install()
    d = synth.hbb_batch(seed=seed, num_stages=stages)
    fine_cfg = fine_cfg or synth.HBB_FINE_CFG
    ext_cfg = ext_cfg or synth.HBB_EXT_CFG
    cap = 100
    ok = True
    # ---- reference
    head = ref_shim.build_ref_mil_head(ns, num_stages=stages, top_k=topk, seed=seed)
    P = hbb.MilHeadParams(num_stages=stages, seed=seed)
    sd = head.state_dict()
    for k, v in P.state_dict().items():
        assert torch.equal(sd[k], v), k
    pb = [b[:cap].clone() for b in d["pseudo_boxes"]]
    gb = [b[:cap].clone() for b in d["gt_boxes"]]
    pp = [b[:cap].clone() for b in d["pseudo_points"]]
    pl = [b[:cap].clone() for b in d["pseudo_labels"]]
    x = (d["feat"],)
    t0 = time.perf_counter()
    ref_losses = {}
    with torch.no_grad():
        rpb = pb
        for s in range(stages):
            props, valids, refs, reals = ns.syn.MIL_gen_proposals_from_cfg(pp, rpb, fine_cfg[s], gb, d["img_metas"])
            negs = d["neg_boxes"][s]
            nw = [((ns.bbox_overlaps(negs[i], props[i]) < 0.3).sum(1) == props[i].shape[0]) for i in range(len(negs))]
            l, rpb = head.MIL_head_burn_in_step2(x, d["img_metas"], props, valids, refs, reals, negs, nw,
                                                 rpb, pl, ext_cfg[s], s)
            rpb = list(rpb)
            ref_losses.update(l)
    t_ref = time.perf_counter() - t0
    t0 = time.perf_counter()
    with torch.no_grad():
        ob, op, ol, aux = hbb.phase2_refine(P, x, [d["stride"]], d["img_metas"], d["pseudo_boxes"],
                                            d["pseudo_points"], d["pseudo_labels"], d["gt_boxes"], fine_cfg,
                                            ext_cfg, num_stages=stages, cap=cap, alpha=(1.0, 1.0), topk=topk,
                                            injected_negs=d["neg_boxes"])
    t_or = time.perf_counter() - t0
    print(f"seed {seed}: reference {t_ref:.2f}s, oracle {t_or:.2f}s")
    for i in range(len(pb)):
        ok &= _eq(ob[i][:cap], rpb[i], f"refined boxes img{i}", 0.0)
    for k, v in ref_losses.items():
        ok &= _eq(ol[k], v, k, 1e-6)
    return ok


if __name__ == "__main__":
    good = check_hbb(0)
    good &= check_hbb(1, topk=3)
    good &= check_hbb(2, stages=2)
    print("ALL OK" if good else "MISMATCH")
    sys.exit(0 if good else 1)
