"""Multi-GPU plumbing of the phase-2 path: one process per GPU, images sharded across ranks,
NO data-path collective (bags, RoIs, scores and selection are independent per image; an image's RoIs only read that
image's feature map).  The only exchange steps are
  * the mean of the logged loss scalars across ranks, as mmdet's ``BaseDetector._parse_losses`` does
    (HBB_TOD/mmdet/models/detectors/base.py: ``dist.all_reduce(loss_value.div_(dist.get_world_size()))``), and
  * one flat-bucket all-reduce (average) of the MIL-head parameter gradients per step (DDP semantics of
    HBB_TOD/tools/train.py -> mmdet.apis.train_detector -> MMDistributedDataParallel), restricted to the parameters
    the path actually uses (``shared_fcs_reg``, ``shared_fcs_bag``, ``fc_cls``, ``fc_ins``, ``fc_reg``): the
    constructed-but-unused ``shared_fcs``, ``shared_fcs_refine`` and ``fc_iou`` never receive a gradient.
Per-rank loss denominators (``avg_factor = K_local``, ``num_sample_local``) are kept per process exactly like the
reference under DDP.  Backend: NCCL over NVLink on the GPUs, gloo on CPU (tests)."""
import os

import torch
import torch.distributed as dist

USED_PREFIXES = ("shared_fcs_reg.", "shared_fcs_bag.", "fc_cls.", "fc_ins.", "fc_reg.")


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def image_shard(num_images, rank_=None, world_=None):
    """Contiguous, balanced partition of image indices: the first ``num_images % world`` ranks own one more."""
    r = rank() if rank_ is None else rank_
    w = world() if world_ is None else world_
    base, extra = divmod(num_images, w)
    start = r * base + min(r, extra)
    return list(range(start, start + base + (1 if r < extra else 0)))


def reduce_mean_losses(losses, group=None):
    """Mean over ranks of every logged scalar (one collective on a stacked vector); identity for one process."""
    if world() == 1:
        return dict(losses)
    keys = sorted(losses)
    vec = torch.stack([losses[k].detach().reshape(()).float() for k in keys])
    dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    vec /= world()
    return {k: vec[i] for i, k in enumerate(keys)}


def _all_reduce_avg_(t, group=None, async_op=False):
    """In-place mean over ranks.  NCCL averages inside the collective (ncclAvg); gloo (CPU tests) has no AVG."""
    if dist.get_backend(group) == "nccl":
        return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
    t /= world()
    return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class MILGradBucket:
    """The gradients of the MIL-head parameters the path uses, laid out so that the backward kernels write them IN
    PLACE and each branch can be reduced the moment its last kernel is enqueued.

    ``flat`` (fp32, 27.8 M values = 111 MB per stage) is cut into one contiguous group per (stage, branch):

        reg: [fc_reg.W | fc_reg.b | reg.0.b | reg.1.b | pad]  [reg.1.W]  [reg.0.W]
        bag: [fc_cls.W fc_ins.W | fc_cls.b fc_ins.b | bag.0.b | bag.1.b | pad]  [bag.1.W]  [bag.0.W]
             '---- "small" region: accumulated with atomics, zeroed per step ----'  '-- written whole by the wgrad GEMMs --'

    * direct mode (``train.Phase2Trainer``, the hand-written backward): ``targets(stage)`` hands these views to the
      kernels -- the FC1 weight gradient lands in the operand's bin-major column order straight from the GEMM
      epilogue, ``reduce_group_(stage, branch)`` launches that group's all-reduce (average; asynchronous on NCCL's
      stream, so it runs under the other branch's GEMMs / the RoIAlign backward) and ``finish_()`` waits, un-permutes
      the two FC1 gradients into the parameters' ``c*49 + bin`` order and publishes ``param.grad`` (views of ``flat``
      for everything else).  No pack / unpack copies, no separate 1/world pass.
    * legacy mode (``pack_`` / ``all_reduce_`` / ``unpack_``): gradients that autograd left in ``param.grad`` are
      copied in (parameter order for every entry), reduced with one collective and copied back (``loss.backward()``
      flows, gloo tests)."""

    def __init__(self, head, prefixes=USED_PREFIXES):
        params = dict(head.named_parameters())
        stages = sorted({int(n.split(".")[1]) for n in params if n.startswith("fc_reg.")})
        self.groups, order = [], []            # groups: dict(stage, branch, names, lo, hi, small_hi)
        off = 0

        def pad4(x):
            return (x + 3) // 4 * 4
        self.offsets = {}
        for s in stages:
            for branch, heads in (("reg", [f"fc_reg.{s}"]), ("bag", [f"fc_cls.{s}", f"fc_ins.{s}"])):
                fcs = f"shared_fcs_{branch}.{s}"
                names = [h + ".weight" for h in heads] + [h + ".bias" for h in heads] + \
                        [f"{fcs}.0.bias", f"{fcs}.1.bias", f"{fcs}.1.weight", f"{fcs}.0.weight"]
                names = [n for n in names if n in params and n.startswith(prefixes) and params[n].requires_grad]
                lo = off
                for n in names:
                    if n.endswith((".1.weight", ".0.weight")) and n.startswith("shared_fcs"):
                        off = pad4(off)
                    if n == f"{fcs}.1.weight":
                        small_hi = off
                    self.offsets[n] = off
                    off += params[n].numel()
                off = pad4(off)
                self.groups.append(dict(stage=s, branch=branch, names=names, lo=lo, hi=off, small_hi=small_hi))
                order += names
        self.named = [(n, params[n]) for n in order]
        if not self.named:
            raise ValueError("no MIL-head parameters found")
        self.numel = sum(p.numel() for _, p in self.named)
        dev = self.named[0][1].device
        self.flat = torch.zeros((off,), dtype=torch.float32, device=dev)
        self.views = [self.flat[self.offsets[n]:self.offsets[n] + p.numel()].view_as(p) for n, p in self.named]
        self._view = {n: v for (n, _), v in zip(self.named, self.views)}
        self._w1_grad = {}                     # persistent parameter-order gradient of the two FC1 weights per stage
        self._works = []
        self._forks = []
        self._w1_done = set()
        self.head = head
        # SMs the persistent GEMMs leave to NCCL while a reduction of this bucket is in flight (0 = none);
        # PTB200_NCCL_SM_RESERVE overrides it for measurements (tools/mb_train_dist.py)
        # measured (training step, max over ranks): N = 8 1.61-1.64 ms with 0, 1.55 / 1.56 / 1.58 / 1.52 ms with
        # 16 / 24 / 32 / 40; N = 2 1.43 ms with 0-16, 1.45-1.46 ms with 24-40
        self.sm_reserve = int(os.environ.get("PTB200_NCCL_SM_RESERVE", "16" if world() >= 4 else "0"))

    def names(self):
        return [n for n, _ in self.named]

    # ---------------------------------------------------------------- direct mode
    def targets(self, stage):
        """Views the backward kernels of ``stage`` write into: {'reg'|'bag': dict(Wh, bh, b1, b2, W2, W1p)}."""
        out = {}
        v = self._view
        for branch, heads in (("reg", [f"fc_reg.{stage}"]), ("bag", [f"fc_cls.{stage}", f"fc_ins.{stage}"])):
            fcs = f"shared_fcs_{branch}.{stage}"
            o0 = self.offsets[heads[0] + ".weight"]
            rows = sum(v[h + ".weight"].shape[0] for h in heads)
            D = v[heads[0] + ".weight"].shape[1]
            ob = self.offsets[heads[0] + ".bias"]
            out[branch] = dict(Wh=self.flat[o0:o0 + rows * D].view(rows, D), bh=self.flat[ob:ob + rows],
                               b1=v[f"{fcs}.0.bias"], b2=v[f"{fcs}.1.bias"], W2=v[f"{fcs}.1.weight"],
                               W1p=v[f"{fcs}.0.weight"])
        return out

    def zero_small_(self):
        for g in self.groups:
            self.flat[g["lo"]:g["small_hi"]].zero_()

    def reduce_group_(self, stage, branch, group=None):
        """Called the moment the last parameter-gradient kernel of (stage, branch) is enqueued.  The FC1 weight
        gradient (51 MB, bin-major in the bucket) is put into parameter order on the forked stream, under the rest of
        the backward; with more than one rank that copy and the remainder of the group's slice are averaged over the
        ranks by two asynchronous all-reduces (joined by ``finish_``)."""
        from . import ops
        g = next(x for x in self.groups if x["stage"] == stage and x["branch"] == branch)
        w1 = f"shared_fcs_{branch}.{stage}.0.weight"
        lo, hi = g["lo"], g["hi"]
        # PTB200_REDUCE_SCHEME=split|group overrides: 'split' un-permutes on the forked stream and reduces that copy on
        # its own; 'group' reduces the whole slice with one all-reduce and un-permutes in finish_()
        scheme = os.environ.get("PTB200_REDUCE_SCHEME", "split" if world() == 1 else "group")
        if scheme == "split" and w1 in self.offsets and self.flat.is_cuda:
            p = dict(self.named)[w1]
            buf = self._w1_grad.get(w1)
            if buf is None:
                buf = self._w1_grad[w1] = torch.empty_like(p)
            with ops.fork() as f:
                ops.unpermute_dw1(self._view[w1], self.head.in_channels, self.head.roi_feat_area, buf, accumulate=False)
                if world() > 1:     # issued from the forked stream: NCCL orders itself behind the un-permute
                    self._works.append(_all_reduce_avg_(buf.view(-1), group=group, async_op=True))
            self._forks.append(f)
            self._w1_done.add(w1)
            hi = self.offsets[w1]                           # FC1 W is the last entry of the group's slice
        if world() == 1:
            return
        self._works.append(_all_reduce_avg_(self.flat[lo:hi], group=group, async_op=True))
        if self.sm_reserve and self.flat.is_cuda:
            ops.set_gemm_sm_reserve(self.sm_reserve)        # until finish_(): the GEMMs leave NCCL's SMs alone

    def finish_(self):
        """Join the forked un-permutes and the outstanding reductions, publish ``.grad``."""
        from . import ops
        for f in self._forks:
            f.join()
        self._forks = []
        for w in self._works:
            w.wait()
        self._works = []
        if self.sm_reserve and self.flat.is_cuda:
            ops.set_gemm_sm_reserve(0)
        head = self.head
        for (n, p), v in zip(self.named, self.views):
            if n.startswith("shared_fcs") and n.endswith(".0.weight"):
                g = self._w1_grad.get(n)
                if g is None:
                    g = self._w1_grad[n] = torch.empty_like(p)
                if n not in self._w1_done:      # a branch whose reduce_group_ was not called (direct callers)
                    ops.unpermute_dw1(v, head.in_channels, head.roi_feat_area, g, accumulate=False)
                p.grad = g
            else:
                p.grad = v
        self._w1_done = set()
        return self

    # ---------------------------------------------------------------- legacy mode
    def pack_(self):
        for (_, p), v in zip(self.named, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
        return self.flat

    def all_reduce_(self, group=None, async_op=False):
        self.pack_()
        if world() == 1:
            return self.unpack_()
        work = _all_reduce_avg_(self.flat, group=group, async_op=async_op)
        if async_op:
            return work
        return self.unpack_()

    def unpack_(self):
        for (_, p), v in zip(self.named, self.views):
            if p.grad is None:
                p.grad = v.clone()
            elif p.grad.data_ptr() != v.data_ptr():
                p.grad.copy_(v)
        return self
