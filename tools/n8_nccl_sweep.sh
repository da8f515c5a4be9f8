#!/bin/bash
# N-GPU training-step time under different NCCL footprints / GEMM launch modes (measurement only).
N=${1:-8}
run() {
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/sweep_$tag.json 2> gpurun_out/sweep_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/sweep_{tag}.json").read().splitlines() if l.startswith("{")][-1])
    print(tag, "fwd ms", round(d["ms_per_step"], 4), "train ms", round(d["train_step"]["ms_per_step"], 4), flush=True)
except Exception as e:
    print(tag, "ERR", e, open(f"gpurun_out/sweep_{tag}.err").read()[-600:], flush=True)
PY
}
run base X=1
run ctas8 NCCL_MAX_CTAS=8
run ctas4 NCCL_MAX_CTAS=4
run nocoop PTB200_GEMM_NO_COOP=1
run ctas16 NCCL_MAX_CTAS=16
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep -i "channels\|nvls\|algo\|Connected" | head -20
