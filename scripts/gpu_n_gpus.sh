#!/bin/bash
# bench.py on N GPUs of one box through torchrun (NCCL); usage: gpurun --gpus N -- 'bash scripts/gpu_n_gpus.sh N'
N=${1:-4}
OUT=gpurun_out/n$N
mkdir -p $OUT
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 ) > $OUT/bench.log 2>&1; echo "bench n$N exit $?"; grep "^{" $OUT/bench.log > $OUT/bench.json; cut -c1-300 $OUT/bench.json; python -c "
import json
d=json.load(open('$OUT/bench.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['by_entry_point_ms'])"
