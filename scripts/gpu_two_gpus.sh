#!/bin/bash
# round 2, call X (2 GPUs): attention tests, then the 2-GPU bench through torchrun (NCCL), then the 1-GPU line on the same box
OUT=${1:-gpurun_out/r2x}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "attention" > $OUT/attn_tests.log 2>&1; echo "attention tests: $(tail -1 $OUT/attn_tests.log)"
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_full_config.py -q -m gpu --no-header -p no:cacheprovider -k "golden or c3_teacher or c4_ or attention_c4" > $OUT/model_tests.log 2>&1; echo "model tests: $(tail -1 $OUT/model_tests.log)"
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 ) > $OUT/bench_n2.log 2>&1; echo "bench n2 exit $?"; grep "^{" $OUT/bench_n2.log > $OUT/bench_n2.json; cut -c1-420 $OUT/bench_n2.json; tail -3 $OUT/bench_n2.log | cut -c1-300
( timeout 600 python bench.py --no-extras ) > $OUT/bench_n1.log 2>&1; grep "^{" $OUT/bench_n1.log > $OUT/bench_n1.json; cut -c1-200 $OUT/bench_n1.json
