#!/bin/bash
# Multi-GPU measurements (run under `gpurun --gpus N` from the repo root): weak and strong
# scaling of the headline step, the config-5 boxed multi-input shape, and the
# ImitationLearner.step loop (il_exp-shaped training: RMSprop, warm start, theta changes
# every step) at N GPUs.
N=$1; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --single-dtype > $O/r2s_weak_${N}gpu.json 2> $O/r2s_${N}.err
$TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --single-dtype --scaling strong > $O/r2s_strong_${N}gpu.json 2>> $O/r2s_${N}.err
if [ "$N" != "8" ]; then   # the 8-GPU call is charged 8x: headline weak / strong + the learner loop only
$TR bench.py --gpus $N --config lindx --ns 8 --nc 4 --boxed --batch 65536 --steps 3 --warmup 1 > $O/r2s_lindx84_${N}gpu.json 2>> $O/r2s_${N}.err
$TR bench.py --gpus $N --config lindx --ns 4 --nc 2 --boxed --batch 1048576 --scaling strong --steps 3 --warmup 1 > $O/r2s_lindx42_1M_strong_${N}gpu.json 2>> $O/r2s_${N}.err
fi
$TR tools/il_learner_bench.py > $O/r2s_learner_${N}gpu.json 2>> $O/r2s_${N}.err
if [ "$N" = "2" ]; then python -m pytest tests/test_dist_gloo.py -m gpu -q > $O/r2s_dist_tests.txt 2>&1; fi
tail -c 400 $O/r2s_weak_${N}gpu.json; echo; tail -c 300 $O/r2s_learner_${N}gpu.json; echo; tail -3 $O/r2s_${N}.err
