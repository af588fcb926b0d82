#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun from the repo root); everything lands
# in gpurun_out/r2m_*.  Numbers quoted in DESIGN.md / profiles/ come from here.
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2m_gputests.txt 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2m_bench.json 2> $O/r2m_bench.err
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $O/r2m_bench_reference_arm.json 2>> $O/r2m_bench.err
python bench.py --dtype f32 --steps 20 --warmup 5 --no-cpu-baseline --single-dtype > $O/r2m_bench_f32.json 2>> $O/r2m_bench.err
python bench.py --regime 2b --steps 20 --warmup 5 --no-cpu-baseline > $O/r2m_bench_2b.json 2>> $O/r2m_bench.err
python bench.py --broadcast-cost --steps 20 --warmup 5 --no-cpu-baseline --single-dtype > $O/r2m_bench_broadcast.json 2>> $O/r2m_bench.err
python bench.py --config rocket --steps 3 --warmup 2 > $O/r2m_rocket_f64.json 2>> $O/r2m_bench.err
python bench.py --config rocket --dtype f32 --steps 3 --warmup 2 > $O/r2m_rocket_f32.json 2>> $O/r2m_bench.err
tools/bin/fp64_latency > $O/r2m_fp64_latency.jsonl 2>&1
# launch list of one bench run (per-launch times are cold-cache / serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r2m_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --single-dtype --no-ncu > /dev/null 2>&1
python tools/launch_summary.py $O/r2m_launches_bench.csv > $O/r2m_launches_bench_summary.txt 2>&1
# full captures of the kernels of the headline step
for k in ilqr_iter_kernel ilqr_gains_kernel adjoint_pass_kernel sens_theta_kernel adjoint_factor_kernel lam_tables_kernel ilqr_begin_kernel; do
  ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 3 --launch-count 1 \
      -o $O/r2m_$k -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --single-dtype --no-ncu > /dev/null 2>&1
  python tools/ncu_summary.py $O/r2m_$k.ncu-rep > $O/r2m_$k.txt 2>&1
done
ncu --set full --clock-control none --import-source on --kernel-name regex:group_sweep_kernel --launch-skip 3 --launch-count 1 \
    -o $O/r2m_group_sweep_rocket -f python bench.py --config rocket --steps 1 --warmup 1 > /dev/null 2>&1
python tools/ncu_summary.py $O/r2m_group_sweep_rocket.ncu-rep > $O/r2m_group_sweep_rocket.txt 2>&1
# config 5: synthetic LinDx sweep on one GPU
: > $O/r2m_lindx_sweep.jsonl
for shape in "4 1" "4 2" "8 1" "8 2" "8 4" "16 1" "16 2" "16 4"; do
  set -- $shape
  for boxed in "" "--boxed"; do
    for T in 10 50 200; do
      python bench.py --config lindx --ns $1 --nc $2 --horizon $T $boxed --batch 65536 --steps 2 --warmup 1 2>/dev/null | tail -1 >> $O/r2m_lindx_sweep.jsonl
    done
  done
done
for shape in "4 1" "4 2" "8 2"; do
  set -- $shape
  python bench.py --config lindx --ns $1 --nc $2 --horizon 50 --boxed --batch 1048576 --steps 2 --warmup 1 2>/dev/null | tail -1 >> $O/r2m_lindx_sweep.jsonl
done
echo done > $O/r2m_done
