#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun from the repo root); everything lands
# in gpurun_out/r2m_*.  Numbers quoted in DESIGN.md / profiles/ come from here.  Every stage has
# its own timeout and is logged with its wall time (gpurun_out/r2m_stages.log).
O=gpurun_out
stage() { local n=$1; local lim=$2; shift 2; local t0=$(date +%s); timeout $lim "$@"; echo "$n rc=$? $(( $(date +%s) - t0 ))s" >> $O/r2m_stages.log; }
: > $O/r2m_stages.log
B="python bench.py"
stage bench_f64 300 bash -c "$B --gpus 1 --steps 20 --warmup 5 > $O/r2m_bench.json 2> $O/r2m_bench.err"
stage bench_f32 200 bash -c "$B --dtype f32 --steps 20 --warmup 5 --no-cpu-baseline --single-dtype --no-ncu > $O/r2m_bench_f32.json 2>> $O/r2m_bench.err"
stage bench_2b 200 bash -c "$B --regime 2b --steps 20 --warmup 5 --no-cpu-baseline --single-dtype --no-ncu > $O/r2m_bench_2b.json 2>> $O/r2m_bench.err"
stage bench_bcast 200 bash -c "$B --broadcast-cost --steps 20 --warmup 5 --no-cpu-baseline --single-dtype --no-ncu > $O/r2m_bench_broadcast.json 2>> $O/r2m_bench.err"
stage rocket_f64 200 bash -c "$B --config rocket --steps 3 --warmup 2 > $O/r2m_rocket_f64.json 2>> $O/r2m_bench.err"
stage rocket_f32 200 bash -c "$B --config rocket --dtype f32 --steps 3 --warmup 2 > $O/r2m_rocket_f32.json 2>> $O/r2m_bench.err"
stage ref_arm 300 bash -c "$B --impl reference --gpus 1 --steps 3 --warmup 1 > $O/r2m_bench_reference_arm.json 2>> $O/r2m_bench.err"
stage fp64lat 60 bash -c "tools/bin/fp64_latency > $O/r2m_fp64_latency.jsonl 2>&1"
# launch list of one bench run (per-launch times are cold-cache / serialised: shares, not absolutes)
stage launches 300 bash -c "ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r2m_launches_bench.csv $B --steps 2 --warmup 1 --no-cpu-baseline --single-dtype --no-ncu > /dev/null 2>&1"
python tools/launch_summary.py $O/r2m_launches_bench.csv > $O/r2m_launches_bench_summary.txt 2>&1
# full capture of the dominant kernel (report kept: source page is read in the build container)
stage ncu_iter 300 bash -c "ncu --set full --clock-control none --import-source on --kernel-name regex:ilqr_iter_kernel --launch-skip 26 --launch-count 1 -o $O/r2m_iter -f $B --steps 1 --warmup 1 --no-cpu-baseline --single-dtype --no-ncu > /dev/null 2>&1"
python tools/ncu_summary.py $O/r2m_iter.ncu-rep > $O/r2m_iter.txt 2>&1
# the other kernels of the step: one report with the second step's launches, summaries only
stage ncu_rest 420 bash -c "ncu --set full --clock-control none --kernel-name 'regex:ilqr_begin|ilqr_gains|lam_tables|adjoint_factor|adjoint_pass|sens_theta|tile_cost|commit' --launch-skip 24 --launch-count 16 -o $O/r2m_rest -f $B --steps 1 --warmup 1 --no-cpu-baseline --single-dtype --no-ncu > /dev/null 2>&1"
python tools/ncu_summary.py $O/r2m_rest.ncu-rep > $O/r2m_rest.txt 2>&1
rm -f $O/r2m_rest.ncu-rep
stage ncu_rocket 300 bash -c "ncu --set full --clock-control none --import-source on --kernel-name regex:group_sweep_kernel --launch-skip 3 --launch-count 1 -o $O/r2m_group_sweep_rocket -f $B --config rocket --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1"
python tools/ncu_summary.py $O/r2m_group_sweep_rocket.ncu-rep > $O/r2m_group_sweep_rocket.txt 2>&1
rm -f $O/r2m_group_sweep_rocket.ncu-rep
ncu -i $O/r2m_iter.ncu-rep --page source --csv --print-source cuda,sass > $O/r2m_iter_cs.csv 2>/dev/null
rm -f $O/r2m_iter.ncu-rep
# config 5: synthetic LinDx sweep on one GPU (forward + KKT backward), B = 65536 and 1 M
if [ -z "$SKIP_SWEEP" ]; then
: > $O/r2m_lindx_sweep.jsonl
for shape in "4 1" "4 2" "8 2" "8 4" "16 4"; do
  set -- $shape
  for boxed in "" "--boxed"; do
    for T in 10 50 200; do
      timeout 120 python bench.py --config lindx --ns $1 --nc $2 --horizon $T $boxed --batch 65536 --steps 2 --warmup 1 2>/dev/null | tail -1 >> $O/r2m_lindx_sweep.jsonl
    done
  done
done
for shape in "4 1 1048576" "4 2 1048576" "8 2 262144"; do   # 8+2 at 1 M would need ~150 GB
  set -- $shape
  timeout 200 python bench.py --config lindx --ns $1 --nc $2 --horizon 50 --boxed --batch $3 --steps 2 --warmup 1 2>/dev/null | tail -1 >> $O/r2m_lindx_sweep.jsonl
done
fi
echo "lindx_sweep done $(wc -l < $O/r2m_lindx_sweep.jsonl) lines" >> $O/r2m_stages.log
cat $O/r2m_stages.log
du -sh $O
