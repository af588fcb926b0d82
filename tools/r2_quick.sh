#!/bin/bash
# Quick GPU pass: parity tests, smoke, the default bench line.
O=gpurun_out; P=${1:-r2q}
python -m pytest tests -m gpu -q --timeout 600 > $O/${P}_gputests.txt 2>&1; tail -12 $O/${P}_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/${P}_smoke.txt 2>&1; tail -2 $O/${P}_smoke.txt
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${P}_bench.json 2> $O/${P}_bench.err; cut -c1-600 $O/${P}_bench.json
if [ -n "$2" ]; then   # second argument: also capture the dominant kernel with ncu (source + stalls)
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:ilqr_iter_kernel --launch-skip 13 --launch-count 1 -o $O/${P}_iter -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --single-dtype --no-ncu > /dev/null 2>&1
  python tools/ncu_summary.py $O/${P}_iter.ncu-rep > $O/${P}_iter.txt 2>&1
  ncu -i $O/${P}_iter.ncu-rep --page source --csv --print-source cuda,sass > $O/${P}_iter_cs.csv 2>/dev/null
  rm -f $O/${P}_iter.ncu-rep
fi
