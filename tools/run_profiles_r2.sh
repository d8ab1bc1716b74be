#!/bin/bash
# The measurement pass behind profiles/r2_*: GPU test suite, default bench, reference arm, ncu launch lists per configuration, full
# ncu captures of the hot kernels (run under gpurun from the repo root; writes into gpurun_out/ with the prefix given as $1).
P=${1:-r2z}
python -m pytest tests -m gpu -x -q > gpurun_out/${P}_pytest.log 2>&1; tail -2 gpurun_out/${P}_pytest.log
python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; tail -c 300 gpurun_out/${P}_bench.err
python bench.py --impl reference > gpurun_out/${P}_ref.json 2>/dev/null
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-configs --no-cpu-baseline"
for c in cfg3 cfg4 cfg5; do $B --config $c > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${P}_launches_$c.csv $B --config $c > /dev/null 2>&1; done
D="python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-e2e-pinned"
$D > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches_bench_default_cfg2.csv $D > /dev/null 2>&1
$B --config cfg2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k2_quad_tiled' -s 2 -c 1 -o gpurun_out/${P}_full_cfg2 $B --config cfg2 > gpurun_out/${P}_ncu_cfg2.log 2>&1
for c in cfg3 cfg4 cfg5; do $B --config $c > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k2_duo_tiled|k_duo_unpermute' -s 2 -c 2 -o gpurun_out/${P}_full_$c $B --config $c > gpurun_out/${P}_ncu_$c.log 2>&1; done
ls -la gpurun_out | grep ${P}
