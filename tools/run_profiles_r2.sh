#!/bin/bash
# The measurement pass behind profiles/r2_*: default bench, reference arm, ncu launch lists per configuration, full ncu captures
# of the duo kernels (run under gpurun from the repo root; writes into gpurun_out/).
python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -c 300 gpurun_out/r2q_bench.err
python bench.py --impl reference > gpurun_out/r2q_ref.json 2>/dev/null
for c in cfg3 cfg4 cfg5; do python bench.py --steps 2 --warmup 1 --no-e2e --no-configs --no-cpu-baseline --config $c > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2q_launches_$c.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-configs --no-cpu-baseline --config $c > /dev/null 2>&1; done
python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-e2e-pinned > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches_bench_default_cfg2.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-e2e-pinned > /dev/null 2>&1
for c in cfg3 cfg4; do python bench.py --steps 2 --warmup 1 --no-e2e --no-configs --no-cpu-baseline --config $c > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k2_duo_tiled|k_duo_unpermute' -s 2 -c 2 -o gpurun_out/r2q_full_$c python bench.py --steps 2 --warmup 1 --no-e2e --no-configs --no-cpu-baseline --config $c > gpurun_out/r2q_ncu_$c.log 2>&1; done
ls -la gpurun_out | grep r2q
