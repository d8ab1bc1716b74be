python -m pytest tests/test_gpu_duo.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -2
one() { python bench.py "$@" --no-e2e --no-configs --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'][:24], d['ms_per_step'], d['value'], d['roofline']['frac'])"; }
for cfg in cfg3 cfg4 cfg5; do one --config $cfg; done
echo "== pipeline"
export ELX_DUO_PIPELINE=1
for cfg in cfg3 cfg4 cfg5; do one --config $cfg; done
ELX_DUO_TIMELINE=1 python bench.py --config cfg3 --no-e2e --no-configs --no-cpu-baseline --steps 2 --warmup 1 2>&1 >/dev/null | tail -2
unset ELX_DUO_PIPELINE
ELX_DUO_TIMELINE=1 python bench.py --config cfg3 --no-e2e --no-configs --no-cpu-baseline --steps 2 --warmup 1 2>&1 >/dev/null | tail -2
