ELX_DUO_PIPELINE=1 python -m pytest tests/test_gpu_duo.py tests/test_gpu_configs.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
one() { python bench.py "$@" --no-e2e --no-configs --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'][:24], d['ms_per_step'], d['value'], d['roofline']['frac'])"; }
for cfg in cfg3 cfg4 cfg5; do one --config $cfg; done
python tools/fuzz_duo.py 60 7 2>&1 | tail -3; ELX_DUO_PIPELINE=1 python tools/fuzz_duo.py 60 8 2>&1 | tail -3
