python -m pytest tests -m gpu -x -q 2>&1 | tail -5
one() { python bench.py "$@" --no-e2e --no-configs --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'][:24], d['ms_per_step'], d['value'], d['roofline']['frac'])"; }
for cfg in cfg2 cfg3 cfg4 cfg5; do one --config $cfg; done
one --config cfg2 --sites 9699328
