python -m pytest tests/test_gpu_duo.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
one() { python bench.py "$@" --no-e2e --no-configs --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'][:24], d['ms_per_step'], d['value'], d['roofline']['frac'])"; }
for cfg in cfg3 cfg4 cfg5; do one --config $cfg; done
echo "== quad wave test"
one --config cfg2 --sites 9699328
one --config cfg2 --sites 10000000
one --config cfg2 --sites 10911744
echo "== transport timeline (fresh)"
PHASES_ONLY="expanded / fresh" ELX_TRANSPORT_DEBUG=1 python tools/bench_e2e_phases.py 10000000 2 cfg2
echo "== 15 threads"
PHASES_ONLY="expanded / fresh" ELX_HOST_THREADS=15 python tools/bench_e2e_phases.py 10000000 3 cfg2
PHASES_ONLY="expanded / fresh" ELX_HOST_THREADS=16 python tools/bench_e2e_phases.py 10000000 3 cfg2
PHASES_ONLY="expanded / pinned" ELX_HOST_THREADS=15 python tools/bench_e2e_phases.py 10000000 3 cfg2
