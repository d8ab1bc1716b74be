one() { python bench.py "$@" --no-e2e --no-configs --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'][:24], d['ms_per_step'], d['value'], d['roofline']['frac'])"; }
export ELX_DUO_PIPELINE=0
for lib in "" elynx_b200/libelx_var_fma.so; do echo "== $lib"; for cfg in cfg3 cfg4 cfg5; do ELX_LIB=$lib one --config $cfg; done; done
ELX_LIB=elynx_b200/libelx_var_fma.so python -m pytest tests/test_gpu_duo.py -m gpu -x -q 2>&1 | tail -2
