# A/B of the split-Philox builds (parity first), then the host-transport unit-order experiment
python -m pytest tests/test_gpu_parity.py tests/test_gpu_duo.py -m gpu -x -q 2>&1 | tail -3
for lib in "" elynx_b200/libelx_var_auto.so; do
  echo "== lib=$lib"
  for cfg in cfg2 cfg3 cfg5; do
    ELX_LIB=$lib python bench.py --config $cfg --no-e2e --no-configs --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'][:20], d['ms_per_step'], d['value'], d['roofline']['frac'])"
  done
done
bash tools/exp_e1.sh
