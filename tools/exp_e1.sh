set -x
export PHASES_ONLY="expanded / fresh"
for cfg in "" "ELX_UNPACK_ORDER=1" "ELX_UNPACK_COLS=1048576" "ELX_UNPACK_COLS=1048576 ELX_UNPACK_NO_NT=1" "ELX_UNPACK_ORDER=1 ELX_UNPACK_NO_NT=1" "ELX_HOST_BATCH_LOG2=22 ELX_UNPACK_COLS=4194304" "ELX_HOST_BATCH_LOG2=22 ELX_UNPACK_COLS=4194304 ELX_UNPACK_NO_NT=1" "ELX_HOST_BATCH_LOG2=21 ELX_UNPACK_COLS=2097152" ; do
  echo "== $cfg"
  env $cfg python tools/bench_e2e_phases.py 10000000 4 cfg2
done
export PHASES_ONLY="expanded / pinned"
for cfg in "" "ELX_UNPACK_ORDER=1" "ELX_UNPACK_COLS=1048576"; do
  echo "== pinned $cfg"
  env $cfg python tools/bench_e2e_phases.py 10000000 4 cfg2
done
nproc; lscpu | head -30; cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag
