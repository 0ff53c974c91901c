export GAR_DEBUG_TUNING=1
# RING_MB S TPI PROWS AHEAD
for cfg in "40 1024 1 128 1" "40 2048 1 128 1" "40 1024 1 64 1" "40 1024 2 64 1" "40 2048 1 64 1" "40 1024 1 128 0" "40 1024 1 64 2"; do
  set -- $cfg
  echo "RING_MB=$1 S=$2 TPI=$3 PROWS=$4 AHEAD=$5"
  export GAR_CHAIN_RING_MB=$1 GAR_CHAIN_S=$2 GAR_CHAIN_TPI=$3 GAR_CHAIN_PROWS=$4 GAR_CHAIN_AHEAD=$5
  timeout 120 python tools/bench_chain.py --reps 5 --only C1 2>&1 | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['chain_kernel']['device_ms'], r['two_launches']['device_ms'], r['chain_kernel']['kernels'][0])"
done
