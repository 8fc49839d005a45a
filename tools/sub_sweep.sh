# images per pass x fused growth launches x streams: ms/step, conv fraction, SM clock (bench.py, config 2)
for cfg in "16 0 1" "16 1 1" "8 0 1" "8 1 1" "4 1 1" "8 1 2" "4 1 2" "4 0 2" "2 1 2" "8 0 2"; do
  set -- $cfg
  r=$(ESR_SUBBATCH=$1 ESR_FUSE_RDB=$2 ESR_STREAMS=$3 python bench.py --steps 20 --warmup 3 --no-zopt --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms %.3f e2e %.1f conv_frac %.4f clocks %s' % (b['ms_per_step'], b['e2e']['value'], b['roofline']['frac'], b['clocks']['sm_mhz']))")
  echo "sub=$1 fuse=$2 streams=$3 -> $r"
done
