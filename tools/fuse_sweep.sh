for cfg in "0 0 4" "1 4 4" "1 8 4" "1 16 4" "1 2 4" "1 8 2" "1 4 2" "1 8 1"; do
  set -- $cfg
  r=$(ESR_FUSE_RDB=$1 ESR_RDB_CHUNK=$2 ESR_RDB_LAYERS=$3 python bench.py --steps 10 --warmup 3 --no-zopt --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms %.3f conv_frac %.4f clocks %s' % (b['ms_per_step'], b['roofline']['frac'], b['clocks']['sm_mhz']))")
  echo "fuse=$1 chunk=$2 layers=$3 -> $r"
done
