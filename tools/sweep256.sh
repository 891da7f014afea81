for cfg in "${1:-220 1}"; do
  set -- $cfg
  echo "=== STEP_NS=$1 XCAP=$2"
  B200AMG_BLOCK_STEP_NS=$1 B200AMG_BLOCK_XCAP=$2 B200AMG_BLOCK_VERBOSE=1 timeout 300 python tools/block_smoke.py --sizes 256 --no-oracle --reps 3 2>&1 | grep -v "iter" | grep "level\|plan:" | cut -c1-175
done
