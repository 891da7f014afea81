# round-1 sweep kernels on the 256^3 hierarchy, per level, for a few values of B200AMG_GS_POLL_MASKED
for pm in 1 2; do
  echo "=== GS_BLOCK=0 POLL_MASKED=$pm"
  B200AMG_GS_BLOCK=0 B200AMG_GS_POLL_MASKED=$pm timeout 400 python tools/block_smoke.py --sizes 256 --no-oracle --reps 3 2>&1 | grep "level"
done
