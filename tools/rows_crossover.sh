#!/bin/bash
# rows kernel vs CTA kernel as the pupil grows (about 1.5 M rays each)
for side in 8 12 16 22 32 45; do
  p=$((side*side)); b=$((65536/p)); [ $b -lt 3 ] && b=3
  echo "-- pupil $p, $b lenses"
  TL_ROWS_MAX_PUPIL=100000 python tools/profile_batched.py $b $side 2>&1 | tail -1 | sed 's/^/rows: /'
  TL_NO_ROWS=1 python tools/profile_batched.py $b $side 2>&1 | tail -1 | sed 's/^/cta : /'
done
