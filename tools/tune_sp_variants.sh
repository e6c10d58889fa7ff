#!/bin/bash
# development aid: time the single-phase step (default settings) for each tuning build under build/variants
for so in build/variants/*.so; do
  echo "== $so"
  HYPERELASTIC_B200_LIB=$PWD/$so python tools/sp_quick.py "$@"
done
