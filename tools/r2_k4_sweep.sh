#!/bin/bash
for mult in 1 2; do for tw in 1 2 3 4 6 99; do
  echo "wave_mult $mult tail_waves $tw"; VEX_K4_WAVE_MULT=$mult VEX_K4_TAIL_WAVES=$tw timeout 120 python tools/k4_ab.py one
done; done
