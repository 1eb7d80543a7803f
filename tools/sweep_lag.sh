#!/bin/bash
# lag / ring sweep of the TMA-fed two-pass launch at one length (scratch; knobs DSC_TMA_LAG, DSC_TMA_RING_MB)
lg=${1:-16}
export DSC_NO_CLUSTER=1
for ring in 32 64 128; do
  for lag in 2 4 8 16 24 32 48 64; do
    r=$(DSC_TMA_RING_MB=$ring DSC_TMA_LAG=$lag timeout 60 python tools/check_tma.py 0 $lg 2>&1 | grep "rows=[0-9][0-9][0-9]" | sed -E 's/.*fwd\+inv +([0-9]+) GB.*/\1/')
    echo "lg=$lg ring=${ring}MB lag=$lag: $r GB/s"
  done
done
