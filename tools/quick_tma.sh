#!/bin/bash
# quick look at the TMA-fed two-pass launch: parity + GB/s at the given lengths, then the pass-through pipeline alone
# usage: tools/quick_tma.sh tag [lg ...]       (environment knobs are passed through)
tag=$1; shift
lgs=${@:-15 16 17 18 19 20}
export DSC_NO_CLUSTER=${DSC_NO_CLUSTER-1}
(timeout 300 python tools/check_tma.py 0 $lgs 2>&1; echo "rc=$?") | grep "rows=[0-9][0-9][0-9]\|rc=\|FAILED\|stuck" | sed "s/^/$tag /"
# the pass-through experiment needs a build with the hook: make -C dsc_b200/csrc DSC_TMA_EXPERIMENTS=1 BUILD=build_exp TARGET=../libdsc_exp.so
[ -f dsc_b200/libdsc_exp.so ] || exit 0
for lg in 16 18 20; do DSC_LIB=dsc_b200/libdsc_exp.so DSC_TMA_DEBUG_SKIP=1 timeout 60 python tools/check_tma.py 0 $lg 2>&1 | grep "rows=[0-9][0-9][0-9]" | sed "s/^/$tag skip=1 /"; done
