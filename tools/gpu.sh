#!/bin/bash
# build the product (abort on any compile error), then run the given command on a B200 via gpurun
set -e
cd "$(dirname "$0")/.."
if ! make -s -j8 -C dsc_b200/csrc > /tmp/dsc_build.log 2>&1; then grep -E "error" -A3 /tmp/dsc_build.log | head -30; echo "BUILD FAILED"; exit 1; fi
exec /usr/local/graft/bin/gpurun "$@"
