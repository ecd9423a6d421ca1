#!/bin/bash
# build, then run a command on the GPU box (never ship a stale libswb200.so)
set -e
cd "$(dirname "$0")/.."
make -s all 2>&1 | grep -E "error|warning" && exit 1
T=${GRUN_TIMEOUT:-900}
exec gpurun ${GRUN_GPUS:+--gpus $GRUN_GPUS} --timeout $T -- "$@"
