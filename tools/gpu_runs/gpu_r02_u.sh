#!/bin/bash
mkdir -p gpurun_out
lscpu | grep -E "Thread|Core|Socket|Model name|^CPU\(s\)" | head -6
python -m pytest tests/test_gpu_host.py -x -q -m gpu 2>&1 | tail -2
for mode in 0 1 0 1; do
  SWB_ASYNC_INFLATE=$mode python tools/bench_wgs.py --lanes 4 --reads-per-file 2000000 --devices 1 --dir /tmp/synwgs_gz8 --reuse > gpurun_out/wgs_async_$mode.json 2> gpurun_out/wgs_async.err || SWB_ASYNC_INFLATE=$mode python tools/bench_wgs.py --lanes 4 --reads-per-file 2000000 --devices 1 --dir /tmp/synwgs_gz8 > gpurun_out/wgs_async_$mode.json 2>> gpurun_out/wgs_async.err
  python -c "
import json; d=json.load(open('gpurun_out/wgs_async_$mode.json')); print('async=$mode: 8 files, slowest file', d['slowest_file_s'], 'pipeline Mreads/s', round(d['pipeline_reads_per_s']/1e6,1))"
done
