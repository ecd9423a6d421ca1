#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_host.py tests/test_gpu_fastq.py -x -q -m gpu ) 2>&1 | tail -15
SWB_STAMPS=1 python tools/bench_wgs.py --bgzf --reads-per-file 4000000 --devices 1 --dir /tmp/synwgs --readers 1 > gpurun_out/wgs_r1.json 2> gpurun_out/wgs_r1.err; cut -c1-700 gpurun_out/wgs_r1.json
SWB_STAMPS=1 python tools/bench_wgs.py --bgzf --reads-per-file 4000000 --devices 1 --dir /tmp/synwgs --readers 2 --reuse > gpurun_out/wgs_r2.json 2> gpurun_out/wgs_r2.err; cut -c1-700 gpurun_out/wgs_r2.json
python tools/bench_wgs.py --bgzf --reads-per-file 4000000 --dir /tmp/synwgs --reuse --io-ceiling
