#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mid_path or mixed" 2>&1 | tail -2
SWB_DEBUG=1 python tools/bench_wgs.py --bgzf --reads-per-file 2000000 --devices 1 --dir /tmp/synwgs > gpurun_out/wgs_probe.json 2> gpurun_out/wgs_probe.err; cat gpurun_out/wgs_probe.json | cut -c1-500; grep -E "^\[main\]|^\[wgs\]" gpurun_out/wgs_probe.err
cd /tmp/synwgs && GPU_CHUNK_SIZE_READS=100000 WGS_DATA_DIR=/tmp/synwgs WGS_SAMPLE_ID=SYN WGS_CHECKPOINT_DIR=/tmp/synwgs strace -f -c -o /tmp/strace.txt /root/repo/build/rustseq_mini --full-wgs --gpu > /dev/null 2>&1; head -25 /tmp/strace.txt
( time /root/repo/build/rustseq_mini -1 ACGT -2 ACGT --gpu ) 2>&1 | tail -4
