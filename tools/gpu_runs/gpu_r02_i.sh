#!/bin/bash
mkdir -p gpurun_out
for mb in 32 48 64 96 128; do
  SWB_CHUNK_MB=$mb python bench.py --steps 5 --warmup 3 --no-aux --long-pairs 0 --long-wave 0 --strong-pairs 0 --bgzf-reads 0 --cpu-passes 1 > gpurun_out/bench_chunk$mb.json 2> gpurun_out/bench_chunk$mb.err
  python -c "import json;d=json.load(open('gpurun_out/bench_chunk$mb.json'));print($mb,'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'csr',d['e2e_csr_windows']['ms_per_step'],'resident',d['e2e_resident_reference']['ms_per_step'])"
done
python -m pytest tests/test_gpu_ranges_multi.py tests/test_gpu_parity.py -x -q -m gpu -k "ranges or chunk or host_path or reference" 2>&1 | tail -2
