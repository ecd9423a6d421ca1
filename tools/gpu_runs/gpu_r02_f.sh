#!/bin/bash
mkdir -p gpurun_out
python tools/probe_sustained.py > gpurun_out/probe_sustained_r02.json 2> gpurun_out/probe_sustained.err; tail -3 gpurun_out/probe_sustained.err; head -c 6000 gpurun_out/probe_sustained_r02.json
for v in 9 12 13 14; do
  SWB_LIB_OVERRIDE=tests/native/libswb200_variants.so python bench.py --variant $v --steps 10 --warmup 3 --no-aux --long-pairs 0 --long-wave 0 --strong-pairs 0 --bgzf-reads 0 --cpu-passes 1 > gpurun_out/bench_var$v.json 2> gpurun_out/bench_var$v.err
  python -c "import json;d=json.load(open('gpurun_out/bench_var$v.json'));print($v,d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e_resident_reference']['ms_per_step'])"
done
python -m pytest tests/test_gpu_ranges_multi.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
