#!/bin/bash
# round 2, call C: variant 12 (adds on the FMA pipe) timing, parity of every stream variant, full-population parity tests
mkdir -p gpurun_out
for v in 9 12; do
  python bench.py --variant $v --steps 10 --warmup 3 --no-aux --long-pairs 0 > gpurun_out/bench_var$v.json 2> gpurun_out/bench_var$v.err
  python -c "import json;d=json.load(open('gpurun_out/bench_var$v.json'));print($v,d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e']['value'],d['config']['host_path_equals_device_path'])"
done
( time python -m pytest tests/test_gpu_parity.py -x -q -m gpu ) > gpurun_out/pytest_parity.log 2>&1
tail -5 gpurun_out/pytest_parity.log
