#!/bin/bash
# round 2, call B: dynamic couple distribution (variants 8-11) parity + timing
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variant or golden_vectors_as_one_batch" > gpurun_out/pytest_variants.log 2>&1
tail -3 gpurun_out/pytest_variants.log
for v in 4 8 9 10 11; do
  python bench.py --variant $v --steps 10 --warmup 3 --no-aux --long-pairs 0 > gpurun_out/bench_var$v.json 2> gpurun_out/bench_var$v.err
  python -c "import json;d=json.load(open('gpurun_out/bench_var$v.json'));print($v,d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e']['value'],d['config']['host_path_equals_device_path'])"
done
