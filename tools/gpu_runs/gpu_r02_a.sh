#!/bin/bash
# round 2, call A: issue-rate microbench (fp16x2 pipes), stream kernel variant 4 vs 7, sanitizer timing probe, H2D ceiling N=1
mkdir -p gpurun_out
build/issue_rate_bench > gpurun_out/issue_rate_r02.json 2> gpurun_out/issue_rate_r02.err
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variant or golden_vectors_as_one_batch" > gpurun_out/pytest_variants.log 2>&1
tail -3 gpurun_out/pytest_variants.log
for v in 4 7; do
  python bench.py --variant $v --steps 10 --warmup 3 --no-aux --long-pairs 0 > gpurun_out/bench_var$v.json 2> gpurun_out/bench_var$v.err
  python -c "import json;d=json.load(open('gpurun_out/bench_var$v.json'));print($v,d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e']['value'])"
done
python tools/h2d_ceiling.py > gpurun_out/h2d_ceiling_n1.json 2> gpurun_out/h2d_ceiling_n1.err; cat gpurun_out/h2d_ceiling_n1.json
( time timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden_vectors_as_one_batch or ragged" ) > gpurun_out/memcheck_probe.log 2>&1
tail -5 gpurun_out/memcheck_probe.log
