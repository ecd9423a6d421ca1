#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --no-aux --long-pairs 0 --long-wave 0 --bgzf-reads 0 --cpu-passes 1 --strong-pairs 20000000 > gpurun_out/bench_strong_probe.json 2> gpurun_out/bench_strong_probe.err
python -c "import json;d=json.load(open('gpurun_out/bench_strong_probe.json'));print(d['config2_strong'])"
