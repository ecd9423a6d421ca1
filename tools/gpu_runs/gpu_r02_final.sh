#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_r02_final.log 2>&1; tail -4 gpurun_out/pytest_gpu_r02_final.log
python -c "import __graft_entry__ as g; g.smoke()"
( time python bench.py --impl reference ) > gpurun_out/bench_reference_arm_r02_v4.json 2> gpurun_out/bench_reference_arm_r02_v4.err; cut -c1-200 gpurun_out/bench_reference_arm_r02_v4.json; tail -3 gpurun_out/bench_reference_arm_r02_v4.err
( time python bench.py ) > gpurun_out/bench_r02_v4.json 2> gpurun_out/bench_r02_v4.err; tail -3 gpurun_out/bench_r02_v4.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r02_v4.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'csr',d['e2e_csr_windows']['value'],'resident',d['e2e_resident_reference']['value'])
print('strong',d['config2_strong']['gcups'],d['config2_strong']['score_ms_max_over_ranks'],'long',d['aux_long_pairs']['full']['gcups'],'bgzf',d['aux_bgzf_ingest']['reads_per_s'],'cpu',d['cpu_baseline']['value'],'clocks',d['clocks'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_v4.csv python bench.py --steps 2 --warmup 3 --no-aux --long-pairs 0 --long-wave 0 --strong-pairs 0 --bgzf-reads 0 --cpu-passes 1 > gpurun_out/ncu_launch_v4.log 2>&1; tail -2 gpurun_out/ncu_launch_v4.log | cut -c1-200
