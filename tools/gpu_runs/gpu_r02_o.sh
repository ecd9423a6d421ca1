#!/bin/bash
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/bench_r02_v3.json 2> gpurun_out/bench_r02_v3.err; tail -3 gpurun_out/bench_r02_v3.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r02_v3.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'csr',d['e2e_csr_windows']['value'],'resident',d['e2e_resident_reference']['value'])
print('strong',d['config2_strong']['gcups'],d['config2_strong']['score_ms_max_over_ranks'],'long',d['aux_long_pairs']['full']['gcups'],'bgzf',d['aux_bgzf_ingest']['reads_per_s'],'cpu',d['cpu_baseline']['value'])"
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_reference_arm_r02_v3.json 2> gpurun_out/bench_reference_arm_r02_v3.err; cut -c1-300 gpurun_out/bench_reference_arm_r02_v3.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
