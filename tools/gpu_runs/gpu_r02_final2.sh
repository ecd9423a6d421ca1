#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_r02_final3.log 2>&1; head -3 gpurun_out/pytest_gpu_r02_final3.log
python -c "import __graft_entry__ as g; g.smoke()"
( time python bench.py ) > gpurun_out/bench_r02_v5.json 2> gpurun_out/bench_r02_v5.err; tail -3 gpurun_out/bench_r02_v5.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r02_v5.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'csr',d['e2e_csr_windows']['value'],'resident',d['e2e_resident_reference']['value'])
print('strong',d['config2_strong']['gcups'],d['config2_strong']['score_ms_max_over_ranks'],'long',d['aux_long_pairs']['full']['gcups'],'bgzf',d['aux_bgzf_ingest']['reads_per_s'],'cpu',d['cpu_baseline']['value'],'launches',d['gpu_launches'])"
