#!/bin/bash
# round 2, call E: smoke + the new bench line (N=1) + launch list + ncu --set full of the stream kernel and the packing kernels
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke_r02.log 2>&1; tail -2 gpurun_out/smoke_r02.log
( time python bench.py ) > gpurun_out/bench_r02_v1.json 2> gpurun_out/bench_r02_v1.err; tail -3 gpurun_out/bench_r02_v1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_v1.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'kernel_ms',d['roofline']['kernel_ms'])
for k in ('e2e','e2e_csr_windows','e2e_resident_reference'): print(k,d[k]['value'],d[k]['ms_per_step'],d[k]['h2d_bytes_per_step'],d[k]['equals_device_resident_results'])
print('strong',d['config2_strong'])
print('long',d['aux_long_pairs'])
print('bgzf',d['aux_bgzf_ingest'])
print('cpu',d['cpu_baseline'])
print('aux',{k:v['gcups'] for k,v in d['aux'].items()})
PY
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_r02_v1.json 2> gpurun_out/bench_ref_r02_v1.err; cat gpurun_out/bench_ref_r02_v1.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_v1.csv python bench.py --steps 2 --warmup 3 --no-aux --long-pairs 0 --long-wave 0 --strong-pairs 0 --bgzf-reads 0 --cpu-passes 1 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sw_stream_kernel -s 3 -c 1 -o gpurun_out/stream_r02_v1 -f python bench.py --steps 2 --warmup 3 --no-aux --long-pairs 0 --long-wave 0 --strong-pairs 0 --bgzf-reads 0 --cpu-passes 1 > gpurun_out/ncu_stream.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pack2bit_kernel -s 7 -c 1 -o gpurun_out/pack_r02_v1 -f python bench.py --steps 2 --warmup 3 --no-aux --long-pairs 0 --long-wave 0 --strong-pairs 0 --bgzf-reads 0 --cpu-passes 1 > gpurun_out/ncu_pack.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:classify_kernel -s 3 -c 1 -o gpurun_out/classify_r02_v1 -f python bench.py --steps 2 --warmup 3 --no-aux --long-pairs 0 --long-wave 0 --strong-pairs 0 --bgzf-reads 0 --cpu-passes 1 > gpurun_out/ncu_classify.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
