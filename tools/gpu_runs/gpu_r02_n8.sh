#!/bin/bash
# round 2, the 8-GPU call: H2D ceiling at N = 1, 2, 4, 8; bench.py at N = 8 (and 4, 2); one-process multi-device batch; 2-device test;
# --full-wgs on BGZF at size on 1 and 8 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | head -8
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
python tools/h2d_ceiling.py > gpurun_out/h2d_ceiling_n1.json 2> gpurun_out/h2d_ceiling.err
for n in 2 4 8; do $TR --nproc-per-node $n --master-port 2951$n tools/h2d_ceiling.py > gpurun_out/h2d_ceiling_n$n.json 2>> gpurun_out/h2d_ceiling.err; done
cat gpurun_out/h2d_ceiling_n*.json
python -m pytest tests/test_gpu_ranges_multi.py -x -q -m gpu -k multi 2>&1 | tail -2
for n in 1 8; do python tools/bench_multi.py --gpus $n > gpurun_out/bench_multi_n$n.json 2> gpurun_out/bench_multi_n$n.err; cat gpurun_out/bench_multi_n$n.json | cut -c1-600; done
for n in 8; do
  ( time $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n ) > gpurun_out/bench_r02_v2_n$n.json 2> gpurun_out/bench_r02_v2_n$n.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_r02_v2_n$n.json'))
print($n,'value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'csr',d['e2e_csr_windows']['value'],'resident',d['e2e_resident_reference']['value'],'strong',d['config2_strong']['gcups'],d['config2_strong']['score_ms_max_over_ranks'],d['config2_strong']['checksum64'],d['config2_strong']['oracle_equal'])"
done
# WGS at size: 16 files x 16 M reads
python tools/bench_wgs.py --bgzf --reads-per-file ${WGS_READS:-16000000} --devices 1 --dir /tmp/synwgs > gpurun_out/wgs_e2e_bgzf_n1_r02.json 2> gpurun_out/wgs_n1.err; cat gpurun_out/wgs_e2e_bgzf_n1_r02.json | cut -c1-900
SWB_DEBUG=1 python tools/bench_wgs.py --bgzf --reads-per-file ${WGS_READS:-16000000} --devices 8 --dir /tmp/synwgs --reuse > gpurun_out/wgs_e2e_bgzf_n8_r02.json 2> gpurun_out/wgs_n8.err; cat gpurun_out/wgs_e2e_bgzf_n8_r02.json | cut -c1-900
grep "\[wgs\]" gpurun_out/wgs_n8.err | head -20
python tools/bench_wgs.py --bgzf --reads-per-file ${WGS_READS:-16000000} --devices 8 --dir /tmp/synwgs --reuse > gpurun_out/wgs_e2e_bgzf_n8_r02_b.json 2> gpurun_out/wgs_n8_b.err; cat gpurun_out/wgs_e2e_bgzf_n8_r02_b.json | cut -c1-400
df -h /tmp | tail -1
