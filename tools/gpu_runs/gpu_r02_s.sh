#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_ranges_multi.py -x -q -m gpu 2>&1 | tail -4
python tools/probe_mid.py > gpurun_out/mid_path_r02_v2.json 2> gpurun_out/mid_path_v2.err; python -c "
import json; d=json.load(open('gpurun_out/mid_path_r02_v2.json'))
for k,v in d.items(): print(k, {kk: (vv['ms'], vv['gcups']) for kk,vv in v.items() if isinstance(vv, dict)}, v.get('identical_results'))"
python tests/tools/fuzz_parity.py --seconds 40 --seed 404 2>&1 | tail -1
