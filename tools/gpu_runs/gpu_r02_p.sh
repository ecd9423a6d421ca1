#!/bin/bash
mkdir -p gpurun_out
python tests/tools/fuzz_parity.py --seconds 110 --seed 202 > gpurun_out/fuzz_r02_a.log 2>&1; tail -2 gpurun_out/fuzz_r02_a.log
python tests/tools/fuzz_parity.py --seconds 110 --seed 303 > gpurun_out/fuzz_r02_b.log 2>&1; tail -2 gpurun_out/fuzz_r02_b.log
