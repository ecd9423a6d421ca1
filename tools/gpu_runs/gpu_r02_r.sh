#!/bin/bash
mkdir -p gpurun_out

timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:sw_stream_kernel<.int.32" -s 2 -c 1 -o gpurun_out/mid_r02_v1 -f python tools/probe_mid.py > gpurun_out/ncu_mid.log 2>&1; tail -3 gpurun_out/ncu_mid.log | cut -c1-200
ls -la gpurun_out/mid_r02_v1.ncu-rep
