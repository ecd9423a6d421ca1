#!/bin/bash
mkdir -p gpurun_out
python tools/probe_ranges.py > gpurun_out/probe_ranges.log 2>&1; cat gpurun_out/probe_ranges.log | tail -120
