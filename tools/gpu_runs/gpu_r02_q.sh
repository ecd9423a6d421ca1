#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_r02_final2.log 2>&1; tail -4 gpurun_out/pytest_gpu_r02_final2.log
python tools/bench_wgs.py --reads-per-file 2000000 --devices 1 --dir /tmp/synwgs_gz > gpurun_out/wgs_gzip_own_r02.json 2> gpurun_out/wgs_gzip_own.err; cut -c1-900 gpurun_out/wgs_gzip_own_r02.json
SWB_HOST_INFLATE=zlib python tools/bench_wgs.py --reads-per-file 2000000 --devices 1 --dir /tmp/synwgs_gz --reuse > gpurun_out/wgs_gzip_zlib_r02.json 2> gpurun_out/wgs_gzip_zlib.err; cut -c1-900 gpurun_out/wgs_gzip_zlib_r02.json
python tools/bench_wgs.py --reads-per-file 2000000 --devices 1 --dir /tmp/synwgs_gz --reuse > gpurun_out/wgs_gzip_own_r02_b.json 2>> gpurun_out/wgs_gzip_own.err; cut -c1-500 gpurun_out/wgs_gzip_own_r02_b.json
python tools/bench_wgs.py --reads-per-file 2000000 --devices 1 --dir /tmp/synwgs_gzn --quals noisy > gpurun_out/wgs_gzip_noisy_own_r02.json 2> gpurun_out/wgs_gzip_noisy.err; cut -c1-500 gpurun_out/wgs_gzip_noisy_own_r02.json
SWB_HOST_INFLATE=zlib python tools/bench_wgs.py --reads-per-file 2000000 --devices 1 --dir /tmp/synwgs_gzn --quals noisy --reuse > gpurun_out/wgs_gzip_noisy_zlib_r02.json 2>> gpurun_out/wgs_gzip_noisy.err; cut -c1-500 gpurun_out/wgs_gzip_noisy_zlib_r02.json
nproc
