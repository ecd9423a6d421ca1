#!/bin/bash
# 8-GPU box: where does --full-wgs lose its scaling?  io ceiling, N=1 vs N=8, readers per file x segment size
mkdir -p gpurun_out
R=${WGS_READS:-8000000}
run() {  # tag devices readers seg_mb
  SWB_STAMPS=1 SWB_BGZF_SEGMENT_MB=$4 python tools/bench_wgs.py --bgzf --reads-per-file $R --devices $2 --dir /tmp/synwgs --readers $3 --reuse > gpurun_out/wgs8_$1.json 2> gpurun_out/wgs8_$1.err
  python - <<PY
import json; d=json.load(open("gpurun_out/wgs8_$1.json"))
print("$1 devices=$2 readers=$3 seg=$4MB: wall", d["wall_s"], "files", d["slowest_file_s"], "pipeline Mreads/s", round(d["pipeline_reads_per_s"]/1e6,1), "e2e Mreads/s", round(d["reads_per_s"]/1e6,1))
PY
  grep -E "context created|pinned|all files|gpu available|full wgs returned|subprocess" gpurun_out/wgs8_$1.err | tr '\n' ';' | cut -c1-900; echo
}
python tools/bench_wgs.py --bgzf --reads-per-file $R --devices 1 --dir /tmp/synwgs --io-ceiling > gpurun_out/wgs8_io_ceiling.json 2> gpurun_out/wgs8_gen.err; cat gpurun_out/wgs8_io_ceiling.json
run n1_r1_s112 1 1 112
run n8_r1_s112 8 1 112
run n8_r2_s112 8 2 112
run n8_r1_s48 8 1 48
run n8_r2_s48 8 2 48
run n8_r2_s24 8 2 24
run n8_r4_s24 8 4 24
run n1_r1_s48 1 1 48
run n8_r1_s112_again 8 1 112
nproc; free -g | head -2
