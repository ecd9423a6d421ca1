#!/bin/bash
# 8-GPU box: --full-wgs at BASELINE.json configs[4] size (16 files x 37.5 M reads = 600 M reads) on 1 and on 8 GPUs, final build
mkdir -p gpurun_out
df -h /tmp | tail -1
FREE=$(df --output=avail -BG /tmp | tail -1 | tr -dc 0-9)
R=37500000; [ "$FREE" -lt 70 ] && R=16000000
echo "reads per file: $R (free: ${FREE}G)"
run() {  # tag devices extra-args
  SWB_STAMPS=1 python tools/bench_wgs.py --bgzf --reads-per-file $R --devices $2 --dir /tmp/synwgs --reuse --clone-files > gpurun_out/wgs_full_$1.json 2> gpurun_out/wgs_full_$1.err
  python - <<PY
import json; d=json.load(open("gpurun_out/wgs_full_$1.json"))
print("$1 devices=$2: wall", d["wall_s"], "slowest file", d["slowest_file_s"], "pipeline Mreads/s", round(d["pipeline_reads_per_s"]/1e6,1), "TCUPS", d["pipeline_gcups"]/1e3, "e2e Mreads/s", round(d["reads_per_s"]/1e6,1), "startup", d["startup_s"], "files", d["files_done"])
PY
  grep -E "context created|pinned|all files|gpu available|full wgs returned|subprocess" gpurun_out/wgs_full_$1.err | tr '\n' ';' | cut -c1-700; echo
}
( time python tools/bench_wgs.py --bgzf --reads-per-file $R --devices 1 --dir /tmp/synwgs --clone-files --io-ceiling ) > gpurun_out/wgs_full_io_ceiling.json 2> gpurun_out/wgs_full_gen.err; cat gpurun_out/wgs_full_io_ceiling.json; tail -3 gpurun_out/wgs_full_gen.err
run n8 8
run n1 1
run n8_again 8
run n4 4
run n2 2
