#!/bin/bash
mkdir -p gpurun_out
R=${WGS_READS:-16000000}
run() {  # tag devices env...
  tag=$1; dev=$2; shift 2
  env SWB_STAMPS=1 "$@" python tools/bench_wgs.py --bgzf --reads-per-file $R --devices $dev --dir /tmp/synwgs --reuse --clone-files > gpurun_out/wgs_d_$tag.json 2> gpurun_out/wgs_d_$tag.err
  python - <<PY
import json; d=json.load(open("gpurun_out/wgs_d_$tag.json"))
print("$tag devices=$dev $*: wall", d["wall_s"], "slowest file", d["slowest_file_s"], "pipeline Mreads/s", round(d["pipeline_reads_per_s"]/1e6,1), "startup", d["startup_s"])
PY
  grep -E "pinned|device consumer" gpurun_out/wgs_d_$tag.err | cut -c1-220
}
python tools/bench_wgs.py --bgzf --reads-per-file $R --devices 1 --dir /tmp/synwgs --clone-files --io-ceiling > /dev/null 2>&1
run base 8 SWB_X=0
run block 8 SWB_BLOCKING_SYNC=1
run block_r1 8 SWB_BLOCKING_SYNC=1 SWB_READERS_PER_FILE=1
run block_s64 8 SWB_BLOCKING_SYNC=1 SWB_BGZF_SEGMENT_MB=64
run block_s16 8 SWB_BLOCKING_SYNC=1 SWB_BGZF_SEGMENT_MB=16
run n1 1 SWB_X=0
