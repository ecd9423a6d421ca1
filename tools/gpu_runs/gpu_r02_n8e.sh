#!/bin/bash
mkdir -p gpurun_out
R=${WGS_READS:-37500000}
run() {  # tag devices env...
  tag=$1; dev=$2; shift 2
  env SWB_STAMPS=1 "$@" python tools/bench_wgs.py --bgzf --reads-per-file $R --devices $dev --dir /tmp/synwgs --reuse --clone-files > gpurun_out/wgs_e_$tag.json 2> gpurun_out/wgs_e_$tag.err
  python - <<PY
import json; d=json.load(open("gpurun_out/wgs_e_$tag.json"))
print("$tag devices=$dev $*: wall", d["wall_s"], "slowest file", d["slowest_file_s"], "pipeline Mreads/s", round(d["pipeline_reads_per_s"]/1e6,1), "TCUPS", round(d["pipeline_gcups"]/1e3,2), "startup", d["startup_s"])
PY
  grep -E "pinned|device consumer" gpurun_out/wgs_e_$tag.err | cut -c1-220
}
python tools/bench_wgs.py --bgzf --reads-per-file $R --devices 1 --dir /tmp/synwgs --clone-files --io-ceiling > /dev/null 2>&1
run n8 8 SWB_X=0
run n8_r2 8 SWB_READERS_PER_FILE=2
run n8_again 8 SWB_X=0
