#!/bin/bash
# ThreadSanitizer over the two threaded pieces of the host side (no GPU): the BGZF readers' ordered hand-off
# (rsm_debug_bgzf_segments: the driver's reader threads against an in-order consumer), hgz::AsyncGunzip, and the parallel reader
# hgz::ParallelGunzip (tests/tools/pgunzip_fuzz.cpp: mutated copies of the gzip file, random chunk sizes and thread counts).
# usage: tests/tools/tsan/run.sh <file.bgzf.gz> <file.gz>     (any BGZF file of a few MB, any gzip file)
set -e
here="$(cd "$(dirname "$0")" && pwd)"; root="$here/../../.."; out=${TMPDIR:-/tmp}/swb_tsan; mkdir -p "$out"
g++ -O1 -g -fsanitize=thread -std=c++17 -pthread -I/usr/local/cuda/include -c "$root/mini_parallel_b200/csrc/rustseq_host.cpp" -o "$out/host.o"
g++ -fsanitize=thread -pthread -std=c++17 -o "$out/bgzf_readers" "$here/bgzf_readers_main.cpp" "$here/swb_stubs.cpp" "$out/host.o" -lz
g++ -O1 -g -fsanitize=thread -std=c++17 -pthread -o "$out/async_gunzip" "$here/async_gunzip_main.cpp" -lz
"$out/bgzf_readers" "$1" | tail -4
"$out/async_gunzip" "$2" | tail -4
g++ -O1 -g -fsanitize=thread -std=c++17 -pthread -o "$out/pgunzip_fuzz" "$here/../pgunzip_fuzz.cpp" -lz
"$out/pgunzip_fuzz" "$2" "$out/scratch.gz" ${3:-60} 1 | tail -4
echo "tsan: no report above = no data race seen"
