#!/bin/bash
mkdir -p gpurun_out
SAN_TOOLS=memcheck SAN_TIMEOUT=60 tools/sanitize.sh
