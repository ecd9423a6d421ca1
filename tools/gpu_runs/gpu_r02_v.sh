#!/bin/bash
mkdir -p gpurun_out
SAN_TOOLS=memcheck SAN_TIMEOUT=30 tools/sanitize.sh
