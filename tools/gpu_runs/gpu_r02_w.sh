#!/bin/bash
python -m pytest tests/test_gpu_fastq.py tests/test_gpu_host.py -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()"
