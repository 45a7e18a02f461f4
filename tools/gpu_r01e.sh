#!/bin/bash
# last seconds of the round's GPU budget: the CLI (new host batch layout, -O3) against the golden outputs and on CRAM
mkdir -p gpurun_out
( timeout 48 python -m pytest tests/test_cli_gpu.py -m gpu -q -n 8 -x -k "(golden and (mix or big or gff or gtf or list3k or paf)) or cram" ) > gpurun_out/e_cli.log 2>&1; echo "rc=$?" >> gpurun_out/e_cli.log
tail -n 6 gpurun_out/e_cli.log
