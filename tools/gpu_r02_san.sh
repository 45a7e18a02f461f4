#!/bin/bash
mkdir -p gpurun_out
( timeout 100 python tools/san_bucket.py ) > gpurun_out/san_bucket_plain.log 2>&1; tail -n 1 gpurun_out/san_bucket_plain.log
( timeout 110 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/san_bucket.py ) > gpurun_out/san_bucket_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 3 gpurun_out/san_bucket_racecheck.log
( timeout 110 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/san_bucket.py ) > gpurun_out/san_bucket_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 3 gpurun_out/san_bucket_memcheck.log
