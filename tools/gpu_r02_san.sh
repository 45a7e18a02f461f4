#!/bin/bash
# compute-sanitizer over the kernels changed after the final trip: bucketed unsorted path (tools/san_bucket.py) and the inflate kernels
# with the CRC32 check, the speculative record chain and the entry reuse (BGZF tests)
mkdir -p gpurun_out
if [ "$1" = bucket ]; then
( timeout 110 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/san_bucket.py ) > gpurun_out/san_bucket_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 3 gpurun_out/san_bucket_racecheck.log
( timeout 110 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/san_bucket.py ) > gpurun_out/san_bucket_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 3 gpurun_out/san_bucket_memcheck.log
else
( timeout 100 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_bgzf_gpu.py -x -q -k "crc32 or (literal_dominated and 2) or (fixtures and k.sam)" ) > gpurun_out/san_bgzf_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 4 gpurun_out/san_bgzf_racecheck.log
( timeout 60 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_bgzf_gpu.py -x -q -k "crc32 or (fixtures and k.sam)" ) > gpurun_out/san_bgzf_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 4 gpurun_out/san_bgzf_memcheck.log
fi
