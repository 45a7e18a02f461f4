#!/bin/bash
# cfg2 / cfg3 / cfg4 end to end through the real binaries (3 Gb x 50x, 1 G reads, flavour A)
mkdir -p gpurun_out
export REPS=2 THREADS="${THREADS:-32 64 128}"
bash tools/bench_cli.sh 3000 A 2>&1 | tee gpurun_out/cli_cfg2.log
L=$((3000*1000000/9)); SPEC=""; for i in 1 2 3 4 5 6 7 8 9; do SPEC="$SPEC,chr$i:$L"; done; SPEC=${SPEC:1}
python tools/make_gff.py /tmp/pdcli_3000_A/genes.gff $SPEC
THREADS="64 128" bash tools/bench_cli.sh 3000 A -g /tmp/pdcli_3000_A/genes.gff 2>&1 | tee gpurun_out/cli_cfg4.log
REPS=1 THREADS="64" bash tools/bench_cli.sh 3000 A -w 100 2>&1 | tee gpurun_out/cli_cfg3.log
for t in 32 64 128; do PANDEPTH_TIMING=1 pandepth_b200/bin/pandepth -i /tmp/pdcli_3000_A/s.bam -o /tmp/pdcli_3000_A/x -t $t 2>&1 | grep timing | sed "s/^/t=$t /"; done | tee gpurun_out/cli_cfg2_timing.log
PANDEPTH_TIMING=1 pandepth_b200/bin/pandepth -i /tmp/pdcli_3000_A/s.bam -o /tmp/pdcli_3000_A/x -t 64 -w 100 2>&1 | grep timing | tee -a gpurun_out/cli_cfg2_timing.log
python tools/bench_windows.py 3e9 100 2>&1 | tail -1 | tee gpurun_out/windows_line.json; python tools/bench_windows.py 3e9 1000 2>&1 | tail -1 | tee -a gpurun_out/windows_line.json
