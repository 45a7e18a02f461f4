#!/bin/bash
# external wall time of the CLI vs its own phase timers (300 Mb and 3 Gb)
bash tools/bench_cli.sh 300 A > /dev/null 2>&1
for i in 1 2 3; do /usr/bin/env time -f "external %e s" true 2>/dev/null; s=$(date +%s.%N); PANDEPTH_TIMING=1 pandepth_b200/bin/pandepth -i /tmp/pdcli_300_A/s.bam -o /tmp/x -t 16 2>&1 | grep -E "reduce\+report|decode\+push"; e=$(date +%s.%N); python -c "print('external wall %.3f s' % ($e-$s))"; done
echo "== orderly exit"; s=$(date +%s.%N); PANDEPTH_ORDERLY_EXIT=1 PANDEPTH_TIMING=1 pandepth_b200/bin/pandepth -i /tmp/pdcli_300_A/s.bam -o /tmp/x -t 16 2>&1 | grep -E "pd_destroy"; e=$(date +%s.%N); python -c "print('external wall %.3f s' % ($e-$s))"
REPS=3 THREADS="16" bash tools/bench_cli.sh 3000 A 2>&1 | tee gpurun_out/cli_cfg2.log
for i in 1 2; do s=$(date +%s.%N); PANDEPTH_TIMING=1 pandepth_b200/bin/pandepth -i /tmp/pdcli_3000_A/s.bam -o /tmp/x -t 16 2>&1 | grep -E "reduce\+report|decode\+push"; e=$(date +%s.%N); python -c "print('external wall %.3f s' % ($e-$s))"; done
