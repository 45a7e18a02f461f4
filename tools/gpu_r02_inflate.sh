#!/bin/bash
# next round, first decode-side trip: GPU BGZF inflate probe on flavour A (SEQ-less) and flavour B (SEQ+QUAL) BAMs
mkdir -p gpurun_out /tmp/pdinf
G=oracle/_ref/bamgen
spec=$(python -c "print(','.join(f'chr{i+1}:{30000000//9}' for i in range(9)))")
$G -o /tmp/pdinf/A.bam -c $spec -n 10000000 -s 7 -t 16
$G -o /tmp/pdinf/B.bam -c $spec -n 2000000 -s 7 -q -t 16
for f in A B; do ( timeout 120 build/gpu_inflate_probe /tmp/pdinf/$f.bam ) 2>&1 | tee -a gpurun_out/inflate_probe.log; done
for f in A B; do ( timeout 120 build/gpu_decode_probe /tmp/pdinf/$f.bam; PANDEPTH_DRY=1 pandepth_b200/bin/pandepth -i /tmp/pdinf/$f.bam -o /tmp/pdinf/x -s | grep dry ) 2>&1 | tee -a gpurun_out/decode_probe.log; done
