#!/bin/bash
# real end-to-end: pandepth (ours) vs the reference binary on the same BAM, decode included, whole-process wall time.
# usage: bench_cli.sh <genome_mb> <flavour A|B> [extra args...]     env: REPS (3), THREADS ("16"), REF_T (52)
MB=${1:-300}; FL=${2:-A}; shift 2
D=/tmp/pdcliP_${MB}_${FL}; mkdir -p $D gpurun_out
L=$((MB*1000000/9)); SPEC=""; for i in 1 2 3 4 5 6 7 8 9; do SPEC="$SPEC,chr$i:$L"; done; SPEC=${SPEC:1}
READS=$((MB*1000000*50/150))
if [ ! -f $D/s.bam.bai ]; then
  t0=$(date +%s.%N); oracle/_ref/bamgen -P -o $D/s.bam -c $SPEC -n $READS -s 7 -t 16 $([ $FL = B ] && echo -q); t1=$(date +%s.%N)
  echo "bamgen: $READS reads, $(du -h $D/s.bam | cut -f1), $(python -c "print('%.1f' % ($t1-$t0))") s"
fi
run() { # name cmd...
  name=$1; shift
  best=999999
  for rep in $(seq 1 ${REPS:-3}); do
    t0=$(date +%s.%N); "$@" > $D/$name.log 2>&1; t1=$(date +%s.%N)
    dt=$(python -c "print($t1-$t0)"); best=$(python -c "print(min($best,$dt))")
  done
  python -c "print('%-34s best of ${REPS:-3}: %.3f s  %.1f Mreads/s' % ('$name', $best, $READS/$best/1e6))"
}
run ref_t${REF_T:-52}   oracle/_ref/pandepth_ref -i $D/s.bam -o $D/ref "$@" -t ${REF_T:-52}
for t in ${THREADS:-16}; do
  run ours_t${t}_device_decode  pandepth_b200/bin/pandepth -i $D/s.bam -o $D/our "$@" -t $t
  for f in $D/ref.*.gz; do b=$(basename $f); o=$D/our${b#ref}; if cmp -s <(zcat $f) <(zcat $o); then echo "  identical (device decode): $b"; else echo "  DIFFERENT (device decode): $b"; fi; done
  run ours_t${t}_host_decode    env PANDEPTH_GPU_DECODE=0 pandepth_b200/bin/pandepth -i $D/s.bam -o $D/ourh "$@" -t $t
  for f in $D/ref.*.gz; do b=$(basename $f); o=$D/ourh${b#ref}; if cmp -s <(zcat $f) <(zcat $o); then echo "  identical (host decode): $b"; else echo "  DIFFERENT (host decode): $b"; fi; done
done
echo "-- phase times (PANDEPTH_TIMING=1), device decode then host decode"
PANDEPTH_TIMING=1 pandepth_b200/bin/pandepth -i $D/s.bam -o $D/our "$@" -t ${THREADS:-16} 2>&1 | grep timing
PANDEPTH_TIMING=1 PANDEPTH_GPU_DECODE=0 pandepth_b200/bin/pandepth -i $D/s.bam -o $D/ourh "$@" -t ${THREADS:-16} 2>&1 | grep timing
