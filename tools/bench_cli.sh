#!/bin/bash
# real end-to-end: pandepth (ours) vs the reference binary on the same BAM, decode included.
# usage: bench_cli.sh <genome_mb> <flavour A|B> [extra args...]
MB=${1:-300}; FL=${2:-A}; shift 2
D=/tmp/pdcli_${MB}_${FL}; mkdir -p $D gpurun_out
L=$((MB*1000000/9)); SPEC=""; for i in 1 2 3 4 5 6 7 8 9; do SPEC="$SPEC,chr$i:$L"; done; SPEC=${SPEC:1}
READS=$((MB*1000000*50/150))
if [ ! -f $D/s.bam.bai ]; then
  t0=$(date +%s.%N); oracle/_ref/bamgen -o $D/s.bam -c $SPEC -n $READS -s 7 -t 16 $([ $FL = B ] && echo -q); t1=$(date +%s.%N)
  echo "bamgen: $READS reads, $(du -h $D/s.bam | cut -f1), $(echo "$t1-$t0" | bc -l 2>/dev/null || python -c "print($t1-$t0)") s"
fi
run() { # name cmd...
  name=$1; shift
  best=999999
  for rep in $(seq 1 ${REPS:-3}); do
    t0=$(date +%s.%N); "$@" > $D/$name.log 2>&1; t1=$(date +%s.%N)
    dt=$(python -c "print($t1-$t0)"); best=$(python -c "print(min($best,$dt))")
  done
  python -c "print('%-28s best of ${REPS:-3}: %.3f s  %.1f Mreads/s' % ('$name', $best, $READS/$best/1e6))"
}
run ref_t52   oracle/_ref/pandepth_ref -i $D/s.bam -o $D/ref "$@" -t 52
for t in ${THREADS:-16 52 128}; do
  run ours_t$t  pandepth_b200/bin/pandepth -i $D/s.bam -o $D/our "$@" -t $t
done
for f in $D/ref.*.gz; do b=$(basename $f); o=$D/our${b#ref}; if cmp -s <(zcat $f) <(zcat $o); then echo "identical: $b"; else echo "DIFFERENT: $b"; fi; done
