# experiment: the smallest config (configs[0], 10k records) is launch-latency bound -- which kernel variants shorten its chain?
for opts in "" "--cigar-kernel 2" "--cigar-kernel 1" "--k3-fold" "--cigar-kernel 2 --k3-fold" "--no-overlap" "--cigar-kernel 2 --no-overlap"; do
  python bench.py --config 0 --no-e2e-bam --no-cpu-baseline --steps 200 --warmup 10 $opts 2>/dev/null | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-40s step %.1f us  e2e %.3e' % ('$opts', d['ms_per_step']*1e3, d['e2e']['value']))"
done
