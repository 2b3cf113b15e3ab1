for g in 4 16 48; do
  BOCF_SCRATCH_GIB=$g python bench.py --config cfg5 --candidates 200000 --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-mixed 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('cfg5 scratch $g GiB', round(l['value']), 'evals/s', round(l['ms_per_step'],1),'ms', {k:round(v,1) for k,v in l['roofline'].get('kernel_ms',{}).items()})"
done
for g in 4 12; do
  BOCF_SCRATCH_GIB=$g python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-extras --no-mixed 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('cfg3 scratch $g GiB', round(l['value']), 'evals/s', round(l['ms_per_step'],1),'ms', {k:round(v,1) for k,v in l['roofline'].get('kernel_ms',{}).items()})"
done
