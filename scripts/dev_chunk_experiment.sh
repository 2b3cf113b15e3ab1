# chunk size vs wave quantisation (K*: 2 blocks of 128 candidates per SM = 296 slots; contractions: 148 persistent CTAs)
for g in 3.6 3.8 3.95 4.0 4.4 5.2; do
  BOCF_SCRATCH_GIB=$g python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-mixed 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); k=l['roofline']['kernel_ms']; n=l['roofline']['launches']
print('scratch $g GiB', round(l['value']), round(l['ms_per_step'],1),'ms launches/step', n/3.0/1.0, {a:round(b/3,1) for a,b in k.items()})"
done
