# parts per candidate tile at cfg 5: does sharing a unit's A tile between CTAs keep its re-streaming in L2?
# (library default = automatic rule, split_gemm.cu: parts_for; BOCF_SPLIT_NP overrides)
for np in auto 8; do
  if [ $np = auto ]; then unset BOCF_SPLIT_NP; else export BOCF_SPLIT_NP=$np; fi
  python bench.py --config cfg5 --candidates 200000 --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-mixed 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('cfg5 NP=$np', round(l['value']), 'evals/s', round(l['ms_per_step'],1),'ms', {k:round(v,1) for k,v in l['roofline'].get('kernel_ms',{}).items()})"
done
unset BOCF_SPLIT_NP
python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-extras --no-mixed 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('cfg3 NP=auto', round(l['value']), 'evals/s', round(l['ms_per_step'],1),'ms', {k:round(v,1) for k,v in l['roofline'].get('kernel_ms',{}).items()})"
