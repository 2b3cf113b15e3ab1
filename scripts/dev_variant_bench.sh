#!/bin/bash
# bench-only comparison of library variants: VARIANTS="b200 rep1 ..." -> bocf_b200/csrc/libbocf_<v>.so
mkdir -p gpurun_out
for v in ${VARIANTS:-b200}; do
  BOCF_LIB_PATH=$PWD/bocf_b200/csrc/libbocf_$v.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mixed --no-extras > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  echo "variant=$v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_$v.json')); print(round(d['value']), round(d['ms_per_step'],1), {k:round(v/3,1) for k,v in d['roofline']['kernel_ms'].items()}, d['clocks']['sm_mhz'])"
done
