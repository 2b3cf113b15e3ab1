#!/bin/bash
# K* launch-bounds experiment: the same bench with posterior.cu built at -DKV_LB=1 (default lib), 3 and 4.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > gpurun_out/kstar_lb_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/kstar_lb_tests.log
for v in b200 lb3 lb4; do
  BOCF_LIB_PATH=$PWD/bocf_b200/csrc/libbocf_$v.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mixed --no-extras > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  echo "variant=$v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_$v.json')); print(round(d['value']), round(d['ms_per_step'],1), {k:round(v/3,1) for k,v in d['roofline']['kernel_ms'].items()}, d['clocks']['sm_mhz'])"
done
