for e in ${KO_LIST:-0 1 2 4}; do
  BOCF_LIB_PATH=$PWD/bocf_b200/csrc/libbocf_ko.so BOCF_SPLIT_EXP=$e timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-mixed > gpurun_out/bench_ko.json 2> gpurun_out/bench_ko.err
  echo "EXP=$e rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_ko.json')); print(round(d['value']), round(d['ms_per_step'],1), {k:round(v/3,1) for k,v in d['roofline']['kernel_ms'].items()}, d['clocks']['sm_mhz'])"
done
