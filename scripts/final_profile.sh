# Final-state evidence of a round (one B200): full GPU test suite, the default bench line, the ncu launch list of the same
# command and one --set full capture of the three hot kernels.  Outputs under gpurun_out/ (copied to profiles/ by hand).
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/final_tests.log
python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-mixed > gpurun_out/final_plain.json 2>> gpurun_out/final_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-mixed > gpurun_out/final_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'split_gemm_kernel|kstar_kernel' -s 30 -c 3 -f -o gpurun_out/final_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-mixed > gpurun_out/final_ncu2.log 2>&1
ls -la gpurun_out/final_*
