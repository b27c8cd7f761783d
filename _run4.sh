set -x
mkdir -p gpurun_out/r2d
python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2d/tests.log
for pdl in 1 0; do
  DT_PDL=$pdl timeout 600 python bench.py --dtype bf16x3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2d/bench_bf16x3_pdl$pdl.json 2> gpurun_out/r2d/bench_bf16x3_pdl$pdl.err
  DT_PDL=$pdl timeout 600 python bench.py --dtype bf16 --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2d/bench_bf16_pdl$pdl.json 2> gpurun_out/r2d/bench_bf16_pdl$pdl.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r2d/launches_bf16x3.csv python bench.py --dtype bf16x3 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2d/ncu_bench.log 2>&1
tail -8 gpurun_out/r2d/tests.log
