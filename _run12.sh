mkdir -p gpurun_out/r2m
( time python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r2m/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2m/summary.txt
( time python bench.py --steps 10 --warmup 3 --layers ) > gpurun_out/r2m/bench_auto.json 2> gpurun_out/r2m/bench_auto.err
echo "bench rc=$?" >> gpurun_out/r2m/summary.txt
cp gpurun_out/conv_layers.json gpurun_out/r2m/conv_layers_bf16x3.json
( time python bench.py --train --steps 5 --warmup 3 ) > gpurun_out/r2m/train1.json 2> gpurun_out/r2m/train1.err
echo "train rc=$?" >> gpurun_out/r2m/summary.txt
cat gpurun_out/r2m/summary.txt; tail -6 gpurun_out/r2m/pytest_gpu.log; tail -2 gpurun_out/r2m/smoke.log; tail -4 gpurun_out/r2m/bench_auto.err; cat gpurun_out/r2m/train1.json
