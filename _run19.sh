mkdir -p gpurun_out/r2v
( time python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r2v/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2v/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2v/summary.txt
( time python bench.py --steps 10 --warmup 3 --layers ) > gpurun_out/r2v/bench_auto.json 2> gpurun_out/r2v/bench_auto.err
echo "bench rc=$?" >> gpurun_out/r2v/summary.txt
cp gpurun_out/conv_layers.json gpurun_out/r2v/conv_layers_bf16x3.json
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2v/bench_ref.json 2> gpurun_out/r2v/bench_ref.err
echo "ref rc=$?" >> gpurun_out/r2v/summary.txt
cat gpurun_out/r2v/summary.txt; tail -6 gpurun_out/r2v/pytest_gpu.log; tail -2 gpurun_out/r2v/smoke.log; tail -3 gpurun_out/r2v/bench_auto.err; tail -3 gpurun_out/r2v/bench_ref.err; cut -c1-600 gpurun_out/r2v/bench_ref.json
