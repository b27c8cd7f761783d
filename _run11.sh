mkdir -p gpurun_out/r2k
( time python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r2k/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2k/summary.txt
( time python bench.py --steps 10 --warmup 3 --layers ) > gpurun_out/r2k/bench_auto.json 2> gpurun_out/r2k/bench_auto.err
echo "bench rc=$?" >> gpurun_out/r2k/summary.txt
cp gpurun_out/conv_layers.json gpurun_out/r2k/conv_layers_bf16x3.json
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2k/bench_ref.json 2> gpurun_out/r2k/bench_ref.err
echo "ref rc=$?" >> gpurun_out/r2k/summary.txt
python tools/bench_wgrad.py > gpurun_out/r2k/bench_wgrad.log 2>&1
python tools/bench_conv.py --dtype bf16x3 --n 8 > gpurun_out/r2k/bench_conv_bf16x3.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc --csv --log-file gpurun_out/r2k/conv_dram.csv python bench.py --dtype bf16x3 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2k/ncu_dram.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc -c 1 -o gpurun_out/r2k/prof_r02_fpn_p2_bf16x3 python tools/bench_conv.py --dtype bf16x3 --iters 1 --n 8 --only "fpn_posthoc_P2" > gpurun_out/r2k/ncu_full.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:wgrad -c 1 -o gpurun_out/r2k/prof_r02_wgrad_res4 python tools/bench_wgrad.py --iters 1 --only "res4 branch2b" > gpurun_out/r2k/ncu_wgrad.log 2>&1
cat gpurun_out/r2k/summary.txt; tail -6 gpurun_out/r2k/pytest_gpu.log; cat gpurun_out/r2k/smoke.log | tail -2; tail -4 gpurun_out/r2k/bench_auto.err; tail -3 gpurun_out/r2k/bench_ref.err; cat gpurun_out/r2k/bench_wgrad.log | tail -9; ls -la gpurun_out/r2k
