mkdir -p gpurun_out/r2s
python tools/bench_wgrad.py --planes > gpurun_out/r2s/bench_wgrad.log 2>&1
echo "bench_wgrad rc=$?" >> gpurun_out/r2s/summary.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc --csv --log-file gpurun_out/r2s/conv_dram.csv python bench.py --dtype bf16x3 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2s/ncu_dram.log 2>&1
echo "ncu dram rc=$?" >> gpurun_out/r2s/summary.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc -c 1 -o gpurun_out/r2s/prof_r02_fpn_p2_bf16x3 python tools/bench_conv.py --dtype bf16x3 --iters 1 --n 8 --only "fpn_posthoc_P2" > gpurun_out/r2s/ncu_full_conv.log 2>&1
echo "ncu conv rc=$?" >> gpurun_out/r2s/summary.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:wgrad_nhwc -c 1 -o gpurun_out/r2s/prof_r02_wgrad_nhwc_p2 python tools/bench_wgrad.py --iters 1 --only "fpn posthoc P2" > gpurun_out/r2s/ncu_full_wgrad.log 2>&1
echo "ncu wgrad rc=$?" >> gpurun_out/r2s/summary.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:wgrad_nhwc -c 1 -o gpurun_out/r2s/prof_r02_wgrad_nhwc_res4 python tools/bench_wgrad.py --iters 1 --only "res4 branch2b" > gpurun_out/r2s/ncu_full_wgrad2.log 2>&1
( time python bench.py --train --steps 10 --warmup 3 ) > gpurun_out/r2s/train1.json 2> gpurun_out/r2s/train1.err
echo "train rc=$?" >> gpurun_out/r2s/summary.txt
cat gpurun_out/r2s/summary.txt; cat gpurun_out/r2s/bench_wgrad.log; ls -la gpurun_out/r2s; tail -1 gpurun_out/r2s/train1.json | cut -c1-900
