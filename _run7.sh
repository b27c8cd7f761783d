mkdir -p gpurun_out/r2g
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/debug_wgrad.py > gpurun_out/r2g/debug_wgrad.log 2>&1
for f in test_gpu_train_ops test_gpu_trainer test_gpu_parity_e2e; do
  timeout 900 python -X faulthandler -m pytest tests/$f.py -q -m gpu -s > gpurun_out/r2g/$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/r2g/summary.txt
done
timeout 600 python bench.py --train --steps 5 --warmup 3 > gpurun_out/r2g/train1.json 2> gpurun_out/r2g/train1.err
echo "train rc=$?" >> gpurun_out/r2g/summary.txt
cat gpurun_out/r2g/summary.txt; cat gpurun_out/r2g/debug_wgrad.log | head -20
for f in gpurun_out/r2g/test*.log; do echo "== $f"; grep -n "e2e parity\|passed\|failed\|Error\|max grad errs\|assert" $f | head -20; done
cat gpurun_out/r2g/train1.json; tail -3 gpurun_out/r2g/train1.err
