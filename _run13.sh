mkdir -p gpurun_out/r2n
for f in test_gpu_targets test_gpu_train_heads test_gpu_trainer test_gpu_tracking; do
  timeout 900 python -X faulthandler -m pytest tests/$f.py -q -m gpu -s -x > gpurun_out/r2n/$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/r2n/summary.txt
done
cat gpurun_out/r2n/summary.txt
for f in gpurun_out/r2n/test*.log; do echo "== $f"; grep -n "passed\|failed\|Error\|error\|assert\|losses\|grad (max" $f | head -30; done
