mkdir -p gpurun_out/r2e
for f in test_gpu_train_ops test_gpu_trainer test_gpu_api test_gpu_parity_e2e test_gpu_tools test_gpu_engine_tube; do
  timeout 900 python -X faulthandler -m pytest tests/$f.py -q -m gpu -x -s > gpurun_out/r2e/$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/r2e/summary.txt
done
timeout 900 python -X faulthandler -m pytest tests -q -m gpu --deselect tests/test_gpu_train_ops.py --deselect tests/test_gpu_trainer.py --deselect tests/test_gpu_api.py --deselect tests/test_gpu_parity_e2e.py --deselect tests/test_gpu_tools.py --deselect tests/test_gpu_engine_tube.py > gpurun_out/r2e/rest.log 2>&1
echo "rest rc=$?" >> gpurun_out/r2e/summary.txt
timeout 600 python bench.py --train --steps 5 --warmup 3 > gpurun_out/r2e/train1.json 2> gpurun_out/r2e/train1.err
echo "train rc=$?" >> gpurun_out/r2e/summary.txt
cat gpurun_out/r2e/summary.txt
for f in gpurun_out/r2e/*.log; do echo "== $f"; tail -5 $f; done
