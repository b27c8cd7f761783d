mkdir -p gpurun_out/r2t
for f in test_gpu_targets test_gpu_trainer test_gpu_tools test_gpu_train_ops; do
  timeout 900 python -X faulthandler -m pytest tests/$f.py -q -m gpu -x > gpurun_out/r2t/$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/r2t/summary.txt
done
( timeout 300 python bench.py --train --steps 10 --warmup 3 ) > gpurun_out/r2t/train1.json 2> gpurun_out/r2t/train1.err
echo "train rc=$?" >> gpurun_out/r2t/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t/launches_train.csv python bench.py --train --steps 1 --warmup 3 > gpurun_out/r2t/ncu_train.log 2>&1
cat gpurun_out/r2t/summary.txt
for f in gpurun_out/r2t/test*.log; do echo "== $f"; tail -4 $f; done
tail -1 gpurun_out/r2t/train1.json | cut -c1-700
