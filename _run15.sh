mkdir -p gpurun_out/r2p
for f in test_gpu_train_ops test_gpu_trainer test_gpu_tools; do
  timeout 900 python -X faulthandler -m pytest tests/$f.py -q -m gpu -x > gpurun_out/r2p/$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/r2p/summary.txt
done
( time python bench.py --train --steps 10 --warmup 3 ) > gpurun_out/r2p/train1.json 2> gpurun_out/r2p/train1.err
echo "train rc=$?" >> gpurun_out/r2p/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p/launches_train.csv python bench.py --train --steps 1 --warmup 3 > gpurun_out/r2p/ncu_train.log 2>&1
cat gpurun_out/r2p/summary.txt
for f in gpurun_out/r2p/test*.log; do echo "== $f"; tail -3 $f; done
cat gpurun_out/r2p/train1.json; tail -3 gpurun_out/r2p/train1.err
