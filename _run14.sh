mkdir -p gpurun_out/r2o
for f in test_gpu_train_heads test_gpu_trainer test_gpu_tools; do
  timeout 900 python -X faulthandler -m pytest tests/$f.py -q -m gpu -s -x > gpurun_out/r2o/$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/r2o/summary.txt
done
( time python bench.py --train --steps 5 --warmup 3 ) > gpurun_out/r2o/train1.json 2> gpurun_out/r2o/train1.err
echo "train rc=$?" >> gpurun_out/r2o/summary.txt
( time python bench.py --train --train-trunk --steps 5 --warmup 3 ) > gpurun_out/r2o/train_trunk.json 2> gpurun_out/r2o/train_trunk.err
echo "trunk rc=$?" >> gpurun_out/r2o/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2o/launches_train.csv python bench.py --train --steps 1 --warmup 3 > gpurun_out/r2o/ncu_train.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r2o/summary.txt
cat gpurun_out/r2o/summary.txt
for f in gpurun_out/r2o/test*.log; do echo "== $f"; grep -n "passed\|failed\|Error\|error\|assert\|losses" $f | head -20; done
cat gpurun_out/r2o/train1.json; tail -5 gpurun_out/r2o/train1.err; cat gpurun_out/r2o/train_trunk.json
