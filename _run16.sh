mkdir -p gpurun_out/r2q
timeout 600 python -X faulthandler -m pytest tests/test_gpu_train_ops.py -q -m gpu > gpurun_out/r2q/train_ops.log 2>&1
echo "train_ops rc=$?" >> gpurun_out/r2q/summary.txt
timeout 600 python -X faulthandler -m pytest tests/test_gpu_trainer.py -q -m gpu -x > gpurun_out/r2q/trainer.log 2>&1
echo "trainer rc=$?" >> gpurun_out/r2q/summary.txt
( timeout 300 python bench.py --train --steps 10 --warmup 3 ) > gpurun_out/r2q/train_nhwc.json 2> gpurun_out/r2q/train_nhwc.err
echo "train nhwc rc=$?" >> gpurun_out/r2q/summary.txt
( DT_WGRAD_PLANES=1 timeout 300 python bench.py --train --steps 10 --warmup 3 ) > gpurun_out/r2q/train_planes.json 2> gpurun_out/r2q/train_planes.err
echo "train planes rc=$?" >> gpurun_out/r2q/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2q/launches_train.csv python bench.py --train --steps 1 --warmup 3 > gpurun_out/r2q/ncu_train.log 2>&1
cat gpurun_out/r2q/summary.txt; tail -5 gpurun_out/r2q/train_ops.log; tail -3 gpurun_out/r2q/trainer.log
python -c "
import json
for f in ('nhwc','planes'):
    try:
        d=json.loads(open('gpurun_out/r2q/train_%s.json'%f).read().strip().split('\n')[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('loss_kps'))
    except Exception as e: print(f, 'ERR', e)
"
