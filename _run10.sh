mkdir -p gpurun_out/r2j
nvidia-smi -L > gpurun_out/r2j/gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --train --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2j/train2.json 2> gpurun_out/r2j/train2.err
echo "train2 rc=$?" >> gpurun_out/r2j/summary.txt
python bench.py --train --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2j/train1.json 2> gpurun_out/r2j/train1.err
echo "train1 rc=$?" >> gpurun_out/r2j/summary.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2j/infer2.json 2> gpurun_out/r2j/infer2.err
echo "infer2 rc=$?" >> gpurun_out/r2j/summary.txt
cat gpurun_out/r2j/summary.txt gpurun_out/r2j/train2.json gpurun_out/r2j/train1.json; tail -3 gpurun_out/r2j/train2.err; python -c "
import json; d=json.load(open('gpurun_out/r2j/infer2.json')); print('infer2', d['value'], d['e2e']['value'], d['n_gpus'])"
