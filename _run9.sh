mkdir -p gpurun_out/r2i
timeout 900 python -X faulthandler -m pytest tests/test_gpu_trainer.py tests/test_gpu_conv.py tests/test_gpu_engine.py -q -m gpu -k "trainer or f16 or bf16x3h" -s > gpurun_out/r2i/test_f16.log 2>&1
echo "f16+trainer rc=$?" >> gpurun_out/r2i/summary.txt
timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity_e2e.py -q -m gpu -s -k "bf16x3h" > gpurun_out/r2i/test_e2e.log 2>&1
echo "e2e rc=$?" >> gpurun_out/r2i/summary.txt
for m in bf16x3h bf16x3; do
  timeout 600 python bench.py --dtype $m --steps 10 --warmup 3 --no-cpu-baseline --no-extras --layers > gpurun_out/r2i/bench_$m.json 2> gpurun_out/r2i/bench_$m.err
  cp gpurun_out/conv_layers.json gpurun_out/r2i/conv_layers_$m.json
done
cat gpurun_out/r2i/summary.txt
grep -n "passed\|failed\|Error\|assert" gpurun_out/r2i/test_f16.log | head
grep -n "e2e parity\|passed\|failed\|Error" gpurun_out/r2i/test_e2e.log | head -20
for m in bf16x3h bf16x3; do python -c "
import json; d=json.load(open('gpurun_out/r2i/bench_$m.json')); print('$m', d['value'], d['e2e']['value'], d['ms_per_step'])"; tail -2 gpurun_out/r2i/bench_$m.err; done
