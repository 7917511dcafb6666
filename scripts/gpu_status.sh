#!/bin/bash
# One GPU call: conv-engine + iteration parity, default-size bench (20 steps), optional extras.
B=${B:-1024}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_conv_engine_gpu.py tests/test_iteration_gpu.py tests/test_graphed_step_gpu.py -q -x 2>&1 | tail -3
timeout -k 10 900 python bench.py --batch $B --steps ${STEPS:-20} --warmup 3 --no_cpu_baseline 2>&1 | grep -v "Warn\|Consider\|run_backward" | tail -1 | tee gpurun_out/bench_b$B.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.0f videos/s  %.2f ms/step  e2e %.0f (%.2f ms)  launches/step %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches']/d['steps']))
for k,v in r['kernels'].items(): print('  %-22s %6.2f ms  %5.0f TF/s  n=%d' % (k, v['ms_per_step_in_kernel'], v['achieved'] or 0, v['launches_per_step']))
print('  conv engine %.2f ms %.0f TF/s; nominal step frac %.3f' % (r['conv_engine_all']['ms_per_step'], r['conv_engine_all']['achieved'], r['step_nominal_frac_of_sustained_peak']))
"
