# e2e variance probe: default bench (b = 2048) with different prefetcher chunk sizes, back to back on one box
for mb in 48 8 48 8; do T2V_PF_CHUNK_MB=$mb python bench.py --steps 10 --no_cpu_baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk $mb MB: resident %.2f e2e %.2f ms  %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks']['reasons']))"; done
