# e2e probe: default bench (b = 2048) with the prefetcher's bulk copy on the copy engine (0) or on N resident CTAs
for c in 0 16 48 148; do T2V_PF_SM_COPY_CTAS=$c python bench.py --steps 10 --no_cpu_baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('sm-copy ctas $c: res %.1f  e2e %.1f  e2e_u8 %.1f  res_again %.1f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_uint8']['ms_per_step'], d['resident_again_ms_per_step']))"; done
