# end-to-end C2 step from pinned host memory: chunk size of the staging ring
for mb in 64 32 128; do
  echo "== OXLI_B200_CHUNK_MB=$mb"
  OXLI_B200_CHUNK_MB=$mb timeout 200 python bench.py --no-cpu-baseline --no-parity --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value %.2f G/s  %.2f ms | e2e %.2f G/s  %.2f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e'].get('ms_per_step', 0)), d['roofline']['pass_ms_per_step'])"
done
