#!/usr/bin/env bash
# e2e leg of bench.py with and without the chunked host transfers of svgdb_step_host
set -u
mkdir -p gpurun_out
for hc in 1 0; do
  SVGDB_HOST_CHUNKS=$hc python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/e2e_hc$hc.json 2> gpurun_out/e2e_hc$hc.err
  tail -1 gpurun_out/e2e_hc$hc.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('host chunks $hc:', 'ms/step %.3f' % d['ms_per_step'], 'e2e ms %.3f' % d['e2e']['ms_per_step'], 'finite', d['config']['finite'])"
done
