#!/bin/bash
# CTA pairs (cta_group::2) against single-CTA tiles on the pooled layers that used pairs until round 2's last pass:
# each layer alone (20 launches per graph, warm L2) at several batch sizes, then the whole bench step both ways.
for b in 1 8 64 256; do echo "BATCH $b  (first line of a pair: without CTA pairs; second: with)"; ITERS=40 BATCH=$b python scripts/b1_layers.py 2>&1 | grep -E "^conv[234]2|without"; done
for i in 1 2; do for pv in 1 0; do
  python - <<PY
import json, os, subprocess, sys
out = subprocess.run([sys.executable, 'bench.py', '--no-configs', '--no-dmha', '--no-extras', '--no-cpu-baseline'] + (['--pairs'] if '$pv' == '1' else []), capture_output=True, text=True).stdout
d = json.loads(out.strip().splitlines()[-1])
print('bench step, %s:' % ('conv22/32/42 on CTA pairs (bench.py --pairs)' if '$pv' == '1' else 'single-CTA tiles (default)'), round(d['value']), 'emb/s', round(d['ms_per_step'], 3), 'ms',
      {k: round(v) for k, v in d['roofline']['per_layer_tflops'].items()}, 'SM clock', d['clocks']['sm_mhz'])
PY
done; done
