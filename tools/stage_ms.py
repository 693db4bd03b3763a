#!/usr/bin/env python
"""Print the interesting fields of bench.py JSON lines read from stdin."""
import json
import sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith('{'):
        continue
    d = json.loads(line)
    r = d.get('roofline', {})
    print('value={:.0f} img/s  ms/step={:.3f}  e2e={:.0f}  collect: {:.1f} GB/s frac={:.3f}  path_frac={:.3f}  stages={}  valid={}'.format(
        d['value'], d['ms_per_step'], d['e2e']['value'], r.get('achieved') or 0, r.get('frac') or 0,
        r.get('path_frac') or 0, {k: round(v, 4) for k, v in d.get('stage_ms', {}).items() if k != 'note'},
        d.get('mean_valid_detections')))
