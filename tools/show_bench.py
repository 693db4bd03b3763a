#!/usr/bin/env python
"""Prints the headline fields of a bench.py JSON line (file argument)."""
import json
import sys
for fn in sys.argv[1:]:
    try:
        d = json.loads(open(fn).read().strip().splitlines()[-1])
    except Exception as e:
        print(fn, 'unreadable:', e)
        continue
    st = {k: round(v, 4) for k, v in d.get('stage_ms', {}).items() if isinstance(v, float)}
    print('{}: {:.0f} img/s  {:.4f} ms/step (eager {:.4f})  path_frac {:.3f}  stages {}'.format(
        d['config']['workload'][:12], d['value'], d['ms_per_step'], d.get('ms_per_step_eager', 0),
        d['roofline']['path_frac'], st))
