#!/bin/bash
# min over repetitions of each host phase of forest creation: tools/exporter_phases.sh [threads] [reps] [cfg]
HERE=$(cd "$(dirname "$0")" && pwd)
BARTINT_TIMING=1 python "$HERE/exporter_bench.py" --threads ${1:-1} --reps ${2:-7} --cfg ${3:-2} 2>&1 | python -c "
import sys, collections
m = collections.defaultdict(lambda: 1e9)
for l in sys.stdin:
    if l.startswith('[bartint]'):
        name = l[10:38].strip(); m[name] = min(m[name], float(l[38:].split()[0]))
    elif l.startswith('{'): print(l.strip())
for k, v in m.items(): print(f'{k:32s} {v:8.3f} ms')
print('sum of phase minima', round(sum(m.values()), 3))
"
