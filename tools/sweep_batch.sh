#!/bin/bash
# throughput vs batch size (device-resident), one line per batch
for B in "$@"; do
  python bench.py --steps 200 --warmup 10 --latency-calls 0 --no-cpu-baseline --batch $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('B %6d value %6.1fM ms/step %.4f e2e %.1fM ctrl %.1fM' % (d['e2e']['h2d_bytes_per_step']//1300, d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e_controller']['value']/1e6))"
done
