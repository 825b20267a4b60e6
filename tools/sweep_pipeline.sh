#!/bin/bash
# chunk / worker-stream sweep of the pipelined sequence call (authoring aid; run on the GPU box)
for cs in "64 2" "64 3" "64 4" "128 2" "128 3" "128 4" "256 2" "32 4" "96 4"; do
  set -- $cs
  python bench.py --steps 20 --warmup 3 --cpu-sample 0 --chunk $1 --streams $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('chunk $1 streams $2: value %.0f fps (%.2f ms)  e2e %.0f fps (%.2f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done
