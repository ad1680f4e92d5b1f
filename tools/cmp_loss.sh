#!/bin/bash
# final KD loss of the bench's training run under the A/B switches of the gradient path (same data, same seeds)
run() { env "$@" python bench.py --workload kd_train --steps 10 --warmup 3 > /tmp/l.log 2>&1; tail -1 /tmp/l.log > /tmp/l.json
  python -c "import json,sys; d=json.load(open('/tmp/l.json')); print(' '.join(sys.argv[1:]) or 'default', d['ms_per_step'], d['config']['final_loss'], 'graph', d['config']['cuda_graph'])" "$@" || tail -5 /tmp/l.log; }
for a in "$@"; do run $a; done
