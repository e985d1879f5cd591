#!/bin/bash
# usage: tools/bench_sweep.sh "ENV1=a ENV2=b" "ENV1=c" ...   -> one short bench line per setting
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --no-cpu --no-e2e --steps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), d['fit'])"
done
