#!/bin/bash
# usage: tools/ab_env.sh VAR "<bench args>" val1 val2 ...   (val "-" = unset)
var=$1; wl="$2"; shift 2
for v in "$@"; do
  if [ "$v" = "-" ]; then unset $var; else export $var=$v; fi
  python bench.py $wl --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip())
print('$var=$v', 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'])"
done
