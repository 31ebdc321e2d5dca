#!/bin/bash
# development: C3 under a few plan options
for o in '{"cta_log2":8}' '{"cta_log2":7}' '{"cta_log2":8,"max_dense_ops":40}' '{"tile_bits":11,"cta_log2":7}'; do
  echo -n "$o  "
  C3_OPTS="$o" timeout 100 python tools/c3_bench.py 2>&1 | grep '"C3"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['plan']['n_passes'], d['plan']['n_steps'], round(d['ms_total'],2))"
done
