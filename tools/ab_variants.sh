#!/bin/bash
# A/B of the board sweep switches on the full game (run on the GPU box after tools/build_variants.py): tools/ab_variants.sh [variants...]
# writes gpurun_out/ab_variants.jsonl (one board_probe line per variant)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=${@:-base final}
: > gpurun_out/ab_variants.jsonl
for v in $V; do
  echo -n "{\"variant\": \"$v\", \"probe\": " >> gpurun_out/ab_variants.jsonl
  PRL_LIB_PATH=pokerrl_b200/lib/variants/lib_$v.so timeout 120 python tools/board_probe.py 134459 10 2>gpurun_out/ab_$v.err | tail -1 >> gpurun_out/ab_variants.jsonl
  echo "}" >> gpurun_out/ab_variants.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/ab_variants.jsonl'):
    try:
        d=json.loads(l); p=d['probe']; print(d['variant'], 'sweep_ms %.3f'%p['sweep_ms'], 'it/s %.2f'%p['iterations_per_s'], 'GB/s %.0f'%p['sweep_GBps'], 'expl', p['expl_cur'], p['expl_avg'])
    except Exception as e: print('bad line', l[:200])
PY
