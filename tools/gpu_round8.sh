#!/bin/bash
# slot-scheduled kernel: parity test first, then a sweep of slot geometries and thresholds
set -u
mkdir -p gpurun_out
echo "=== pytest slot + wide tests"
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "slot or wide or octant or c1_exact" 2>&1 | tail -15
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: value %.0f Mrays/s  e2e %.0f  ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['config'].get('sched'))
except Exception as e: print('$1 FAILED', l[-400:])
"; }
B="timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --leaf-size ${LEAF:-3}"
$B 2>&1 | show "persistent wide leaf${LEAF:-3}"
$B --opt wide_nodes=0 2>&1 | show "persistent pairs-oct leaf${LEAF:-3}"
for cfg in "3 768" "2 1024" "4 512" "2 768" "3 512"; do set -- $cfg
  $B --kernel slots --opt slot_slots=$1 --opt slot_threads=$2 2>&1 | show "slots K=$1 T=$2"
done
for tn in 12 16 24 28; do $B --kernel slots --opt slot_tn=$tn 2>&1 | show "slots 3x768 TN=$tn"; done
for tl in 4 8 16 20; do $B --kernel slots --opt slot_tl=$tl 2>&1 | show "slots 3x768 TL=$tl"; done
for tw in 4 12 16; do $B --kernel slots --opt slot_tw=$tw 2>&1 | show "slots 3x768 TW=$tw"; done
for ts in 12 16 24 28; do $B --kernel slots --opt slot_ts=$ts --opt slot_tr=$ts 2>&1 | show "slots 3x768 TS=TR=$ts"; done
