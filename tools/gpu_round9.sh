#!/bin/bash
# slot-scheduled kernel, burst scheduler: parity, then sweeps
set -u
mkdir -p gpurun_out
echo "=== pytest slot tests"
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "slot" 2>&1 | tail -8
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); s=d['config'].get('sched'); print('$1: %.0f Mrays/s  ms/step %.3f' % (d['value'], d['ms_per_step']), ' '.join('%s %.1fM@%.1f'%(k[:5]+k[-2:],v[0]/1e6,v[1]) for k,v in s.items()) if s else '')
except Exception as e: print('$1 FAILED', l[-400:])
"; }
for LEAF in 3 2 4; do
B="timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --leaf-size $LEAF"
for cfg in "3 768" "2 768" "3 512" "4 512"; do set -- $cfg
  $B --kernel slots --opt slot_slots=$1 --opt slot_threads=$2 2>&1 | show "leaf$LEAF slots K=$1 T=$2"
done
done
B="timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --leaf-size 3"
for tn in 8 12 16 24; do $B --kernel slots --opt slot_tn=$tn 2>&1 | show "leaf3 3x768 TN=$tn"; done
for tl in 4 8 16; do $B --kernel slots --opt slot_tl=$tl 2>&1 | show "leaf3 3x768 TL=$tl"; done
for tw in 4 12 16; do $B --kernel slots --opt slot_tw=$tw 2>&1 | show "leaf3 3x768 TW=$tw"; done
for ts in 8 12 16 24; do $B --kernel slots --opt slot_ts=$ts --opt slot_tr=$ts 2>&1 | show "leaf3 3x768 TS=TR=$ts"; done
