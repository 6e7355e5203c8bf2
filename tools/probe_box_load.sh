#!/bin/bash
# One GPU of the box alone, then all of them at once, the SAME kernel (128-plane slab of the 1024^3 PD3O-TV iteration, no
# communication), with clocks / power sampled meanwhile: how much of the N-GPU step time is the box, not the exchange.
N=${1:-8}
smi() { nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 200 > "$1" & echo $!; }
echo "== GPU 0 alone"
P=$(smi gpurun_out/box_alone_clocks.csv)
python tools/bench_slab_shape.py --planes 128,128 --reps 1500 --spin 500 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rank', d['rank'], [round(r['ms'],4) for r in d['rows']])"
kill $P
echo "== $N GPUs at once"
P=$(smi gpurun_out/box_all_clocks.csv)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 tools/bench_slab_shape.py --planes 128,128 --reps 1500 --spin 500 2>/dev/null | grep "^{" | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('rank', d['rank'], [round(r['ms'],4) for r in d['rows']])" | sort
kill $P
python - <<'PY'
import csv, statistics, collections
for name in ("alone", "all"):
    rows = [r for r in csv.reader(open(f"gpurun_out/box_{name}_clocks.csv")) if len(r) >= 7]
    by = collections.defaultdict(list)
    for r in rows:
        by[r[0].strip()].append(r)
    print(f"-- clocks while '{name}' (median over the samples with power > 300 W):")
    for g, rs in sorted(by.items()):
        busy = [r for r in rs if float(r[3].split()[0]) > 300] or rs
        sm = statistics.median(float(r[1].split()[0]) for r in busy)
        mem = statistics.median(float(r[2].split()[0]) for r in busy)
        pw = statistics.median(float(r[3].split()[0]) for r in busy)
        cap = sum(r[4].strip() == "Active" for r in busy)
        print(f"   gpu {g}: sm {sm:.0f} MHz, mem {mem:.0f} MHz, {pw:.0f} W, sw_power_cap active in {cap}/{len(busy)} samples, hw_slowdown {sum(r[5].strip()=='Active' for r in busy)}, sw_thermal {sum(r[6].strip()=='Active' for r in busy)}")
PY
