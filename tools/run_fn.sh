run() { echo "== $*"; env $1 timeout 250 python bench.py --steps 6 --warmup 3 --no-cpu-baseline ${@:2} 2>&1 | tail -1 > /tmp/l.json; python -c "
import json
d=json.loads(open('/tmp/l.json').read()); k=d['roofline']['kernel_share_ms_per_step']
print(round(d['value'],2), round(d['ms_per_step'],3), {a:round(b,3) for a,b in k.items()})" || tail -5 /tmp/l.json; }
