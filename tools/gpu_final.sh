#!/bin/bash
# Round-end GPU visit: parity tests, the bench lines the driver will ask for, and the ncu launch list of the bench command.
set -u
O=gpurun_out; mkdir -p $O; T=${1:-final}
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $O/${T}_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/${T}_bench_u20.json 2> $O/${T}_bench_u20.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
timeout 300 python bench.py --workload N22 --steps 3 --warmup 3 > $O/${T}_bench_n22.json 2> $O/${T}_bench_n22.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --workers 1 --no-cpu-baseline > $O/${T}_launches.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${T}_smoke.log 2>&1
tail -2 $O/${T}_pytest.log; tail -1 $O/${T}_smoke.log; wc -l $O/${T}_launches.csv
python - <<PY
import json
for f in ("bench_u20", "bench_reference", "bench_n22"):
    try:
        for l in open("$O/${T}_%s.json" % f):
            if l.startswith("{"):
                d = json.loads(l); print(f, d.get("value"), d.get("unit"), (d.get("e2e") or {}).get("value"), d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
