#!/bin/bash
O=gpurun_out
cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_final.log 2>&1; tail -2 $O/pytest_gpu_final.log
timeout 300 python bench.py > $O/bench_final.json 2> $O/bench_final.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_final.json"))
print("value %.0f GB/s, e2e %.1f (ceiling %.1f, frac %.2f), enc %.0f (%.3f) dec %.0f (%.3f), clocks %s" % (d["value"], d["e2e"]["value"], d["e2e"]["ceiling"], d["e2e"]["frac"], d["roofline"]["encode"]["GBps"], d["roofline"]["encode"]["frac"], d["roofline"]["decode"]["GBps"], d["roofline"]["decode"]["frac"], d["clocks"]))
for x in d["extra_configs"]: print(x["workload"][:32], round(x["encode"]["GBps"]), round(x["decode"]["GBps"]), x["parity"][:40])
print(d["parity_gate"])
PY
