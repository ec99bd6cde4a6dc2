#!/bin/bash
# one GPU call: cluster-mode N4 sanity (bounded), timings + bit parity, the GPU suite, smoke, a short bench.
# If the sanity pass fails, cluster mode is switched off (VB_N4_CLUSTER=1) for the rest of the call.
O=gpurun_out/r02c; mkdir -p $O
if timeout 70 python tools/time_cluster.py --quick > $O/cluster_quick.jsonl 2> $O/cluster_quick.err; then
  echo "quick ok"; cat $O/cluster_quick.jsonl
  timeout 100 python tools/time_cluster.py > $O/cluster.jsonl 2> $O/cluster.err; echo "cluster timing rc=$?"
else
  echo "quick FAILED rc=$? -> cluster mode off for the rest"; tail -5 $O/cluster_quick.err; export VB_N4_CLUSTER=1
fi
timeout 200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cohort-leg > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
cat $O/cluster.jsonl
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02c/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step") if k in d}, d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"))
PY
