#!/bin/bash
# last GPU call of round 2: the GPU suite, smoke, a bench line, what the automatic N4 cluster choice picks,
# single-subject wall times through the host entry point, one ncu capture of the cluster kernel (C4 subject)
O=gpurun_out/r02f; mkdir -p $O
timeout 200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cohort-leg > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 80 python tools/time_cluster.py --auto > $O/cluster_auto.jsonl 2> $O/cluster_auto.err; echo "auto rc=$?"; cat $O/cluster_auto.jsonl
timeout 60 python tools/time_big.py > $O/single_subject.txt 2> $O/single_subject.err; echo "time_big rc=$?"; cat $O/single_subject.txt
timeout 120 ncu --set full --clock-control none --import-source on -k regex:n4_cluster_kernel -c 1 -f -o $O/r02_n4_cluster_c4 \
  python tools/time_cluster.py --ncu > $O/ncu_cluster.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu_cluster.log
CW=14,18,22 timeout 60 python tools/sweep_cluster_chain.py > $O/sweep_fine.jsonl 2> $O/sweep_fine.err; echo "fine sweep rc=$?"; cat $O/sweep_fine.jsonl
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02f/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step") if k in d}, d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"))
PY
