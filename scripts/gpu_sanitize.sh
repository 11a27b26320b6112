#!/bin/bash
# compute-sanitizer (memcheck + racecheck + initcheck) on small inputs
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from swarm_b200 import Engine, HostDb, ENUM_FULL, ENUM_HALF, ENUM_JOIN, scoring
for name in ("c1_1k_150", "short_600_20", "w65_300"):
    db = HostDb(f"tests/golden/{name}.fasta")
    for mode in (ENUM_FULL, ENUM_HALF, ENUM_JOIN):
        for fk in (1, 2):
            e = Engine(0, enum_mode=mode, fast_kernel=fk, collect_stats=1)
            e.load(db); e.d1_index(); e.d1_network(); e.d1_cluster(); e.d1_fastidious(); e.close()
    e = Engine(0); e.load(db); e.dn_cluster(2, penalties=scoring()); e.close()
    e = Engine(0, dn_filter=1); e.load(db); e.dn_cluster(3, penalties=scoring()); e.close()
print("sanitizer workload done")
PY
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > $O/sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitizer workload done' $O/sanitizer_$tool.log | tr '\n' ' ')"
done
