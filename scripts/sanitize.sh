#!/bin/bash
# compute-sanitizer passes over both kernels through the CLI (pure CUDA host, fast under the tool):
# a graph with short rows, empty rows, long rows (> 4096) and a huge row (>= 32768, cluster path).
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import __graft_entry__ as e
e.load_package()
from gespmm_b200 import graphs
rng = np.random.default_rng(0)
M = 3000
deg = rng.integers(0, 9, M); deg[rng.random(M) < 0.3] = 0
deg[[5, 700, 701, 2999]] = [40000, 5000, 4097, 9000]
rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
colind = rng.integers(0, M, rowptr[-1]).astype(np.int32)
graphs.write_mtx('gpurun_out/sanitize.mtx', rowptr, colind)
print('rows', M, 'nnz', rowptr[-1])
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  compute-sanitizer --tool $tool --error-exitcode 9 ge-spmm_b200/bin/spmm_test gpurun_out/sanitize.mtx 0 --K 128,200,512 --iters 2 --validate --out gpurun_out/sanitize.csv 2>&1 | grep -E "ERROR SUMMARY|validate|WA|Error|error" | head -12
done
