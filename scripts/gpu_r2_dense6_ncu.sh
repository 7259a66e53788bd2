#!/bin/bash
# ncu capture of k_stage_dense6 on the 24^3 x 10^3 mesh (third launch of the timed loop)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > gpurun_out/d6.py <<'PY'
import bellman_b200 as bb
sa = bb.Solver_attitude(); sa.n_mesh_w, sa.n_mesh_q = 24, 10
T = sa.dense6_tables()
print(bb.dense6_run(T, 4)[2])
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_stage_dense6 -s 2 -c 1 -f -o gpurun_out/r02_dense6 \
  env PYTHONPATH=. python gpurun_out/d6.py > gpurun_out/dense6_ncu.log 2>&1
ncu -i gpurun_out/r02_dense6.ncu-rep --page raw --csv > gpurun_out/r02_dense6_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r02_dense6_raw.csv > gpurun_out/r02_dense6_summary.txt
cat gpurun_out/r02_dense6_summary.txt
