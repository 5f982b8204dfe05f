#!/bin/bash
# GPU session r02d: stratified shearing box, MRI fused A/B (instruction-cache diet), tests
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -n 3 > $O/r02d_pytest.log 2>&1; tail -25 $O/r02d_pytest.log
echo "== MRI A/B"; timeout 300 python tools/mri_ab.py > $O/r02d_mri_ab.log 2>&1; cat $O/r02d_mri_ab.log
echo "== ncu MRI fused (rot_dt=0)"
cat > /tmp/mri_prof.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from ramsesgpu_b200 import set_tuning
set_tuning("rot_dt", int(sys.argv[1]))
sys.argv = ["full_size_check.py", "mri256slab"]
exec(open("tools/full_size_check.py").read())
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_flux -s 3 -c 1 -o $O/r02d_mri_fused_nodt python /tmp/mri_prof.py 0 > $O/r02d_ncu_mri0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_flux -s 3 -c 1 -o $O/r02d_mri_fused_dt python /tmp/mri_prof.py 1 > $O/r02d_ncu_mri1.log 2>&1
tail -2 $O/r02d_ncu_mri0.log
