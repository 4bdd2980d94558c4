#!/bin/sh
# Warps per column of run_ckd_kernel (timing builds only: ECCKD_TIMING_KNOBS=1 python -c "import __graft_entry__ as g; g.build_library(True)").
for lz in 1 2 3 4; do echo "LZ=$lz"; ECCKD_RUN_CKD_LZ=$lz python tools/run_ckd_profile.py 2>&1 | tail -1 | cut -c1-120; done
