for lz in 3 4; do echo "LZ=$lz"; ECCKD_RUN_CKD_LZ=$lz python tools/run_ckd_profile.py 2>&1 | tail -1 | cut -c1-120; done
