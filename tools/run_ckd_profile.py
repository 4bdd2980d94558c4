"""Runs the run_ckd LW core at the bench shape a few times (profiling target): python tools/run_ckd_sweep.py does
the timing; this wrapper only exists so that `ncu ... python tools/run_ckd_profile.py` has a short command."""
import os
import runpy
import sys

sys.argv = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "run_ckd_sweep.py")] + sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
