"""Same-box A/B of two builds of the library: `python scripts/ab_lib.py <path to .so> [bench.py arguments]` runs bench.py with
gyre_b200._native bound to that shared object (measurement scaffolding, not product code)."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gyre_b200 import _native as N  # noqa: E402

N._LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
