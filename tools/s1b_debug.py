"""Debug build of the retirement image (-DEGDST_DEBUG_LATE) run on the S1b case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["EGDST_NVCC_EXTRA"] = "-DEGDST_DEBUG_LATE"
os.environ["EGDST_B200_LIB_DIR"] = "/tmp/_dbglib"
from egdst_b200 import examples
m = examples.retirement(T=40, ngridm=2000, ngridmax=4000, nthrhmax=2000, ny=20)
m.compile(); m.solve()
print("status", m._solution.status())
