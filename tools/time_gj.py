import sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np
from mavmap_b200 import _lib
os.environ["MM_GJ_DEBUG"] = "1"
for m in (448, 1946):
    rng = np.random.default_rng(m); B = rng.normal(size=(m, m + 8)); A = B @ B.T + 0.5 * np.eye(m)
    for rep in range(2):
        out = np.ascontiguousarray(A.copy()); _lib.lib().mm_debug_spd_inverse(out.ctypes.data_as(C.POINTER(C.c_double)), m)
    print("err", np.abs(out @ A - np.eye(m)).max())
