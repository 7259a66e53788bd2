import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bellman_b200 as bb
n0, n1, C = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (128, 128, 16)))
o = bb.Dynamic_Solver()
d = bb.tables.kirk_desc(o.A, o.B, o.Q, o.R, 6, -2.5, 3.0, n0, -40.0, 10.0, C, store_J_all=False, store_idx_all=False)
d.n = [n0, n0]
sw = bb.Sweep(d)
sw.run(2, kernel=bb.KERNEL_WINDOW)
print("ok", sw.last_kernel, sw.get_J().sum())
