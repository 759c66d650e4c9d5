"""How close the BASELINE config-3 default-path parity test runs to its bounds: value / bound of every figure it checks,
largest first (tests/test_gpu_default_path.py::test_bc_config3_default_path_forward_backward_vs_fp64_oracle)."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import test_gpu_default_path as T
from hierarchicalgnn_b200 import ops
rows = []
orig = T.Errs.assert_ok
def report(self):
    rows.extend(self.rows); orig(self)
T.Errs.assert_ok = report
ops.set_precision("auto")
T.test_bc_config3_default_path_forward_backward_vs_fp64_oracle()
rows.sort(key=lambda r: -r[1] / r[2])
for n, v, b in rows[:12]:
    print(f"{v / b:6.3f} of its bound   {n:60s} {v:.3e} < {b:.1e}")
print(len(rows), "figures checked")
