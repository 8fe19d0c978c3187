"""mink_b200 -- batched differential inverse kinematics on NVIDIA B200 behind mink's Python API.

Only the `solve_ik` hot path of kevinzakka/mink is rebuilt here: FK + task
errors/Jacobians + QP assembly + exact active-set solve + integrate, as hand-written sm_100a kernels
in libbik.so (include/bik.h), driven through ctypes.  See DESIGN.md and INTEGRATION.md.
"""

__version__ = "0.1.0"
