"""ORACLE / TEST INFRASTRUCTURE -- run the reference's OWN tests on top of the oracle shims.

Puts `oracle/` (mujoco, qpsolvers, robot_descriptions stand-ins) and /root/reference on
sys.path, then runs /root/reference/tests with pytest.  This is the gate that pins the
restated third-party arithmetic.  Only runs where /root/reference
exists (this container); the GPU box uses the committed golden vectors instead.

Usage: python oracle/run_reference_tests.py [pytest args]
"""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MINK_REFERENCE", "/root/reference")


def main() -> int:
    if not os.path.isdir(REF):
        print("reference tree not present; nothing to run")
        return 0
    sys.path[:0] = [HERE, REF]
    import pytest

    # test_collision_avoidance_limit.py::test_contact_normal_jac_matches_mujoco needs real
    # MuJoCo contact/efc_J machinery -> deselected, documented.
    args = [os.path.join(REF, "tests"), "-q", "-p", "no:cacheprovider",
            "-k", "not test_contact_normal_jac_matches_mujoco",
            "--rootdir", "/tmp"] + sys.argv[1:]
    return pytest.main(args)


if __name__ == "__main__":
    sys.exit(main())
