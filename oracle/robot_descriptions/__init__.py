"""ORACLE / TEST INFRASTRUCTURE -- offline stand-in for `robot_descriptions` (see loaders/mujoco.py)."""
