"""Trajectory input (reference: src/io/loader.py)."""
from .loader import TrajectoryLoader  # noqa: F401
