"""config/maze.py:104-165 of the reference: horizon 150, planning horizon 15 (body shared in config/base.py)."""
from .base import PointEnvConfigModule


class MazeConfigModule(PointEnvConfigModule):
    ENV_NAME = "maze"
    TASK_HORIZON = 150
    PLAN_HOR = 15


CONFIG_MODULE = MazeConfigModule
