"""config/navigation2.py:104-160 of the reference: horizon 100, planning horizon 5 (body shared in config/base.py)."""
from .base import PointEnvConfigModule


class Navigation2ConfigModule(PointEnvConfigModule):
    ENV_NAME = "navigation2"


CONFIG_MODULE = Navigation2ConfigModule
