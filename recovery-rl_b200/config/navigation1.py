"""config/navigation1.py:104-160 of the reference: horizon 100, planning horizon 5 (body shared in config/base.py)."""
from .base import PointEnvConfigModule


class Navigation1ConfigModule(PointEnvConfigModule):
    ENV_NAME = "navigation1"


CONFIG_MODULE = Navigation1ConfigModule
