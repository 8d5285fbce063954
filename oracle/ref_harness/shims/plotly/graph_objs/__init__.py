class Scatter(object):
    pass
from . import scatter  # noqa: F401,E402
