"""Auto-vivifying attribute dict: the subset of `dotmap.DotMap` the reference's config tree uses
(config/default.py:15-119, recovery_rl/utils.py:84-88).  dotmap itself is not a dependency of this repo."""


class DotMap(dict):
    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        if key not in self:
            self[key] = DotMap()
        return self[key]

    def __setattr__(self, key, value):
        self[key] = value

    def pprint(self):
        print(dict(self))
