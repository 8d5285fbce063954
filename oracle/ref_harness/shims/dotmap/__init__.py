class DotMap(dict):
    """Auto-vivifying attribute dict (subset of dotmap.DotMap used by the reference)."""

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
