def __getattr__(name):
    def _noop(*a, **k):
        raise RuntimeError("matplotlib shim: plotting is out of scope (%s)" % name)
    return _noop
