class ImageSequenceClip(object):
    def __init__(self, *a, **k):
        raise RuntimeError("moviepy shim: gif output is out of scope")
