class Rectangle(object):
    def __init__(self, *a, **k):
        pass
