class Line(object):
    pass
