class Box(object):
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class Dict(dict):
    def __init__(self, spaces=None, **kw):
        super().__init__(spaces or {}, **kw)
