class SinglePointCalculator:
    def __init__(self, atoms=None, **kw):
        self.kw = kw
