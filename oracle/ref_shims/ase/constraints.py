class FixAtoms:
    def __init__(self, mask=None, indices=None):
        self.mask = mask
