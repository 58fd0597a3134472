class Trajectory:
    def __init__(self, *a, **k):
        self.frames = []

    def write(self, atoms):
        self.frames.append(atoms)

    def close(self):
        pass


def read(*a, **k):
    raise NotImplementedError
