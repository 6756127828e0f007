class Data(object):
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return self.x.size(0)


class Batch(Data):
    pass
