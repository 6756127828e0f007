"""Stand-in for torch_sparse==0.6.8: the reference only uses `SparseTensor` in isinstance checks
(mp/cell_mp.py:166,199,263,332; data/complex.py:144,244,363,385,453); nothing ever constructs one."""


class SparseTensor(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("torch_sparse stand-in: SparseTensor is never constructed on the hot path")
