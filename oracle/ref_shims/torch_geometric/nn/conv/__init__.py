from torch_geometric.nn import GINConv, GINEConv  # noqa: F401


class MessagePassing(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("stand-in: PyG MessagePassing is only used by reference tests as a comparator")
