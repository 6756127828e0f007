"""AtomEncoder / BondEncoder: sum over integer feature columns of nn.Embedding(dim_i, emb_dim), xavier-uniform init."""
import torch
from ogb.utils.features import get_atom_feature_dims, get_bond_feature_dims


class _SumOfEmbeddings(torch.nn.Module):
    _dims = ()
    _list_name = "embedding_list"

    def __init__(self, emb_dim):
        super().__init__()
        embs = torch.nn.ModuleList()
        for dim in self._dims:
            emb = torch.nn.Embedding(dim, emb_dim)
            torch.nn.init.xavier_uniform_(emb.weight.data)
            embs.append(emb)
        setattr(self, self._list_name, embs)

    def forward(self, x):
        embs = getattr(self, self._list_name)
        out = 0
        for i in range(x.shape[1]):
            out = out + embs[i](x[:, i])
        return out


class AtomEncoder(_SumOfEmbeddings):
    _dims = tuple(get_atom_feature_dims())
    _list_name = "atom_embedding_list"


class BondEncoder(_SumOfEmbeddings):
    _dims = tuple(get_bond_feature_dims())
    _list_name = "bond_embedding_list"
