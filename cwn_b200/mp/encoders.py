"""Atom / bond encoders of the ogbg-mol* models: the sum over integer feature columns of one
`nn.Embedding(vocab_i, emb_dim)` each, xavier-uniform initialised (ogb==1.3.1 `mol_encoder.py`, which the
reference imports at `mp/molec_models.py:7`, `mp/layers.py:10`; ogb itself is not vendored in the reference).
Vocabulary sizes default to ogb's `get_atom_feature_dims()` / `get_bond_feature_dims()` and can be overridden."""
import torch

ATOM_FEATURE_DIMS = (119, 4, 12, 12, 10, 6, 6, 2, 2)
BOND_FEATURE_DIMS = (5, 6, 2)


class _ColumnEmbeddingSum(torch.nn.Module):
    def __init__(self, emb_dim, dims, list_name):
        super(_ColumnEmbeddingSum, self).__init__()
        self._list_name = list_name
        tables = torch.nn.ModuleList()
        for vocab in dims:
            emb = torch.nn.Embedding(vocab, emb_dim)
            torch.nn.init.xavier_uniform_(emb.weight.data)
            tables.append(emb)
        setattr(self, list_name, tables)

    def forward(self, x):
        tables = getattr(self, self._list_name)
        out = 0
        for i in range(x.shape[1]):
            out = out + tables[i](x[:, i])
        return out


class AtomEncoder(_ColumnEmbeddingSum):
    def __init__(self, emb_dim, feature_dims=ATOM_FEATURE_DIMS):
        super(AtomEncoder, self).__init__(emb_dim, feature_dims, 'atom_embedding_list')


class BondEncoder(_ColumnEmbeddingSum):
    def __init__(self, emb_dim, feature_dims=BOND_FEATURE_DIMS):
        super(BondEncoder, self).__init__(emb_dim, feature_dims, 'bond_embedding_list')
