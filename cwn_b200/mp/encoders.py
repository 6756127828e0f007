"""Atom / bond encoders of the ogbg-mol* models: the sum over integer feature columns of one
`nn.Embedding(vocab_i, emb_dim)` each, xavier-uniform initialised (ogb==1.3.1 `mol_encoder.py`, which the
reference imports at `mp/molec_models.py:7`, `mp/layers.py:10`; ogb itself is not vendored in the reference).
Vocabulary sizes default to ogb's `get_atom_feature_dims()` / `get_bond_feature_dims()` and can be overridden."""
import os

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
        if (x.is_cuda and x.dim() == 2 and x.size(1) == len(tables) and x.size(0) > 0 and self.fuse_lookup
                and all(type(t) is torch.nn.Embedding and t.weight.dtype == torch.float32 and t.padding_idx is None
                        and t.max_norm is None and not t.sparse and not t.scale_grad_by_freq for t in tables)):
            # ONE row gather over the concatenated tables instead of a lookup + add per feature column, and ONE
            # segmented reduction (CSR plan of the flat ids, cwn_b200.ops) instead of a sort-based embedding backward per
            # column: 9 + 3 columns were ~30 launches forward and ~100 backward per step of the ogbg-mol* models
            return self._lookup_concat(x)
        out = 0
        for i in range(x.shape[1]):
            out = out + tables[i](x[:, i])
        return out

    def _lookup_concat(self, x):
        from cwn_b200 import ops
        tables = getattr(self, self._list_name)
        flat = (x + self._offsets(x.device)).reshape(-1)
        rows = ops.gather_rows(torch.cat([t.weight for t in tables]), flat)
        return rows.view(x.size(0), len(tables), -1).sum(dim=1)

    fuse_lookup = os.environ.get('CWN_B200_FUSE_OGB_LOOKUP', '1') != '0'  # A/B switch

    def _offsets(self, device):
        """[1, C] first row of every column's table inside the concatenation (cached per device)."""
        cache = self.__dict__.setdefault('_offs', {})
        t = cache.get(device)
        if t is None:
            sizes = [tab.num_embeddings for tab in getattr(self, self._list_name)]
            starts = [sum(sizes[:i]) for i in range(len(sizes))]
            t = cache[device] = torch.tensor([starts], dtype=torch.long, device=device)
        return t


class AtomEncoder(_ColumnEmbeddingSum):
    def __init__(self, emb_dim, feature_dims=ATOM_FEATURE_DIMS):
        super(AtomEncoder, self).__init__(emb_dim, feature_dims, 'atom_embedding_list')


class BondEncoder(_ColumnEmbeddingSum):
    def __init__(self, emb_dim, feature_dims=BOND_FEATURE_DIMS):
        super(BondEncoder, self).__init__(emb_dim, feature_dims, 'bond_embedding_list')
