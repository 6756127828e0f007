"""Data loading with the reference's interface (`data/data_loading.py:44-110, 113-218`): `Collater`, `DataLoader(dataset,
batch_size, shuffle, follow_batch, max_dim)` and `load_dataset(name, ...)`, plus `DeviceDataLoader`, the GPU-side
variant (dataset resident in HBM, batches collated by one kernel).

Datasets. Download and lifting pipelines (ZINC / OGB / TU / SR files, graph-tool, gudhi) are out of scope and
there is no network here; `load_dataset` knows
    DUMMY, DUMMYM            the reference's hand-made complexes (`data/datasets/dummy.py`),
    ZINC-SYNTH, MOLHIV-SYNTH seeded synthetic ring-lifted molecules with the ZINC / ogbg-molhiv feature conventions
                             (`cwn_b200.data.synthetic`), ragged, split 80 / 10 / 10,
and raises for every other name with the reason.
"""
from collections.abc import Mapping, Sequence

import torch
from torch.utils.data.dataloader import default_collate

from cwn_b200.data.complex import Cochain, CochainBatch, Complex, ComplexBatch


class Collater(object):
    """Turns a python list of samples into the right storage format (`data/data_loading.py:44-82`)."""

    def __init__(self, follow_batch, max_dim=2):
        self.follow_batch, self.max_dim = follow_batch, max_dim

    def collate(self, batch):
        first = batch[0]
        if isinstance(first, Cochain):
            return CochainBatch.from_cochain_list(batch, self.follow_batch)
        if isinstance(first, Complex):
            return ComplexBatch.from_complex_list(batch, self.follow_batch, max_dim=self.max_dim)
        if isinstance(first, torch.Tensor):
            return default_collate(batch)
        if isinstance(first, float):
            return torch.tensor(batch, dtype=torch.float)
        if isinstance(first, int):
            return torch.tensor(batch)
        if isinstance(first, str):
            return batch
        if isinstance(first, Mapping):
            return {key: self.collate([sample[key] for sample in batch]) for key in first}
        if isinstance(first, tuple) and hasattr(first, '_fields'):
            return type(first)(*(self.collate(group) for group in zip(*batch)))
        if isinstance(first, Sequence):
            return [self.collate(group) for group in zip(*batch)]
        raise TypeError(f'DataLoader found invalid type: {type(first)}')

    def __call__(self, batch):
        return self.collate(batch)


class DataLoader(torch.utils.data.DataLoader):
    """Mini-batches of cochain complexes (`data/data_loading.py:85-110`; same signature)."""

    def __init__(self, dataset, batch_size=1, shuffle=False, follow_batch=(), max_dim=2, **kwargs):
        kwargs.pop('collate_fn', None)
        self.follow_batch = follow_batch
        super(DataLoader, self).__init__(dataset, batch_size, shuffle, collate_fn=Collater(follow_batch, max_dim),
                                         **kwargs)


class DeviceDataLoader(object):
    """The same iteration protocol over a `PackedComplexDataset` (dataset resident in HBM): every batch is collated ON THE
    GPU by one kernel; per batch the host does ~0.2 ms of numpy (segment sizes, prefix sums), one small H2D copy and one
    launch.

        loader = DeviceDataLoader(complex_list, batch_size=128, shuffle=True, max_dim=2, device='cuda')
        for batch in loader: ...          # device-resident, packed ComplexBatch
    """

    def __init__(self, dataset, batch_size=1, shuffle=False, max_dim=2, device='cuda', seed=0, drop_last=False):
        from cwn_b200.data.packed import PackedComplexDataset
        self.packed = dataset if isinstance(dataset, PackedComplexDataset) \
            else PackedComplexDataset(list(dataset), max_dim=max_dim, device=device)
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), shuffle, drop_last
        self._gen = torch.Generator().manual_seed(seed)

    def __len__(self):
        n = len(self.packed)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def id_batches(self):
        n = len(self.packed)
        order = torch.randperm(n, generator=self._gen).tolist() if self.shuffle else list(range(n))
        out = [order[i:i + self.batch_size] for i in range(0, n, self.batch_size)]
        if self.drop_last and out and len(out[-1]) < self.batch_size:
            out.pop()
        return out

    def __iter__(self):
        return prefetched(self.packed, self.id_batches())


def prefetched(packed, id_batches, out=None):
    """Generator of collated batches. The host half of batch i + 1 (`packed.prepare`: segment sizes, prefix sums, the
    table — ~0.2 ms of numpy) is done right after the collate kernel of batch i has been QUEUED, i.e. while the GPU is
    still copying batch i. (A worker thread was tried and measured slower: its numpy work holds the GIL in pieces and
    delays the launching thread. To hide the host half behind the training step itself, call `prepare` for the next
    batch between launching the step and reading its loss, as bench.py's e2e_gpu_collation loop does.)"""
    id_batches = list(id_batches)
    prep = packed.prepare(id_batches[0]) if id_batches else None
    for i in range(len(id_batches)):
        batch = packed.launch(prep, out=out)
        if i + 1 < len(id_batches):
            prep = packed.prepare(id_batches[i + 1])
        yield batch


# --------------------------------------------------------------------------------------------------------- datasets
class ListComplexDataset(torch.utils.data.Dataset):
    """An in-memory list of complexes with the attributes `exp/run_exp.py` reads from a `ComplexDataset`
    (`data/datasets/dataset.py:77-110, 364`): `max_dim`, `num_classes`, `num_tasks`, `num_features_in_dim`,
    `get_idx_split`, `get_tune_idx_split`, `get_split`."""

    def __init__(self, name, complexes, max_dim, num_classes, train_ids, val_ids, test_ids=None, num_tasks=1):
        self.name, self.complexes, self.max_dim, self.num_classes, self.num_tasks = name, complexes, max_dim, num_classes, num_tasks
        self.train_ids, self.val_ids, self.test_ids = train_ids, val_ids, test_ids

    def __len__(self):
        return len(self.complexes)

    def __getitem__(self, idx):
        if isinstance(idx, (list, tuple)):
            return ListComplexDataset(self.name, [self.complexes[i] for i in idx], self.max_dim, self.num_classes,
                                      list(range(len(idx))), [], None, self.num_tasks)
        return self.complexes[int(idx)]

    def num_features_in_dim(self, dim):
        if dim > self.max_dim:
            raise ValueError(f'`dim` {dim} larger than max allowed dimension {self.max_dim}.')
        for comp in self.complexes:
            c = comp.cochains.get(dim)
            if c is not None and c.x is not None:
                return c.x.size(-1)
        return 0

    def get_idx_split(self):
        split = {'train': self.train_ids, 'valid': self.val_ids}
        if self.test_ids is not None:
            split['test'] = self.test_ids
        return split

    def get_tune_idx_split(self):
        return self.get_idx_split()

    def get_split(self, split):
        if split not in ('train', 'valid', 'test'):
            raise ValueError(f'Unknown split {split}.')
        ids = self.get_idx_split().get(split)
        if ids is None:
            raise AssertionError('No split information found.')
        return self[list(ids)]


def _synthetic(name, size, seed, ogb, use_edge_features, include_down_adj):
    from cwn_b200.data import synthetic
    comps = synthetic.zinc_like_complexes(size, seed=seed, ragged=True, ring_count_range=(1, 5), pendant_range=(4, 15),
                                          ogb_features=ogb, edge_features=use_edge_features,
                                          include_down_adj=include_down_adj)
    if ogb:
        for i, c in enumerate(comps):  # binary task, floats as in ogbg-molhiv
            c.y = torch.tensor([[float(c.cochains[0].num_cells % 2)]])
    n_train, n_val = int(0.8 * size), int(0.1 * size)
    ids = list(range(size))
    return ListComplexDataset(name, comps, 2, 1 if not ogb else 2, ids[:n_train], ids[n_train:n_train + n_val],
                              ids[n_train + n_val:], num_tasks=1)


def load_dataset(name, root=None, max_dim=2, fold=0, init_method='sum', n_jobs=2, **kwargs):
    """`data/data_loading.py:113-218` for the datasets available without network or lifting libraries."""
    import os
    from cwn_b200.data import dummy_complexes
    size = int(os.environ.get('CWN_SYNTH_SIZE', '1024'))
    if name == 'DUMMY':
        comps = dummy_complexes.get_testing_complex_list()
        for i, c in enumerate(comps):
            c.y = torch.LongTensor([i % 2])
        ids = list(range(len(comps)))
        return ListComplexDataset(name, comps, 3, 2, ids, ids, ids)
    if name == 'DUMMYM':
        comps = dummy_complexes.get_mol_testing_complex_list()
        for i, c in enumerate(comps):
            c.y = torch.LongTensor([i % 2])
        ids = list(range(len(comps)))
        return ListComplexDataset(name, comps, 2, 2, ids, ids, ids)
    if name == 'ZINC-SYNTH':
        return _synthetic(name, size, 0, False, kwargs.get('use_edge_features', False), kwargs.get('include_down_adj', False))
    if name == 'MOLHIV-SYNTH':
        return _synthetic(name, size, 1, True, kwargs.get('use_edge_features', False), kwargs.get('include_down_adj', False))
    raise NotImplementedError(
        f"cwn_b200: dataset '{name}' needs the reference's download + lifting pipeline (graph-tool / gudhi / ogb / network), "
        f"which is outside the hot path rebuilt here; available: DUMMY, DUMMYM, ZINC-SYNTH, MOLHIV-SYNTH")
