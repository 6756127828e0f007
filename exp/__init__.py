"""Experiment entry points of the reference (`exp/run_mol_exp.py`, `exp/run_exp.py`, `exp/parser.py`,
`exp/train_utils.py`) over the cwn_b200 models and data API, so that the reference's launch lines
(`python -m exp.run_mol_exp --dataset ... --model embed_sparse_cin ...`, `exp/scripts/*.sh`) resolve to the B200 path.
Dataset download / lifting pipelines are out of scope: the registered datasets are the reference's hand-made
DUMMY / DUMMYM sets and seeded synthetic ZINC- / molhiv-shaped sets; see `cwn_b200/data/data_loading.py`."""
