"""The experiment shim (`exp/`) and the data loader with the reference's signatures: the reference's own launch lines
parse and run. CPU tier: the six `cwn_b200.ops` entry points are substituted by torch restatements for one test at a time
(tests/cpu_ops_shim.py — test infrastructure, the product has no CPU path); the GPU tier repeats the runs for real."""
import os
import re

import pytest
import torch

import cpu_ops_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SCRIPTS = '/root/reference/exp/scripts'


def _flags_of(script_text):
    """The argument list of a `python -m exp.run_mol_exp \\ ...` launch line of the reference's scripts."""
    body = script_text.split('exp.run_mol_exp', 1)[1].replace('\\\n', ' ')
    return [tok.strip("'\"") for tok in body.split()]


ZINC_FLAGS = ['--device', '0', '--start_seed', '0', '--stop_seed', '9', '--exp_name', 'cwn-zinc', '--dataset', 'ZINC',
              '--train_eval_period', '20', '--epochs', '1000', '--batch_size', '128', '--drop_rate', '0.0',
              '--drop_position', 'lin2', '--emb_dim', '128', '--max_dim', '2', '--final_readout', 'sum', '--init_method',
              'sum', '--lr', '0.001', '--graph_norm', 'bn', '--model', 'embed_sparse_cin', '--nonlinearity', 'relu',
              '--num_layers', '4', '--readout', 'sum', '--max_ring_size', '18', '--task_type', 'regression',
              '--eval_metric', 'mae', '--minimize', '--lr_scheduler', 'ReduceLROnPlateau', '--use_coboundaries', 'True',
              '--use_edge_features', '--early_stop', '--lr_scheduler_patience', '20', '--dump_curves', '--preproc_jobs', '32']
MOLHIV_FLAGS = ['--device', '0', '--start_seed', '0', '--stop_seed', '9', '--exp_name', 'cwn-molhiv', '--dataset', 'MOLHIV',
                '--model', 'ogb_embed_sparse_cin', '--use_coboundaries', 'True', '--indrop_rate', '0.0', '--drop_rate',
                '0.5', '--graph_norm', 'bn', '--drop_position', 'lin2', '--nonlinearity', 'relu', '--readout', 'mean',
                '--final_readout', 'sum', '--lr', '0.0001', '--lr_scheduler', 'None', '--num_layers', '2', '--emb_dim',
                '64', '--batch_size', '128', '--epochs', '150', '--num_workers', '2', '--preproc_jobs', '32',
                '--task_type', 'bin_classification', '--eval_metric', 'ogbg-molhiv', '--max_dim', '2',
                '--max_ring_size', '6', '--init_method', 'sum', '--train_eval_period', '10', '--use_edge_features',
                '--dump_curves']


def _swap(flags, **over):
    out = list(flags)
    for k, v in over.items():
        if '--' + k in out:
            out[out.index('--' + k) + 1] = str(v)
        else:
            out += ['--' + k, str(v)]
    return out


@pytest.mark.skipif(not os.path.isdir(REF_SCRIPTS), reason='reference scripts not mounted')
def test_flag_lists_are_the_reference_scripts():
    assert _flags_of(open(os.path.join(REF_SCRIPTS, 'cwn-zinc.sh')).read()) == ZINC_FLAGS
    assert _flags_of(open(os.path.join(REF_SCRIPTS, 'cwn-molhiv.sh')).read()) == MOLHIV_FLAGS


def test_parser_has_the_reference_flags_and_defaults():
    from exp.parser import get_parser, validate_args
    args = get_parser().parse_args([])
    assert (args.seed, args.model, args.use_coboundaries, args.num_layers, args.emb_dim, args.batch_size) == \
        (43, 'sparse_cin', 'False', 5, 64, 32)
    assert tuple(args.readout_dims) == (0, 1, 2) and args.graph_norm == 'bn' and args.lr_scheduler == 'StepLR'
    if os.path.isfile('/root/reference/exp/parser.py'):  # every flag the reference declares exists here
        declared = set(re.findall(r"add_argument\('--(\w+)'", open('/root/reference/exp/parser.py').read()))
        assert declared <= set(vars(args))
    validate_args(get_parser().parse_args(ZINC_FLAGS))
    validate_args(get_parser().parse_args(MOLHIV_FLAGS))


def _run(flags, tmp_path, monkeypatch, cpu):
    from exp.parser import get_parser
    from exp.run_exp import main
    if cpu:
        cpu_ops_shim.install(monkeypatch)
        monkeypatch.setattr(torch.cuda, 'is_available', lambda: False)
    monkeypatch.setenv('CWN_SYNTH_SIZE', '96')
    return main(get_parser().parse_args(flags + ['--result_folder', str(tmp_path), '--seed', '0']))


def _dummym_flags():
    return ['--use_coboundaries', 'True', '--graph_norm', 'id', '--lr_scheduler', 'None', '--num_layers', '3', '--emb_dim',
            '8', '--batch_size', '3', '--epochs', '1', '--dataset', 'DUMMYM', '--max_ring_size', '5', '--exp_name',
            'dummym_test', '--readout_dims', '0', '2']


def test_run_exp_on_dummym(tmp_path, monkeypatch):
    """The reference's own `exp/test_run_exp.py`: sparse_cin on the hand-made molecular complexes, splits coincide."""
    curves = _run(_dummym_flags(), tmp_path, monkeypatch, cpu=True)
    assert curves['last_train'] == curves['last_val'] == curves['last_test']
    assert os.path.isfile(os.path.join(str(tmp_path), 'DUMMYM-dummym_test', 'seed-0', 'results.txt'))


@pytest.mark.parametrize('flags,dataset', [(ZINC_FLAGS, 'ZINC-SYNTH'), (MOLHIV_FLAGS, 'MOLHIV-SYNTH')])
def test_reference_launch_lines_build_and_train_the_models(flags, dataset, tmp_path, monkeypatch):
    """`embed_sparse_cin` / `ogb_embed_sparse_cin` built by `exp.run_exp.main` from the reference's OWN flag sets
    (exp/scripts/cwn-zinc.sh, cwn-molhiv.sh), one epoch on the synthetic stand-in of the dataset."""
    flags = _swap(flags, dataset=dataset, epochs=1, batch_size=16, emb_dim=16, num_workers=0)
    flags = [f for f in flags if f != '--dump_curves']
    curves = _run(flags, tmp_path, monkeypatch, cpu=True)
    assert len(curves['train_loss']) == 5 and all(v == v for v in curves['train_loss'])  # 76 train complexes / 16
    assert curves['val'][0] == curves['val'][0] and curves['last_test'] == curves['last_test']  # not NaN


def test_data_loader_signature_and_collation():
    from cwn_b200.data.complex import ComplexBatch
    from cwn_b200.data.data_loading import DataLoader, load_dataset
    ds = load_dataset('DUMMYM')
    loader = DataLoader(ds.get_split('train'), batch_size=3, shuffle=False, max_dim=ds.max_dim)
    batches = list(loader)
    assert all(isinstance(b, ComplexBatch) for b in batches) and sum(b.num_complexes for b in batches) == len(ds)
    ref = ComplexBatch.from_complex_list([ds[i] for i in range(3)], max_dim=2)
    assert torch.equal(batches[0].cochains[1].boundary_index, ref.cochains[1].boundary_index)
    with pytest.raises(NotImplementedError, match='download'):
        load_dataset('ZINC')


@pytest.mark.gpu
def test_run_exp_on_gpu_eager_and_graph_agree(tmp_path, monkeypatch):
    """The ZINC launch line on the B200 path: the CUDA-graph epoch loop (`--step_mode graph`: padded layout, one graph)
    reproduces the training-loss curve of the eager loop."""
    flags = _swap([f for f in ZINC_FLAGS if f != '--dump_curves'], dataset='ZINC-SYNTH', epochs=1, batch_size=16,
                  emb_dim=64, nonlinearity='elu')
    eager = _run(flags + ['--step_mode', 'eager'], tmp_path / 'a', monkeypatch, cpu=False)
    graph = _run(flags + ['--step_mode', 'graph'], tmp_path / 'b', monkeypatch, cpu=False)
    assert len(eager['train_loss']) == len(graph['train_loss']) == 5
    for a, b in zip(eager['train_loss'], graph['train_loss']):
        assert abs(a - b) <= 2e-3 * max(abs(a), 1.0), (eager['train_loss'], graph['train_loss'])


@pytest.mark.gpu
def test_device_data_loader_matches_python_collation():
    from cwn_b200.data import synthetic
    from cwn_b200.data.complex import ComplexBatch
    from cwn_b200.data.data_loading import DeviceDataLoader
    comps = synthetic.zinc_like_complexes(50, seed=3, ragged=True)
    loader = DeviceDataLoader(comps, batch_size=16, shuffle=False, max_dim=2, device='cuda:0')
    assert len(loader) == 4
    for i, batch in enumerate(loader):
        ref = ComplexBatch.from_complex_list(comps[16 * i:16 * (i + 1)], max_dim=2)
        for d in range(ref.dimension + 1):
            for key in ('x', 'upper_index', 'boundary_index', 'shared_coboundaries', 'batch'):
                u, v = getattr(ref.cochains[d], key), getattr(batch.cochains[d], key)
                assert (u is None) == (v is None)
                if u is not None:
                    assert torch.equal(u, v.cpu()), (i, d, key)
        assert torch.equal(ref.y, batch.y.cpu())
