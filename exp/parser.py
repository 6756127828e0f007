"""The reference's experiment flags (`exp/parser.py:8-117`: same names, types and defaults, so its scripts parse
unchanged) and its dataset-dependent sanity checks (`:120-186`)."""
import argparse
import os
import time

ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (flag, type or 'flag' for store_true, default, help)
_FLAGS = [
    ('seed', int, 43, 'random seed'),
    ('start_seed', int, 0, 'first seed when evaluating on several seeds'),
    ('stop_seed', int, 9, 'last seed when evaluating on several seeds'),
    ('device', int, 0, 'which gpu to use'),
    ('model', str, 'sparse_cin', 'model name'),
    ('use_coboundaries', str, 'False', "coboundary features in the up-messages ('True' / 'False', a string as in the reference)"),
    ('include_down_adj', 'flag', False, 'use lower adjacencies (CIN++)'),
    ('indrop_rate', float, 0.0, 'input dropout rate of the molecular models'),
    ('drop_rate', float, 0.0, 'dropout rate'),
    ('drop_position', str, 'lin2', 'where the final dropout is applied'),
    ('nonlinearity', str, 'relu', 'activation function'),
    ('readout', str, 'sum', 'readout function'),
    ('final_readout', str, 'sum', 'final readout function'),
    ('jump_mode', str, None, 'jumping-knowledge mode'),
    ('lr', float, 0.001, 'learning rate'),
    ('lr_scheduler', str, 'StepLR', 'learning-rate scheduler'),
    ('lr_scheduler_decay_steps', int, 50, 'epochs between lr decays'),
    ('lr_scheduler_decay_rate', float, 0.5, 'strength of the lr decay'),
    ('lr_scheduler_patience', float, 10, 'patience of ReduceLROnPlateau'),
    ('lr_scheduler_min', float, 0.00001, 'minimum lr of ReduceLROnPlateau'),
    ('num_layers', int, 5, 'number of message-passing layers'),
    ('emb_dim', int, 64, 'hidden width'),
    ('batch_size', int, 32, 'batch size'),
    ('epochs', int, 100, 'number of epochs'),
    ('num_workers', int, 0, 'data-loader workers'),
    ('dataset', str, 'PROTEINS', 'dataset name'),
    ('task_type', str, 'classification', '(bin_)classification, (mse_)regression or isomorphism'),
    ('eval_metric', str, 'accuracy', 'evaluation metric'),
    ('iso_eps', int, 0.01, 'threshold of (non-)isomorphism'),
    ('minimize', 'flag', False, 'the evaluation metric is to be minimised'),
    ('max_dim', int, '2', 'maximum cellular dimension'),
    ('max_ring_size', int, None, 'maximum ring size to look for'),
    ('result_folder', str, os.path.join(ROOT_DIR, 'exp', 'results'), 'where results are written'),
    ('exp_name', str, None, 'experiment name (default: unix timestamp)'),
    ('dump_curves', 'flag', False, 'dump the training curves to disk'),
    ('untrained', 'flag', False, 'skip training'),
    ('fold', int, None, 'fold index of k-fold cross-validation'),
    ('folds', int, None, 'number of folds'),
    ('init_method', str, 'sum', 'how features of higher cells are initialised (sum, mean)'),
    ('train_eval_period', int, 10, 'how often to evaluate on train'),
    ('tune', 'flag', False, 'use the tuning indexes'),
    ('flow_points', int, 400, 'points of the flow experiment'),
    ('flow_classes', int, 3, 'classes of the flow experiment'),
    ('train_orient', str, 'default', 'orientation of the training complexes'),
    ('test_orient', str, 'default', 'orientation of the testing complexes'),
    ('fully_orient_invar', 'flag', False, 'apply torch.abs from the first layer'),
    ('use_edge_features', 'flag', False, 'use edge features of molecular graphs'),
    ('simple_features', 'flag', False, 'subset of the ogb-mol* features'),
    ('early_stop', 'flag', False, 'stop when the minimum lr is reached'),
    ('paraid', int, 0, 'model id'),
    ('preproc_jobs', int, 2, 'jobs of the dataset preprocessing'),
]


def get_parser():
    parser = argparse.ArgumentParser(description='CWN experiment (cwn_b200).')
    for name, kind, default, text in _FLAGS:
        if kind == 'flag':
            parser.add_argument('--' + name, action='store_true', help=text)
        else:
            if name == 'exp_name':
                default = str(time.time())
            parser.add_argument('--' + name, type=kind, default=default, help=text)
    parser.add_argument('--readout_dims', type=int, nargs='+', default=(0, 1, 2), help='dimensions of the final readout')
    parser.add_argument('--graph_norm', type=str, default='bn', choices=['bn', 'ln', 'id'], help='normalisation layer')
    # cwn_b200 extension: how a training epoch is executed (results are those of the eager loop)
    parser.add_argument('--step_mode', type=str, default='auto', choices=['auto', 'eager', 'graph'],
                        help="'graph': batches are padded to one layout and replayed through one CUDA graph "
                             "(cwn_b200.bucketed); 'auto' uses it for regression on CUDA when the model qualifies")
    return parser


_OGB = ['MOLHIV', 'MOLPCBA', 'MOLTOX21', 'MOLTOXCAST', 'MOLMUV', 'MOLBACE', 'MOLBBBP', 'MOLCLINTOX', 'MOLSIDER', 'MOLESOL',
        'MOLFREESOLV', 'MOLLIPO']


def validate_args(args):
    """Dataset-dependent sanity checks (the reference's rules for the datasets this shim knows)."""
    if args.dataset.startswith('ZINC'):
        assert args.model.startswith('embed')
        if args.model == 'embed_cin++':
            assert args.include_down_adj is True
        assert args.task_type == 'regression' and args.minimize and args.eval_metric == 'mae'
        assert args.lr_scheduler == 'ReduceLROnPlateau' and not args.simple_features
    elif args.dataset.split('-')[0] in _OGB:
        assert args.model in ('ogb_embed_sparse_cin', 'ogb_embed_cin++')
        if args.model == 'ogb_embed_cin++':
            assert args.include_down_adj is True
        assert args.eval_metric == 'ogbg-' + args.dataset.split('-')[0].lower() and args.jump_mode is None
        if args.dataset.split('-')[0] in ('MOLESOL', 'MOLFREESOLV', 'MOLLIPO'):
            assert args.task_type == 'mse_regression' and args.minimize
        else:
            assert args.task_type == 'bin_classification' and not args.minimize
