"""`main(args)`: the common training / evaluation script of the reference (`exp/run_exp.py:20-480`) over the cwn_b200
models: same model names and constructor wiring (`:112-328`), Adam + the same schedulers (`:343-358`), the same epoch
loop, curves dictionary and result files (`:360-476`)."""
import copy
import os
import pickle
import random

import numpy as np
import torch
import torch.optim as optim

from cwn_b200.data.data_loading import DataLoader, load_dataset
from cwn_b200.mp import models as M
from cwn_b200.mp import molec_models as MM
from exp.parser import get_parser, validate_args
from exp.train_utils import Evaluator, eval, train  # noqa: A004


def build_model(args, dataset, device):
    """The model `--model` names, wired from the flags exactly as the reference does."""
    cob = args.use_coboundaries.lower() == 'true'
    readout_dims = tuple(sorted(args.readout_dims))
    plain = dict(dropout_rate=args.drop_rate, max_dim=dataset.max_dim, jump_mode=args.jump_mode,
                 nonlinearity=args.nonlinearity, readout=args.readout)
    sparse = dict(plain, final_readout=args.final_readout, apply_dropout_before=args.drop_position, use_coboundaries=cob,
                  graph_norm=args.graph_norm, readout_dims=readout_dims)
    embed = dict(sparse, embed_edge=args.use_edge_features)
    ogb = dict(embed, indropout_rate=args.indrop_rate)
    name = args.model
    if name == 'cin':
        model = M.CIN0(dataset.num_features_in_dim(0), dataset.num_classes, args.num_layers, args.emb_dim, **plain)
    elif name == 'sparse_cin':
        model = M.SparseCIN(dataset.num_features_in_dim(0), dataset.num_classes, args.num_layers, args.emb_dim, **sparse)
    elif name == 'cin++':
        model = M.CINpp(dataset.num_features_in_dim(0), dataset.num_classes, args.num_layers, args.emb_dim, **sparse)
    elif name == 'edge_orient':
        model = M.EdgeOrient(1, dataset.num_classes, args.num_layers, args.emb_dim, dropout_rate=args.drop_rate,
                             readout=args.readout, nonlinearity=args.nonlinearity, fully_invar=args.fully_orient_invar)
    elif name == 'edge_mpnn':
        model = M.EdgeMPNN(1, dataset.num_classes, args.num_layers, args.emb_dim, dropout_rate=args.drop_rate,
                           readout=args.readout, nonlinearity=args.nonlinearity, fully_invar=args.fully_orient_invar)
    elif name == 'embed_sparse_cin':
        model = MM.EmbedSparseCIN(dataset.num_node_type, dataset.num_edge_type, dataset.num_classes, args.num_layers,
                                  args.emb_dim, **embed)
    elif name == 'embed_cin++':
        model = MM.EmbedCINpp(dataset.num_node_type, dataset.num_edge_type, dataset.num_classes, args.num_layers,
                              args.emb_dim, **embed)
    elif name == 'ogb_embed_sparse_cin':
        model = MM.OGBEmbedSparseCIN(dataset.num_tasks, args.num_layers, args.emb_dim, **ogb)
    elif name == 'ogb_embed_cin++':
        model = MM.OGBEmbedCINpp(dataset.num_tasks, args.num_layers, args.emb_dim, **ogb)
    else:
        raise ValueError(f'Invalid model type {name} (graph baselines and ring-experiment models are outside the rebuilt path).')
    return model.to(device)


def _stepper(args, model, optimizer, dataset, device):
    """A `BucketedStep` (padded layout, one CUDA graph per epoch loop) when the configuration qualifies, else None."""
    if args.step_mode == 'eager' or device.type != 'cuda' or args.task_type != 'regression':
        return None, None
    try:
        from cwn_b200.bucketed import BucketedStep, Capacity, masked_l1
        from cwn_b200.dist import FlatGradBucket
        train_set = dataset.get_split('train')
        complexes = [train_set[i] for i in range(len(train_set))]
        cap = Capacity.from_dataset(complexes, args.batch_size)
        step = BucketedStep(model, masked_l1, FlatGradBucket(model), optimizer, capacity=cap)
        step.capture(complexes[:args.batch_size])
        loader = torch.utils.data.DataLoader(train_set, batch_size=args.batch_size, shuffle=True, collate_fn=list,
                                             num_workers=0)
        return step, loader
    except Exception as exc:  # noqa: BLE001
        if args.step_mode == 'graph':
            raise
        print(f'step_mode auto: CUDA-graph stepping unavailable ({type(exc).__name__}: {exc}); eager loop')
        return None, None


def main(args):
    device = torch.device('cuda:' + str(args.device)) if torch.cuda.is_available() else torch.device('cpu')
    print('Using device', str(device), '| fold', args.fold, '| seed', args.seed)
    print(args)
    for seed_fn in (torch.manual_seed, torch.cuda.manual_seed, torch.cuda.manual_seed_all, np.random.seed, random.seed):
        seed_fn(args.seed)
    if args.task_type == 'isomorphism':
        assert args.dataset.startswith('sr')
        torch.set_default_dtype(torch.float64)
    result_folder = os.path.join(args.result_folder, f'{args.dataset}-{args.exp_name}', f'seed-{args.seed}')
    if args.fold is not None:
        result_folder = os.path.join(result_folder, f'fold-{args.fold}')
    os.makedirs(result_folder, exist_ok=True)

    dataset = load_dataset(args.dataset, max_dim=args.max_dim, fold=args.fold, init_method=args.init_method,
                           emb_dim=args.emb_dim, flow_points=args.flow_points, flow_classes=args.flow_classes,
                           max_ring_size=args.max_ring_size, use_edge_features=args.use_edge_features,
                           include_down_adj=args.include_down_adj, simple_features=args.simple_features,
                           n_jobs=args.preproc_jobs, train_orient=args.train_orient, test_orient=args.test_orient)
    if args.dataset.startswith('ZINC'):  # vocabulary sizes of data/datasets/zinc.py:29-30
        dataset.num_node_type, dataset.num_edge_type = 28, 4
    split_idx = dataset.get_tune_idx_split() if args.tune else dataset.get_idx_split()
    loader_of = lambda split, shuffle: DataLoader(dataset.get_split(split), batch_size=args.batch_size, shuffle=shuffle,  # noqa: E731
                                                  num_workers=args.num_workers, max_dim=dataset.max_dim)
    train_loader, valid_loader = loader_of('train', True), loader_of('valid', False)
    test_loader = loader_of('test', False) if split_idx.get('test') is not None else None
    evaluator = Evaluator(args.eval_metric, eps=args.iso_eps)

    model = build_model(args, dataset, device)
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    print(f'Trainable params: {trainable} | total: {sum(p.numel() for p in model.parameters())}')
    optimizer = optim.Adam(model.parameters(), lr=args.lr, **({'capturable': True} if device.type == 'cuda' else {}))
    if args.lr_scheduler == 'ReduceLROnPlateau':
        scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode='min' if args.minimize else 'max',
                                                               factor=args.lr_scheduler_decay_rate,
                                                               patience=args.lr_scheduler_patience)
    elif args.lr_scheduler == 'StepLR':
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, args.lr_scheduler_decay_steps,
                                                    gamma=args.lr_scheduler_decay_rate)
    elif args.lr_scheduler == 'None':
        scheduler = None
    else:
        raise NotImplementedError(f'Scheduler {args.lr_scheduler} is not currently supported.')
    stepper, raw_loader = (None, None) if args.untrained else _stepper(args, model, optimizer, dataset, device)

    best_val_epoch, train_perf = 0, np.nan
    valid_curve, test_curve, train_curve, train_loss_curve = [], [], [], []
    if not args.untrained:
        for epoch in range(1, args.epochs + 1):
            print(f'=====Epoch {epoch}')
            if stepper is not None:
                losses = train(model, device, raw_loader, optimizer, args.task_type, stepper=stepper)
            else:
                losses = train(model, device, train_loader, optimizer, args.task_type)
            train_loss_curve += losses
            epoch_train_loss = float(np.mean(losses))
            if epoch == 1 or epoch % args.train_eval_period == 0:
                train_perf, _ = eval(model, device, train_loader, evaluator, args.task_type)
            train_curve.append(train_perf)
            valid_perf, val_loss = eval(model, device, valid_loader, evaluator, args.task_type)
            valid_curve.append(valid_perf)
            test_perf, test_loss = eval(model, device, test_loader, evaluator, args.task_type) \
                if test_loader is not None else (np.nan, np.nan)
            test_curve.append(test_perf)
            print(f'Train: {train_perf:.3f} | Validation: {valid_perf:.3f} | Test: {test_perf:.3f} | Train Loss '
                  f'{epoch_train_loss:.3f} | Val Loss {val_loss:.3f} | Test Loss {test_loss:.3f}')
            if scheduler is not None:
                if args.lr_scheduler == 'ReduceLROnPlateau':
                    scheduler.step(valid_perf)
                    if args.early_stop and optimizer.param_groups[0]['lr'] < args.lr_scheduler_min:
                        print('\n!! The minimum learning rate has been reached.')
                        break
                else:
                    scheduler.step()
        best_val_epoch = int(np.argmin(np.array(valid_curve)) if args.minimize else np.argmax(np.array(valid_curve)))
    else:
        for curve in (train_loss_curve, train_curve, valid_curve, test_curve):
            curve.append(np.nan)

    print('Final Evaluation...')
    final_train = final_val = final_test = np.nan
    if not args.dataset.startswith('sr'):
        final_train, _ = eval(model, device, train_loader, evaluator, args.task_type)
        final_val, _ = eval(model, device, valid_loader, evaluator, args.task_type)
    if test_loader is not None:
        final_test, _ = eval(model, device, test_loader, evaluator, args.task_type)
    curves = {'train_loss': train_loss_curve, 'train': train_curve, 'val': valid_curve, 'test': test_curve,
              'last_val': final_val, 'last_test': final_test, 'last_train': final_train, 'best': best_val_epoch}
    msg = (f'========== Result ============\nDataset:        {args.dataset}\n------------ Best epoch -----------\n'
           f'Train:          {train_curve[best_val_epoch]}\nValidation:     {valid_curve[best_val_epoch]}\n'
           f'Test:           {test_curve[best_val_epoch]}\nBest epoch:     {best_val_epoch}\n'
           f'------------ Last epoch -----------\nTrain:          {final_train}\nValidation:     {final_val}\n'
           f'Test:           {final_test}\n-------------------------------\n\n')
    print(msg)
    with open(os.path.join(result_folder, 'results.txt'), 'w') as handle:
        handle.write(msg + str(args))
    if args.dump_curves:
        with open(os.path.join(result_folder, 'curves.pkl'), 'wb') as handle:
            pickle.dump(curves, handle)
    return curves


if __name__ == '__main__':
    _args = get_parser().parse_args()
    validate_args(_args)
    main(copy.copy(_args))
