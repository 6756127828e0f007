"""One training epoch / evaluation pass / metric evaluation with the reference's semantics (`exp/train_utils.py`):
same loss per task type, same skipping of one-sample batches, NaN-target masking, curve of per-batch losses.

`train(..., stepper=...)`: when a `cwn_b200.bucketed.BucketedStep` is supplied (regression on CUDA), a batch of complexes
is padded to the fixed-capacity layout and the whole step is ONE CUDA-graph replay; the eager loop below is the
fallback and the definition of the semantics."""
import logging

import numpy as np
import torch

from cwn_b200.data.complex import ComplexBatch

_LOSSES = {'classification': torch.nn.CrossEntropyLoss(), 'bin_classification': torch.nn.BCEWithLogitsLoss(),
           'regression': torch.nn.L1Loss(), 'mse_regression': torch.nn.MSELoss()}


def _loss_for(task_type, strict=True):
    if task_type not in _LOSSES:
        if strict:
            raise NotImplementedError(f'Training on task type {task_type} not yet supported.')
        return None
    return _LOSSES[task_type]


def _targets(loss_fn, batch, pred):
    if isinstance(loss_fn, torch.nn.CrossEntropyLoss):
        return batch.y.view(-1,)
    return batch.y.to(torch.float32).view(pred.shape)


def _progress(loader, desc):
    try:
        from tqdm import tqdm
        return tqdm(loader, desc=desc)
    except Exception:  # noqa: BLE001
        return loader


def train(model, device, loader, optimizer, task_type='classification', ignore_unlabeled=False, stepper=None):
    """One optimisation pass over the loader; returns the list of batch losses."""
    loss_fn = _loss_for(task_type)
    curve, skipped = [], 0
    model.train()
    for batch in _progress(loader, 'Training iteration'):
        if stepper is not None and isinstance(batch, (list, tuple)):  # raw complexes: padded layout + graph replay
            curve.append(float(stepper.step(batch).item()))
            continue
        batch = batch.to(device)
        if isinstance(batch, ComplexBatch):
            num_samples = batch.cochains[0].x.size(0)
            for dim in range(1, batch.dimension + 1):
                num_samples = min(num_samples, batch.cochains[dim].num_cells)
        else:
            num_samples = batch.x.size(0)
        if num_samples <= 1:  # BatchNorm cannot take a single sample
            skipped += 1
            if float(skipped) / len(loader) >= 0.25:
                logging.warning('Warning! 25% of the batches were skipped this epoch')
            continue
        if num_samples < 10:
            logging.warning('Warning! BatchNorm applied on a batch with only {} samples'.format(num_samples))
        optimizer.zero_grad()
        pred = model(batch)
        targets = _targets(loss_fn, batch, pred)
        mask = ~torch.isnan(targets)  # some ogbg-mol* targets are missing
        loss = loss_fn(pred[mask], targets[mask])
        loss.backward()
        optimizer.step()
        curve.append(loss.detach().cpu().item())
    return curve


def infer(model, device, loader):
    model.eval()
    preds = []
    for batch in _progress(loader, 'Inference iteration'):
        batch = batch.to(device)
        with torch.no_grad():
            preds.append(model(batch).detach().cpu())
    return torch.cat(preds, dim=0).numpy()


def eval(model, device, loader, evaluator, task_type):  # noqa: A001 — the reference's name
    loss_fn = _loss_for(task_type, strict=False)
    model.eval()
    y_true, y_pred, losses = [], [], []
    for batch in _progress(loader, 'Eval iteration'):
        if torch.get_default_dtype() == torch.float64:
            for dim in range(batch.dimension + 1):
                batch.cochains[dim].x = batch.cochains[dim].x.double()
        batch = batch.to(device)
        with torch.no_grad():
            pred = model(batch)
            if task_type != 'isomorphism':
                targets = _targets(loss_fn, batch, pred)
                y_true.append((batch.y if isinstance(loss_fn, torch.nn.CrossEntropyLoss) else batch.y.view(pred.shape)).detach().cpu())
                mask = ~torch.isnan(targets)
                losses.append(loss_fn(pred[mask], targets[mask]).detach().cpu().item())
        y_pred.append(pred.detach().cpu())
    y_true = torch.cat(y_true, dim=0).numpy() if y_true else None
    y_pred = torch.cat(y_pred, dim=0).numpy()
    mean_loss = float(np.mean(losses)) if losses else np.nan
    return evaluator.eval({'y_pred': y_pred, 'y_true': y_true}), mean_loss


class Evaluator(object):
    """accuracy / ap / mae / isomorphism as in the reference; `ogbg-mol*` metrics (the ogb package is absent here) are
    computed with scikit-learn the way `ogb.graphproppred.Evaluator` defines them: per-task over the labelled entries,
    averaged over tasks (rocauc / ap), or rmse."""

    _OGB = {'rocauc': ['molhiv', 'molbace', 'molbbbp', 'molclintox', 'molsider', 'moltox21', 'moltoxcast'],
            'ap': ['molpcba', 'molmuv'], 'rmse': ['molesol', 'molfreesolv', 'mollipo']}

    def __init__(self, metric, **kwargs):
        self.eps, self.p_norm = kwargs.get('eps', 0.01), kwargs.get('p', 2)
        if metric in ('isomorphism', 'accuracy', 'ap', 'mae'):
            self.eval_fn = getattr(self, '_' + metric)
        elif metric.startswith('ogbg-mol'):
            kind = [k for k, names in self._OGB.items() if metric[len('ogbg-'):] in names]
            if not kind:
                raise NotImplementedError(f'Metric {metric} is not yet supported.')
            self._kind, self.eval_fn = kind[0], self._ogb
        else:
            raise NotImplementedError('Metric {} is not yet supported.'.format(metric))

    def eval(self, input_dict):
        return self.eval_fn(input_dict)

    def _isomorphism(self, d):  # failure rate: pairs of embeddings closer than eps
        preds = torch.tensor(d['y_pred'], dtype=torch.float64)
        mm = torch.pdist(preds, p=self.p_norm)
        return (mm < self.eps).sum().item() / mm.shape[0]

    def _accuracy(self, d):
        return float((np.asarray(d['y_true']).reshape(-1) == np.argmax(d['y_pred'], axis=1)).mean())

    def _ap(self, d):
        from sklearn import metrics
        return metrics.average_precision_score(d['y_true'], d['y_pred'])

    def _mae(self, d):
        return float(np.abs(np.asarray(d['y_true'], dtype=np.float64) - np.asarray(d['y_pred'], dtype=np.float64)).mean())

    def _ogb(self, d):
        from sklearn import metrics
        y_true, y_pred = np.asarray(d['y_true'], dtype=np.float64), np.asarray(d['y_pred'], dtype=np.float64)
        if self._kind == 'rmse':
            return float(np.mean([np.sqrt(np.nanmean((y_true[:, i] - y_pred[:, i]) ** 2)) for i in range(y_true.shape[1])]))
        scores = []
        for i in range(y_true.shape[1]):
            labelled = ~np.isnan(y_true[:, i])
            t = y_true[labelled, i]
            if (t == 1).sum() > 0 and (t == 0).sum() > 0:
                fn = metrics.roc_auc_score if self._kind == 'rocauc' else metrics.average_precision_score
                scores.append(fn(t, y_pred[labelled, i]))
        if not scores:
            raise RuntimeError('No positively labeled data available. Cannot compute the metric.')
        return float(np.mean(scores))
