"""`python -m exp.run_mol_exp <flags>`: one run per seed (per fold for CSL), then mean / std of the metric at the best
validation epoch and at the last epoch (`exp/run_mol_exp.py` of the reference)."""
import copy
import os
import subprocess
import sys
from itertools import product

import numpy as np

from exp.parser import get_parser
from exp.run_exp import main


def _sha():
    try:
        return subprocess.check_output(['git', 'describe', '--always'], stderr=subprocess.DEVNULL).strip().decode()
    except Exception:  # noqa: BLE001
        return 'unknown'


def exp_main(passed_args):
    parser = get_parser()
    args = parser.parse_args(copy.copy(passed_args))
    assert args.stop_seed >= args.start_seed
    seeds = range(args.start_seed, args.stop_seed + 1)
    if args.folds is None:
        runs = [['--seed', str(seed)] for seed in seeds]
    else:
        assert args.dataset == 'CSL'
        runs = [['--seed', str(seed), '--fold', str(fold)] for seed, fold in product(seeds, range(args.folds))]
    results = [main(parser.parse_args(copy.copy(passed_args) + extra)) for extra in runs]

    def at_best(key):
        return np.array([r[key][r['best']] for r in results], dtype=float)

    def stats(v):
        v = np.asarray(v, dtype=float)
        return np.mean(v), (np.std(v, ddof=1) if len(v) > 1 else float('nan'))
    lines = ['========= Final result ==========', f'Dataset:                {args.dataset}',
             f'SHA:                    {_sha()}', '----------- Best epoch ----------']
    for label, key in (('Train', 'train'), ('Valid', 'val'), ('Test', 'test')):
        m, s = stats(at_best(key))
        lines.append(f'{label + ":":24s}{m} ± {s}')
    lines += [f'Test Min:               {np.min(at_best("test"))}', f'Test Max:               {np.max(at_best("test"))}',
              '----------- Last epoch ----------']
    for label, key in (('Train', 'last_train'), ('Valid', 'last_val'), ('Test', 'last_test')):
        m, s = stats([r[key] for r in results])
        lines.append(f'{label + ":":24s}{m} ± {s}')
    last_test = [r['last_test'] for r in results]
    lines += [f'Test Min:               {np.min(last_test)}', f'Test Max:               {np.max(last_test)}',
              '---------------------------------', '']
    msg = '\n'.join(lines)
    print(msg)
    filename = os.path.join(args.result_folder, f'{args.dataset}-{args.exp_name}/result.txt')
    print('Writing results at: {}'.format(filename))
    with open(filename, 'w') as handle:
        handle.write(msg + '\n' + str(args))
    return results


if __name__ == '__main__':
    passed = sys.argv[1:]
    assert '--seed' not in passed and '--fold' not in passed
    exp_main(passed)
