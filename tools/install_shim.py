#!/usr/bin/env python
"""Copy shim/ over a uzh-rpg/rpg_ramnet `RAM_Net` directory (see shim/README.md).

    python tools/install_shim.py /path/to/RAM_Net [--dry-run]
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ['model/model.py', 'model/loss.py', 'model/metric.py', 'utils/event_tensor_utils.py',
         'trainer/trainer_no_recurrent.py']


def install(target, dry_run=False):
    if not os.path.isfile(os.path.join(target, 'train.py')):
        raise SystemExit(f'{target} does not look like RAM_Net/ (no train.py)')
    for f in FILES:
        src, dst = os.path.join(ROOT, 'shim', f), os.path.join(target, f)
        print(('would copy ' if dry_run else 'copy ') + f'{src} -> {dst}')
        if not dry_run:
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)


if __name__ == '__main__':
    args = [a for a in sys.argv[1:] if not a.startswith('-')]
    if len(args) != 1:
        raise SystemExit(__doc__)
    install(args[0], dry_run='--dry-run' in sys.argv)
