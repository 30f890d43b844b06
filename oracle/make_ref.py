"""Recipe that stages the UNMODIFIED reference modules of the hot path under the git-ignored `oracle/_ref/`.

TEST / BASELINE INFRASTRUCTURE.  The reference (Geunwoo-Jeon/LarvaNet) is pure Python, so "building" it means copying
the handful of files the path needs from where they lie under /root/reference:

    models/{__init__,base,LarvaNet,LarvaNetV2,LarvaLeg,LarvaLegV2,edsr}.py   validate.py   utils/{__init__,image_utils}.py
    dataloaders/{__init__,base}.py  (validate.py imports the package)

`oracle/_ref/` is listed in .gitignore (the sources never enter this repo's history) but NOT in .gpurunignore, so the
copy travels to the GPU box with the snapshot; there `tests/test_gpu_fullsize.py` and `bench.py --impl reference` run
the real reference modules (oracle/ref_loader.py) instead of the restatement in oracle/torch_port.py.

    python oracle/make_ref.py            # no-op with a message when /root/reference is absent
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
FILES = [
    'models/__init__.py', 'models/base.py', 'models/LarvaNet.py', 'models/LarvaNetV2.py', 'models/LarvaLeg.py',
    'models/LarvaLegV2.py', 'models/edsr.py', 'validate.py', 'utils/__init__.py', 'utils/image_utils.py',
    'dataloaders/__init__.py', 'dataloaders/base.py',
]


def make_ref(src='/root/reference', quiet=False):
    """Copy the path's reference files into oracle/_ref/.  Returns the directory, or None when `src` is absent."""
    if not os.path.isdir(src):
        if not quiet:
            print(f'oracle/make_ref: {src} not present; keeping whatever oracle/_ref/ holds')
        return DST if os.path.isdir(DST) else None
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.exists(s):
            shutil.copyfile(s, d)
        elif rel.endswith('__init__.py'):
            open(d, 'a').close()     # the reference uses namespace packages in places
        else:
            raise FileNotFoundError(s)
    with open(os.path.join(DST, 'PROVENANCE.txt'), 'w') as f:
        f.write(f'verbatim copies from {src} made by oracle/make_ref.py; not part of this repository\n')
    if not quiet:
        print(f'oracle/make_ref: staged {len(FILES)} reference files under {DST}')
    return DST


if __name__ == '__main__':
    sys.exit(0 if make_ref() else 0)
