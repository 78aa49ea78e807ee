"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- makes the unmodified reference importable on the GPU box.

    python oracle/make_ref.py          (run in the build container; __graft_entry__.build() calls it)

`/root/reference` does not exist on the GPU box, but bench.py's baselines must time the REFERENCE's own
modules there (SURVEY.md 8d "Reference timing": `implicit.LocalPclResnetFC.forward` in PyTorch eager on the same
B200, and on the host cores), not a port.  This recipe copies the few Python files the hot path needs from where
they lie under /root/reference into `oracle/_ref/`, byte for byte.  `oracle/_ref/` is git-ignored (reference
sources never enter this repository's history) but NOT gpurun-ignored, so it travels with the snapshot exactly
like a built .so.  `oracle/ref_loader.py` imports from /root/reference when it exists and from `oracle/_ref/`
otherwise.  Nothing under occlusions-4d_b200/ imports either.
"""
import hashlib
import os
import shutil
import sys

SRC = os.environ.get('O4D_REFERENCE_ROOT', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
FILES = ['__init__.py', 'loss.py', 'model/implicit.py', 'model/model.py', 'model/modules.py',
         'model/point_transformer_layer.py', 'utils/geometry.py', 'utils/utils.py']


def main():
    if not os.path.isfile(os.path.join(SRC, 'model', 'implicit.py')):
        print('make_ref: %s not present, nothing to do' % SRC)
        return 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    lines = []
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        lines.append('%s  %s' % (hashlib.sha256(open(dst, 'rb').read()).hexdigest(), rel))
    for d in ('data', 'eval'):                      # __init__.py:51-55 appends these to sys.path; keep them resolvable
        os.makedirs(os.path.join(DST, d), exist_ok=True)
    with open(os.path.join(DST, 'SHA256SUMS'), 'w') as f:
        f.write('\n'.join(lines) + '\n')
    print('make_ref: copied %d reference files into %s' % (len(FILES), DST))
    return 0


if __name__ == '__main__':
    sys.exit(main())
