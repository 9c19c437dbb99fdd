"""TEST / BASELINE INFRASTRUCTURE — stages the UNMODIFIED reference for the GPU box.

The reference (NVlabs/latentfusion) is pure Python: there is nothing of it to compile, and the GPU box only receives
/root/repo.  `__graft_entry__.build()` calls `stage()` in the authoring container: a verbatim copy of
/root/reference/latentfusion (+ configs/) lands in oracle/_ref/ — gitignored (never in history, never product
source), not gpurun-ignored (it travels with the snapshot like the built .so).  `bench.py --impl reference` and its
`reference_cuda` context block then run the reference's own estimator through `oracle/ref_import.py`; when the copy
is absent they fall back to the oracle port and say so (`kind: "port"`)."""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference'
DST = os.path.join(HERE, '_ref')


def stage(force=False):
    if not os.path.isdir(os.path.join(SRC, 'latentfusion')):
        return os.path.isdir(os.path.join(DST, 'latentfusion'))
    marker = os.path.join(DST, '.staged')
    if os.path.exists(marker) and not force:
        return True
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(os.path.join(SRC, 'latentfusion'), os.path.join(DST, 'latentfusion'),
                    ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    shutil.copytree(os.path.join(SRC, 'configs'), os.path.join(DST, 'configs'))
    open(marker, 'w').write('verbatim copy of /root/reference/{latentfusion,configs}; see oracle/stage_ref.py\n')
    return True


if __name__ == '__main__':
    print('staged' if stage(force=True) else 'no reference available')
