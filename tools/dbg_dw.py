import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_train_step import _load, _build
from latentfusion_b200 import ops
from latentfusion_b200.train import ReconTrainStep
z, meta = _load()
dev = torch.device('cuda:0')
ops.set_default_precision(1)
res = {}
for mode in (True, False):
    ops._DW_FFMA = mode
    sc, fu, ph, batch = _build(z, meta, dev)
    cfg = meta['cfg']
    step = ReconTrainStep(sc, fu, ph, lr=cfg['lr'], depth_weight=cfg['depth_weight'], mask_weight=cfg['mask_weight'],
                          beta_weight=cfg['beta_weight'], beta_param=cfg['beta_param'], depth_k=cfg['depth_k'])
    ops.KernelTrace.reset(True)
    step.run_iteration(batch, train=True, is_step=False)
    torch.cuda.synchronize()
    print('mode ffma' if mode else 'mode tc', {k: v['calls'] for k, v in ops.KernelTrace.summary().items() if 'dw' in k or 'bwd_weight' in k})
    ops.KernelTrace.reset(False)
    res[mode] = {f'{n}/{k}': p.grad.clone() for n, m in (('sculptor', sc), ('fuser', fu), ('photographer', ph)) for k, p in m.named_parameters()}
for k in res[True]:
    a, b = res[True][k], res[False][k]
    e = float((a - b).norm() / a.norm().clamp(min=1e-9))
    if e > 1e-4:
        print(f'{k:60s} {tuple(a.shape)} rel {e:.3g}')
