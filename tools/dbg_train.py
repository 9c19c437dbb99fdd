import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_train_step import _load, _build
from latentfusion_b200 import ops
from latentfusion_b200.train import ReconTrainStep
from latentfusion_b200.three.batchview import b2bv, bv2b
from latentfusion_b200.recon.models import gan_normalize
z, meta = _load()
dev = torch.device('cuda:0')
ops.set_default_precision(0)
sc, fu, ph, batch = _build(z, meta, dev)
def err(a, k):
    r = torch.from_numpy(np.array(z[k]))
    a = a.detach().cpu().reshape(r.shape)
    print(k, 'maxabs', float((a - r).abs().max()), 'ref max', float(r.abs().max()))
with torch.no_grad():
    for mode in (False, True):
        for m in (sc, fu, ph): m.train(mode)
        print('train mode', mode)
        color, mask = batch['in']['image'], batch['in']['mask']
        x = torch.cat([bv2b(color), gan_normalize(bv2b(mask))], 1)
        zv, _, _ = sc(x, batch['in']['camera'])
        zo, _ = fu(b2bv(zv, color.shape[1]), [], [], None)
        err(zo, 'fwd.z_obj')
        zo2, _ = sc.encode(fu, batch['in']['camera'], color, None, mask)
        err(zo2, 'fwd.z_obj')
        zg = torch.from_numpy(np.array(z['fwd.z_obj'])).to(dev)
        y, _, _ = ph.decode(zg, batch['out_gt']['camera'], interpret_logits=True)
        err(y['depth'], 'fwd.depth'); err(y['mask_logits'], 'fwd.mask_logits')
        # per-object decode
        cam = batch['out_gt']['camera']
        for b in range(2):
            yb, _, _ = ph.decode(zg[b:b+1], cam[b*2:(b+1)*2], interpret_logits=True)
            r = torch.from_numpy(np.array(z['fwd.depth']))[b:b+1]
            print(' obj', b, float((yb['depth'].cpu() - r).abs().max()))
