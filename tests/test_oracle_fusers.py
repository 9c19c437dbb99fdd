"""CPU: the oracle's LSTM / concat fusers against golden vectors of the unmodified reference
(oracle/make_golden_fusers.py -> tests/golden/fusers_c8_s10.npz)."""
import os

import torch

from tests import parity_helpers as ph
from oracle import lf_oracle as O

PATH = os.path.join(ph.ROOT, 'tests', 'golden', 'fusers_c8_s10.npz')


def test_lstm_and_concat_fusers_vs_golden():
    g = ph.Golden(PATH)
    sd = g.state_dict('lstm')
    out = O.fuse('lstm', g['z_obj'], sd)
    torch.testing.assert_close(out, g['fused.lstm'], atol=2e-5, rtol=1e-4)
    assert torch.equal(O.fuse('concat', g['z_obj']), g['fused.concat'])
    zt = g['z_obj'].clone().requires_grad_(True)
    (O.fuse('lstm', zt, sd) * g['lstm.w']).sum().backward()
    torch.testing.assert_close(zt.grad, g['lstm.grad_z'], atol=2e-5, rtol=1e-3)
