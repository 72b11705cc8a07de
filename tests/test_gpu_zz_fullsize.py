"""Full-size parity cases added at the end of round 2 (collected last, after the per-kernel and per-model suites)."""
import pytest
import torch

import oracle
from helpers import assert_close_fro, rel_fro
from test_gpu_history import _layers

pytestmark = pytest.mark.gpu


def test_encoder_config3_batch_slice_matches_oracle():
    """BASELINE configs[2] encoder shape (B = 8192, H = 50, D = 128, 4 heads, 2 layers) run over the whole batch; sequences
    are independent, so the output and the input gradient of a 256-row slice are compared with the oracle on that slice."""
    import two_tower_models_b200 as tt

    B, H, D, heads, L = 8192, 50, 128, 4, 2
    torch.manual_seed(5)
    enc = tt.UserHistoryEncoder(D, H, heads, L, True)
    with torch.no_grad():
        for layer in enc.multihead_attn_layers:
            layer.in_proj_bias.normal_(0, 0.1)
            layer.out_proj.bias.normal_(0, 0.1)
    p = {k: v.detach().clone() for k, v in enc.state_dict().items()}
    x = torch.randn(B, H, D)
    gout = torch.randn(B, 2, D)
    sl = slice(4000, 4256)
    xr = x[sl].clone().requires_grad_(True)
    yr = oracle.history_encoder(xr, _layers(p), heads, oracle.positional_encoding(H, D))
    (yr * gout[sl]).sum().backward()
    enc = enc.cuda()
    xg = x.cuda().requires_grad_(True)
    y = enc(xg)
    (y * gout.cuda()).sum().backward()
    assert rel_fro(y[sl], yr.detach()) < 2e-2, rel_fro(y[sl], yr.detach())
    assert_close_fro(xg.grad[sl], xr.grad, rtol=6e-2, what="dx of the slice")
