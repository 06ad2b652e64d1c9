"""CPU: the weight foldings the engine applies at load time, as plain torch algebra
against the reference's formulation: the ScaledParallelAdapter folded into the FFN GEMMs
(lib/models.py:383-387, 415-421) and the weight-norm of the positional conv (HF:343-355)."""
import torch


def test_adapter_folds_into_ffn():
    torch.manual_seed(0)
    D, F, A, T, scale = 64, 256, 32, 50, 4.0
    u = torch.randn(T, D, dtype=torch.float64)
    W1, b1 = torch.randn(F, D, dtype=torch.float64) / 8, torch.randn(F, dtype=torch.float64) * 0.1
    W2, b2 = torch.randn(D, F, dtype=torch.float64) / 16, torch.randn(D, dtype=torch.float64) * 0.1
    Wd, bd = torch.randn(A, D, dtype=torch.float64) / 8, torch.randn(A, dtype=torch.float64) * 0.1
    Wu, bu = torch.randn(D, A, dtype=torch.float64) / 6, torch.randn(D, dtype=torch.float64) * 0.1
    gelu = torch.nn.functional.gelu
    ref = (gelu(u @ W1.T + b1) @ W2.T + b2) + scale * (torch.relu(u @ Wd.T + bd) @ Wu.T + bu)
    # engine layout: FFN-up rows then adapter-down rows (GELU | ReLU by column range), FFN-down columns
    # then scale * adapter-up columns, one bias
    W_up = torch.cat([W1, Wd], 0)
    b_up = torch.cat([b1, bd], 0)
    W_down = torch.cat([W2, scale * Wu], 1)
    b_down = b2 + scale * bu
    mid = u @ W_up.T + b_up
    mid = torch.cat([gelu(mid[:, :F]), torch.relu(mid[:, F:])], 1)
    got = mid @ W_down.T + b_down
    assert (got - ref).abs().max() < 1e-12


def test_positional_conv_weight_norm_fold():
    torch.manual_seed(1)
    conv = torch.nn.Conv1d(128, 128, kernel_size=16, padding=8, groups=2)
    conv = torch.nn.utils.parametrizations.weight_norm(conv, name="weight", dim=2)     # HF:343-355
    g = conv.parametrizations.weight.original0.detach()          # [1, 1, taps]
    v = conv.parametrizations.weight.original1.detach()          # [O, I/groups, taps]
    with torch.no_grad():
        g.mul_(1 + 0.3 * torch.randn_like(g))
    folded = g * v / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()   # what w2vseg_finalize_weights packs
    assert (folded - conv.weight.detach()).abs().max() < 1e-6
