"""Test helper: differentiable torch restatement of the host-side reparametrisations of ``arch.MLP`` / ``ModifiedMLP`` /
``PirateNet`` (weight_norm: W = g V / ||V||_col, mlp.py:21-53; random_weight: W = g V, mlp.py:56-92), mapping the model's
flat parameter vector onto the oracle's layout [effective linear layers | embeddings | alphas | fourier kernel]."""
import torch

from oracle import ppsci_oracle as O


def oracle_flat(m, flat: torch.Tensor) -> torch.Tensor:
    parts = []
    for i, (a, b) in enumerate(m._shapes):
        W = flat[m._w_off[i]: m._w_off[i] + a * b].view(a, b)
        if m._wn_layer(i):
            g = flat[m._g_off[i]: m._g_off[i] + b]
            W = W * (g if m.random_weight else g / W.norm(p=2, dim=0, keepdim=True))
        parts += [W.reshape(-1), flat[m._b_off[i]: m._b_off[i] + b]]
    if m._n_blocks:
        parts.append(flat[m._alpha_off: m._alpha_off + m._n_blocks])
    for o, n_b in zip(m._beta_off, m._beta_len):
        parts.append(flat[o: o + n_b])
    if m.fourier:
        nf, dh = m._f_shape
        parts.append(flat[m._f_off: m._f_off + nf * dh])
    return torch.cat(parts)


def oracle_loss_and_grad(m, om, exprs, inp, lab):
    """Losses of the oracle network ``om`` at the model's parameters and their gradient w.r.t. ``m.flat`` (chain rule of
    the reparametrisation by autograd)."""
    flat = m.flat.data.detach().cpu().double().clone().requires_grad_(True)
    of = oracle_flat(m, flat)
    assert om.n_params == of.numel(), (om.n_params, of.numel())
    lo, _, go = O.train_forward_backward(om, of.detach(), exprs, {k: inp[k].cpu().double() for k in om.input_keys},
                                         {k: v.cpu().double() for k, v in lab.items()}, None, "mean")
    (g,) = torch.autograd.grad(of, flat, grad_outputs=go)
    return lo, g
