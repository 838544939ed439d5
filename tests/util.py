"""Shared helpers for the parity tests."""
import torch

# north_star tolerance for floating point: 1e-5 relative (fp32)
RTOL = 1e-5


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)


def assert_close(a, b, rtol=RTOL, what=""):
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = rel_err(a, b)
    assert err <= rtol, f"{what}: max-normalised error {err:.3e} > {rtol:.1e}"


def random_edge_index(n_src, n_dst, e, gen, self_loops=True):
    row = torch.randint(0, n_src, (e,), generator=gen)
    col = torch.randint(0, n_dst, (e,), generator=gen)
    if not self_loops:
        keep = row != col
        row, col = row[keep], col[keep]
    return torch.stack([row, col])


def copy_params(dst: torch.nn.Module, src: torch.nn.Module) -> None:
    """Load src's parameters into dst (moves to dst's device); materialises lazy parameters first."""
    sd = {k: v.detach().clone() for k, v in src.state_dict().items()}
    from torch.nn.parameter import UninitializedParameter
    for name, p in dst.named_parameters():
        if isinstance(p, UninitializedParameter):
            p.materialize(sd[name].shape)
    dst.load_state_dict(sd)
