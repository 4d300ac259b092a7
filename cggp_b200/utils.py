"""cggp/utils.py helpers on the hot path."""
import torch


def add_diagonal(matrix: torch.Tensor, diagonal: torch.Tensor) -> torch.Tensor:
    """Returns ``matrix + diag(diagonal)`` as a fresh tensor (cggp/utils.py:11-17)."""
    out = matrix.clone()
    out.diagonal().add_(diagonal.reshape(-1).to(out.dtype))
    return out
