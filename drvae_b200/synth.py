"""Synthetic L1000-shaped minibatches: the inputs bench.py, the tools and the CLI's --synthetic
mode train on.  The reference's real data file is not distributed (workspace/.MISSING_LARGE_BLOBS)."""
import torch


def synthetic_batch(N, dim_x=978, seed=0, dim_y=2):
    """x1 ~ N(0, I); x2 = x1 + 0.3 N(0, I) for pair rows (every second row), zeros for singletons (the reference's
    dataset wrapper stores a zero row for a missing x2, src/DrVAE.py:924); labels on two rows out of three."""
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn(N, dim_x, generator=g)
    x2 = x1 + 0.3 * torch.randn(N, dim_x, generator=g)
    s = torch.zeros(N, dtype=torch.int32)
    y = torch.randint(0, dim_y, (N,), generator=g).int()
    i = torch.arange(N)
    has_x2 = (i % 2 == 0).int()
    has_y = (i % 3 != 0).int()
    x2[has_x2 == 0] = 0
    return dict(x1=x1, x2=x2, s=s, y=y, has_x2=has_x2, has_y=has_y)
