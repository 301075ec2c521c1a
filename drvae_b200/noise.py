"""Reference-order ε tape -> the row-indexed ε block the kernels read (include/drvae_b200.h,
drvae_eps_layout_t).

The reference draws its normals group by group (src/DrVAE.py:404-428,
:585-608, src/blocks.py:172); the CUDA step never reorders rows, so for parity runs the tape is
scattered once, on the host, to per-row slots.  Production runs do not use this: the step draws
its own Philox normals.
"""
import torch


def group_indices(kind, has_x2, has_y):
    """Row groups in the reference's processing order: list of (index tensor, is_pair, is_labeled)."""
    hx = has_x2.bool() if has_x2 is not None else None
    hy = has_y.bool() if has_y is not None else None
    nz = lambda m: torch.nonzero(m).view(-1)
    if kind == "drvae":  # LS, US, LP, UP (DrVAE.py:565-608)
        return [(nz(hy & ~hx), False, True), (nz(~hy & ~hx), False, False), (nz(hy & hx), True, True),
                (nz(~hy & hx), True, False)]
    if kind == "pvae":  # S, P (PVAE.py:429-448)
        return [(nz(~hx), False, False), (nz(hx), True, False)]
    if kind == "vfae":  # labeled, unlabeled (VFAE.py:418-436)
        return [(nz(hy), False, True), (nz(~hy), False, False)]
    raise ValueError(kind)


def tape_shapes(kind, has_x2, has_y, dim_x, dim_z1, dim_z3, dim_y, L, noisy):
    """Shapes of the draws of one step, in order."""
    shapes = []
    for idx, pair, lab in group_indices(kind, has_x2, has_y):
        n = len(idx)
        if n == 0:
            continue
        if noisy:
            shapes.append((n, dim_x))
            if pair:
                shapes.append((n, dim_x))
        for _ in range(L):
            shapes.append((n, dim_z1))
            if kind != "vfae":
                if pair:
                    shapes.append((n, dim_z1))
                shapes.append((n, dim_z1))
            if kind != "pvae":
                shapes += [(n, dim_z3)] * (1 if lab else dim_y)
    return shapes


def eps_block_from_tape(plan, draws, has_x2=None, has_y=None, noisy=True):
    """Scatter a recorded tape (list of CPU tensors) into one model's ε block (1-D float32)."""
    kind, L, Ncap = plan.kind, plan.L, plan.Ncap
    X, Z, Z3, Y = plan.dim_x, plan.dim_z1, max(1, plan.dim_z3), max(1, plan.dim_y)
    if kind == "pvae":
        Z3, Y = 1, 1
    el = plan.eps_layout
    blk = torch.zeros(el.total, dtype=torch.float32)
    x1 = blk[el.off_x1:el.off_x1 + Ncap * X].view(Ncap, X)
    x2 = blk[el.off_x2:el.off_x2 + Ncap * X].view(Ncap, X)
    z1 = blk[el.off_z1:el.off_z1 + L * Ncap * Z].view(L, Ncap, Z)
    z2 = blk[el.off_z2:el.off_z2 + L * Ncap * Z].view(L, Ncap, Z)
    z2f = blk[el.off_z2f:el.off_z2f + L * Ncap * Z].view(L, Ncap, Z)
    z3 = blk[el.off_z3:el.off_z3 + L * Ncap * Y * Z3].view(L, Ncap, Y, Z3)
    it = iter(draws)
    for idx, pair, lab in group_indices(kind, has_x2, has_y):
        if len(idx) == 0:
            continue
        if noisy:
            x1[idx] = next(it)
            if pair:
                x2[idx] = next(it)
        for l in range(L):
            z1[l, idx] = next(it)
            if kind != "vfae":
                if pair:
                    z2[l, idx] = next(it)
                z2f[l, idx] = next(it)
            if kind != "pvae":
                for jj in range(1 if lab else Y):
                    z3[l, idx, jj] = next(it)
    rest = list(it)
    if rest:
        raise ValueError("tape has %d unused draws" % len(rest))
    return blk
