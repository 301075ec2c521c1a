"""Host-side helpers for the chunk8 operand layout (drvae_b200/csrc/common.cuh).

buf[feature // 8][row][feature % 8], bf16.  Used by tests and by state_dict import/export; the
training step itself never packs on the host.
"""
import torch


def round_up(x, m):
    return (x + m - 1) // m * m


def pack_c8(x, rcap=None, fcap=None):
    """[rows, feats] float tensor -> chunk8 bf16 tensor of shape [fcap // 8, rcap, 8]."""
    rows, feats = x.shape
    rcap = rows if rcap is None else rcap
    fcap = round_up(feats, 8) if fcap is None else fcap
    assert fcap % 8 == 0 and rcap >= rows and fcap >= feats
    buf = torch.zeros(rcap, fcap, dtype=torch.bfloat16, device=x.device)
    buf[:rows, :feats] = x.to(torch.bfloat16)
    return buf.view(rcap, fcap // 8, 8).permute(1, 0, 2).contiguous()


def unpack_c8(buf, rows, feats):
    """chunk8 bf16 tensor [nchunks, rcap, 8] -> [rows, feats] float32."""
    nch, rcap, _ = buf.shape
    x = buf.permute(1, 0, 2).reshape(rcap, nch * 8)
    return x[:rows, :feats].float()
