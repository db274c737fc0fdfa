"""TEST INFRASTRUCTURE ONLY: dense restatement of the attention core of the reference's FocusedAttn.forward
(transoar/models/necks/focused_decoder.py:238-254) in plain torch: scores for every (query, token) pair, additive -inf
mask outside the query's box, softmax over all tokens, weighted sum.  Differentiable; used as the checker for the fused
RoI kernel (tests/) -- never imported by transoar_b200/."""
import torch


def dense_masked_attention(q, k, v, boxes, grid_shape):
    """q [B,Nq,H,HD] (scaled), k/v [B,Nkv,H,HD], boxes int [Nq,6], grid (X,Y,Z) -> [B,Nq,H*HD]."""
    B, Nq, H, HD = q.shape
    X, Y, Z = grid_shape
    mask = torch.ones(Nq, X, Y, Z, dtype=torch.bool, device=q.device)              # True = masked (:150-157)
    for qi, (x1, y1, z1, x2, y2, z2) in enumerate(boxes.tolist()):
        mask[qi, x1:x2, y1:y2, z1:z2] = False
    add = torch.zeros(Nq, X * Y * Z, dtype=q.dtype, device=q.device)
    add[mask.flatten(1)] = float("-inf")                                           # :243-245
    attn = torch.matmul(q.permute(0, 2, 1, 3), k.permute(0, 2, 3, 1)) + add        # :238
    attn = attn.softmax(-1)                                                        # :247
    return (attn @ v.permute(0, 2, 1, 3)).transpose(1, 2).reshape(B, Nq, H * HD)  # :254
