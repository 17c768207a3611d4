"""Algorithmic FLOP model of one APLA fine-tune step: 2*M*N*K per GEMM, attention forward
4*N*D per token per block, attention backward 2.5x forward, input gradients = forward cost, weight gradients only for
the r APLA rows (2*r*D per token) and the head; block 0's attention / qkv / proj input gradients are pruned."""


def flops_per_image(D: int, L: int, patch: int, img: int, r: int, n_classes: int):
    """-> (forward, backward) FLOPs per image."""
    P = (img // patch) ** 2
    N = P + 1
    lin = 24 * D * D
    att = 4 * N * D
    fwd = 2 * 3 * patch * patch * D * P + L * N * (lin + att) + 2 * D * n_classes
    bwd = (L - 1) * N * (lin + 2.5 * att + 2 * r * D) + N * (16 * D * D + 2 * r * D) + 2 * (2 * D * n_classes)
    return float(fwd), float(bwd)
