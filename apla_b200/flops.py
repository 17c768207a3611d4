"""Algorithmic FLOP model of one APLA fine-tune step: 2*M*N*K per GEMM, attention forward
4*N*D per token per block, attention backward 2.5x forward, input gradients = forward cost, weight gradients only for
the r APLA rows (2*r*D per token) and the head; block 0's attention / qkv / proj input gradients are pruned."""


def flops_per_image(D: int, L: int, patch: int, img: int, r: int, n_classes: int, cls_only_last_block: bool = False):
    """-> (forward, backward) FLOPs per image.

    cls_only_last_block: the engine evaluates the per-token tail of the last block (projection 2*D*D, MLP 16*D*D per
    token forward; MLP input gradients 16*D*D per token backward) on the CLS token only -- the other N-1 tokens' values
    never reach the head; mode 2 (= True) also restricts that block's attention to the CLS query (4*N*D forward, 2.5x
    backward, plus the projection input gradient 2*D*D, per skipped token).  True / 1 / 2 = count what is executed (the conservative numerator for a roofline fraction); False =
    the reference's dense algorithm, SURVEY.md Appendix B."""
    P = (img // patch) ** 2
    N = P + 1
    lin = 24 * D * D
    att = 4 * N * D
    fwd = 2 * 3 * patch * patch * D * P + L * N * (lin + att) + 2 * D * n_classes
    bwd = (L - 1) * N * (lin + 2.5 * att + 2 * r * D) + N * (16 * D * D + 2 * r * D) + 2 * (2 * D * n_classes)
    mode = 2 if cls_only_last_block is True else int(cls_only_last_block)
    if mode >= 1:
        fwd -= (N - 1) * 18 * D * D
        bwd -= (N - 1) * 16 * D * D
    if mode >= 2:               # the last block's attention for the CLS query only, and its projection dgrad on CLS rows
        fwd -= (N - 1) * att
        if L > 1:
            bwd -= (N - 1) * (2.5 * att + 2 * D * D)
    return float(fwd), float(bwd)
