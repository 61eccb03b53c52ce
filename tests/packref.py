"""Reference (numpy/torch, CPU) construction of the packed GEMM weights the conv kernels consume.
Used by the tests to validate both the conv kernels (with host-packed weights) and the device
packing kernel (bit-for-bit against this file)."""
import numpy as np
import torch

# taps of the 5x5 kernel (as offsets dy in -2..2) that fold onto low-res tap t in -1..1 for output
# phase a in {0,1}: conv5x5(upsample2(x))[2i+a] reads x[(2i+a+dy)>>1]
PHASE_GROUPS = {0: {-1: (-2, -1), 0: (0, 1), 1: (2,)}, 1: {-1: (-2,), 0: (-1, 0), 1: (1, 2)}}


def gemm_weights_fwd5(W):
    """W [Co][Ci][5][5] -> Wg [Co][25][Ci]"""
    return W.permute(0, 2, 3, 1).reshape(W.shape[0], 25, W.shape[1]).contiguous()


def gemm_weights_dgrad5(W):
    """Wg [Ci][25][Co] with flipped taps: dX = conv(dY, flip(W)^T)."""
    return W.flip(2, 3).permute(1, 2, 3, 0).reshape(W.shape[1], 25, W.shape[0]).contiguous()


def phase_weights(W):
    """W [Co][Ci][5][5] -> Weff [(a,b,co)][9][Ci] (fp32 sums of the folded taps)."""
    Co, Ci = W.shape[:2]
    out = torch.zeros(2, 2, Co, 3, 3, Ci, dtype=W.dtype)
    for a in (0, 1):
        for b in (0, 1):
            for ty in (-1, 0, 1):
                for tx in (-1, 0, 1):
                    acc = torch.zeros(Co, Ci, dtype=W.dtype)
                    for dy in PHASE_GROUPS[a][ty]:
                        for dx in PHASE_GROUPS[b][tx]:
                            acc = acc + W[:, :, dy + 2, dx + 2]
                    out[a, b, :, ty + 1, tx + 1, :] = acc
    return out.reshape(4 * Co, 9, Ci)


def gemm_weights_phase_fwd(W, n_pad=None):
    Wg = phase_weights(W)
    if n_pad and n_pad > Wg.shape[0]:
        Wg = torch.cat([Wg, torch.zeros(n_pad - Wg.shape[0], 9, Wg.shape[2])])
    return Wg


def gemm_weights_phase_dgrad(W, k_pad=None):
    """Wg [Ci][9][(a,b,co)] = Weff[(a,b,co)][flipped tap][ci]."""
    Weff = phase_weights(W)                      # [4Co][9][Ci]
    P = Weff.shape[0]
    Wg = Weff.reshape(P, 3, 3, -1).flip(1, 2).reshape(P, 9, -1).permute(2, 1, 0).contiguous()
    if k_pad and k_pad > P:
        Wg = torch.cat([Wg, torch.zeros(Wg.shape[0], 9, k_pad - P)], dim=2)
    return Wg


def pack_kblocks(Wk):
    """Wk [N][ksteps][16] float -> bf16 [N/NB][ksteps][NB/8][2][8][8] flattened (UMMA no-swizzle
    K-major core matrices, SBO 256 B between 8-row groups, LBO 128 B between the two K halves)."""
    N, ks, _ = Wk.shape
    NB = min(N, 128)
    x = Wk.reshape(N // NB, NB // 8, 8, ks, 2, 8)          # [nb][ng][r][ks][kc][e]
    x = x.permute(0, 3, 1, 4, 2, 5).contiguous()           # [nb][ks][ng][kc][r][e]
    return x.to(torch.bfloat16).reshape(-1)


def pack_generic(Wg):
    """Wg [N][taps][C] (C % 16 == 0) -> packed bf16; K step = (tap, 16-channel group)."""
    N, T, C = Wg.shape
    return pack_kblocks(Wg.reshape(N, T * (C // 16), 16))


def wa_groups(ksize, J):
    """Tap groups of the weights-as-A kernel (csrc/wa_groups.cuh): list of (dy, s, lo) in (filter row, chunk) order."""
    pad, nc = ksize // 2, (ksize + J - 1) // J
    out = []
    for row in range(ksize):
        for k in range(nc):
            lo = -pad + k * J
            out.append((row - pad, min(lo + J - 1, pad), lo))
    return out


def pack_wa(Wg, kb=64, J=1):
    """Wg [N][taps][C] (taps = 25 or 9, C % kb == 0, N * J % 128 == 0) -> packed bf16 for conv_wa.cu: K steps ordered
    (kb-channel block, tap group, 16-channel step); GEMM row r of a 128-row block = (channel r // J, j = r % J) carrying
    tap dx = s - j of its group or zeros (CVAE_PACK_KORDER_BLOCK64/32 | CVAE_PACK_STACK2/4)."""
    N, T, C = Wg.shape
    ksize = 5 if T == 25 else 3
    pad, groups, spu, nblk = ksize // 2, wa_groups(ksize, J), kb // 16, C // kb
    rows = N * J
    Wk = torch.zeros(rows, nblk, len(groups), spu, 16)
    r = torch.arange(rows)
    ch = (r // 128) * (128 // J) + (r % 128) // J
    jj = (r % 128) % J
    for gi, (dy, s, lo) in enumerate(groups):
        for j in range(J):
            dx = s - j
            if dx < lo:
                continue
            tap = (dy + pad) * ksize + dx + pad
            sel = jj == j
            Wk[sel, :, gi] = Wg[ch[sel], tap].reshape(-1, nblk, spu, 16)
    return pack_kblocks(Wk.reshape(rows, nblk * len(groups) * spu, 16))


def pack_block64(Wg):
    """pack_wa without stacking, 64-channel blocks (CVAE_PACK_KORDER_BLOCK64)."""
    return pack_wa(Wg, 64, 1)


def pack_pair8_e0(W):
    """Encoder conv 0: W [32][3][5][5]; 8-channel padded source, 13 K steps of two taps each
    (conv_gemm.cu PAIR8 table)."""
    N = W.shape[0]
    Wk = torch.zeros(N, 13, 16)
    def put(step, half, ky, kx):
        Wk[:, step, half * 8: half * 8 + 3] = W[:, :, ky, kx]
    for i in range(10):
        ky, kx = i // 2, (i % 2) * 2
        put(i, 0, ky, kx); put(i, 1, ky, kx + 1)
    for i in (10, 11):
        ky = (i - 10) * 2
        put(i, 0, ky, 4); put(i, 1, ky + 1, 4)
    put(12, 0, 4, 4)
    return pack_kblocks(Wk)
