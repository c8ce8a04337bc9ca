"""CPU check of the index algebra behind `xattn_kv32_kernel` (csrc/xattn.cu): the kernel enumerates the contraction index of
its m16n8k16 MMAs, the key columns of the score tiles and the head-dim columns of the output tiles in PERMUTED orders so that
every operand moves as one 16-byte piece per thread.  Here the warp is emulated in numpy with the PTX fragment layouts of
`mma.sync.aligned.m16n8k16.row.col` (A: a0/a1 = rows g/g+8 of k-slots 2t,2t+1, a2/a3 = k-slots 2t+8,2t+9; B: b0/b1 = the same
k-slots of column g; C: c0,c1 = row g columns 2t,2t+1, c2,c3 = row g+8) and must reproduce plain softmax attention."""
import numpy as np
import pytest


def mma(c, a, b):
    """c[lane][4] += A(16x16) @ B(16x8) with the operands scattered over the 32 lanes as the PTX ISA specifies."""
    A, B = np.zeros((16, 16)), np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, 2 * t:2 * t + 2], A[g + 8, 2 * t:2 * t + 2] = a[lane][0], a[lane][1]
        A[g, 2 * t + 8:2 * t + 10], A[g + 8, 2 * t + 8:2 * t + 10] = a[lane][2], a[lane][3]
        B[2 * t:2 * t + 2, g], B[2 * t + 8:2 * t + 10, g] = b[lane][0], b[lane][1]
    C = A @ B
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        c[lane][0] += C[g, 2 * t]
        c[lane][1] += C[g, 2 * t + 1]
        c[lane][2] += C[g + 8, 2 * t]
        c[lane][3] += C[g + 8, 2 * t + 1]


def piece(row, off):
    """A 16-byte piece = 8 consecutive bf16 = the four 32-bit words x, y, z, w."""
    v = row[off:off + 8]
    return [v[0:2], v[2:4], v[4:6], v[6:8]]


@pytest.mark.parametrize("D", [64, 128])
def test_permuted_fragment_orders_reproduce_attention(D):
    rng = np.random.default_rng(D)
    KK = D // 32
    Q, K, V = rng.standard_normal((16, D)), rng.standard_normal((32, D)), rng.standard_normal((32, D))
    Vt = V.T.copy()
    scale = D ** -0.5
    S = Q @ K.T * scale
    P = np.exp(S - S.max(1, keepdims=True))
    ref = (P / P.sum(1, keepdims=True)) @ V

    # q pieces: thread (g, t) holds elements 32 kk + 8 t .. + 8 of rows g and g + 8
    q = {(lane, j, kk): piece(Q[(lane >> 2) + 8 * j], kk * 32 + 8 * (lane & 3)) for lane in range(32) for j in range(2) for kk in range(KK)}
    s = [[[0.0] * 4 for _ in range(32)] for _ in range(4)]
    for nt in range(4):
        for kk in range(KK):
            # column g of score tile nt is key 8 (g >> 1) + 2 nt + (g & 1): one 16-byte piece of that K row per thread
            kb = {lane: piece(K[8 * ((lane >> 2) >> 1) + 2 * nt + ((lane >> 2) & 1)], kk * 32 + 8 * (lane & 3)) for lane in range(32)}
            for half in range(2):   # words (x, y) feed k-step 2 kk, (z, w) k-step 2 kk + 1
                a = [[q[l, 0, kk][2 * half], q[l, 1, kk][2 * half], q[l, 0, kk][2 * half + 1], q[l, 1, kk][2 * half + 1]] for l in range(32)]
                mma(s[nt], a, [[kb[l][2 * half], kb[l][2 * half + 1]] for l in range(32)])
    sv = np.array(s) * scale                       # [tile][lane][4]
    e = np.zeros_like(sv)
    inv = np.zeros((32, 2))
    for lane in range(32):
        quad = [(lane & ~3) + i for i in range(4)]
        for j in range(2):                          # j = 0: row g (lanes 0, 1 of the accumulator), j = 1: row g + 8
            m = max(sv[nt][qd][2 * j + x] for nt in range(4) for qd in quad for x in range(2))
            for nt in range(4):
                for x in range(2):
                    e[nt][lane][2 * j + x] = np.exp(sv[nt][lane][2 * j + x] - m)
    for lane in range(32):
        quad = [(lane & ~3) + i for i in range(4)]
        for j in range(2):
            inv[lane][j] = 1.0 / sum(e[nt][qd][2 * j + x] for nt in range(4) for qd in quad for x in range(2))
    # the thread's 8 probabilities per row are the keys 8 t .. 8 t + 7: A quads of the two k-steps of P V
    pa = [[[e[2 * kk][l][0:2] * inv[l][0], e[2 * kk][l][2:4] * inv[l][1], e[2 * kk + 1][l][0:2] * inv[l][0], e[2 * kk + 1][l][2:4] * inv[l][1]]
           for l in range(32)] for kk in range(2)]
    o = [[[0.0] * 4 for _ in range(32)] for _ in range(D // 8)]
    for nt in range(D // 8):
        # column g of output tile nt is head-dim 32 (nt >> 2) + 8 (g >> 1) + 2 (nt & 3) + (g & 1): one piece of that V^T row
        vb = {lane: piece(Vt[32 * (nt >> 2) + 8 * ((lane >> 2) >> 1) + 2 * (nt & 3) + ((lane >> 2) & 1)], 8 * (lane & 3)) for lane in range(32)}
        for kk in range(2):
            mma(o[nt], pa[kk], [[vb[l][2 * kk], vb[l][2 * kk + 1]] for l in range(32)])
    out = np.zeros((16, D))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for j in range(2):
            for G in range(D // 32):               # 8 consecutive head-dims per thread and group of four tiles: one 16-byte store
                out[g + 8 * j, 32 * G + 8 * t:32 * G + 8 * t + 8] = [o[4 * G + m][lane][2 * j + x] for m in range(4) for x in range(2)]
    assert np.abs(out - ref).max() < 1e-12
