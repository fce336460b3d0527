"""CPU model of the fragment-major weight layout and of the operand mapping the megakernel / skinny GEMM
feed to mma.m16n8k16 (sesameai-tts_b200/csrc/api.cu:k_pack_frag, mega.cuh:gemv_groups, skinny.cuh).

This is host-side documentation-as-test: it restates the packing and the PTX fragment layouts in numpy
and checks that the two mappings (16-row groups: weights = A operand, pre-packed quads; 8-row groups:
weights = B operand, activations staged with the pairs of every 8-group as (P0,P2,P1,P3)) reproduce
y = W x exactly.  The CUDA kernels themselves are exercised by the -m gpu parity tests."""
import numpy as np
import pytest


def pack_frag(W, R):
    """k_pack_frag: [group][k block of 32][R*64 bytes] as 16-bit elements; rows beyond W are zero."""
    rows, K = W.shape
    groups, KB = -(-rows // R), K // 32
    Wp = np.zeros((groups * R, K), W.dtype)
    Wp[:rows] = W
    out = np.zeros((groups, KB, R * 32), W.dtype)  # R*64 bytes = R*32 elements per block
    for g in range(groups):
        for kb in range(KB):
            for u in range(R * 4):  # 16-byte units of 8 elements
                L = u & 31
                if R == 16:
                    h = u >> 5
                    r0 = g * 16 + 2 * (L >> 2)
                    k0 = kb * 32 + (L & 3) * 8 + 4 * h
                    unit = np.concatenate([Wp[r0, k0:k0 + 2], Wp[r0 + 1, k0:k0 + 2], Wp[r0, k0 + 2:k0 + 4], Wp[r0 + 1, k0 + 2:k0 + 4]])
                else:
                    unit = Wp[g * 8 + (L >> 2), kb * 32 + (L & 3) * 8: kb * 32 + (L & 3) * 8 + 8]
                out[g, kb, u * 8:(u + 1) * 8] = unit
    return out


def mma_m16n8k16(a, b):
    """PTX mma.sync.m16n8k16 fragment semantics.  a[lane] = 4 registers of 2 elements, b[lane] = 2 registers
    of 2 elements; returns d[lane] = 4 accumulators.  (PTX ISA, 'Matrix fragments for mma.m16n8k16'.)"""
    A = np.zeros((16, 16))
    B = np.zeros((16, 8))
    for lane in range(32):
        g, q = lane >> 2, lane & 3
        for i, (row, col) in enumerate([(g, 2 * q), (g + 8, 2 * q), (g, 2 * q + 8), (g + 8, 2 * q + 8)]):
            A[row, col:col + 2] = a[lane][i]
        for i, k in enumerate([2 * q, 2 * q + 8]):
            B[k:k + 2, g] = b[lane][i]
    D = A @ B
    d = np.zeros((32, 4))
    for lane in range(32):
        g, q = lane >> 2, lane & 3
        d[lane] = [D[g, 2 * q], D[g, 2 * q + 1], D[g + 8, 2 * q], D[g + 8, 2 * q + 1]]
    return d


def regs(unit8):
    """a 16-byte load = 4 registers of 2 consecutive elements"""
    return [unit8[0:2], unit8[2:4], unit8[4:6], unit8[6:8]]


def stage_x(x, natural):
    """stage_x / store_unit: natural order, or the pairs of every 8-group as (P0,P2,P1,P3)"""
    if natural:
        return x.copy()
    y = x.reshape(-1, 4, 2)[:, [0, 2, 1, 3], :]
    return y.reshape(x.shape)


@pytest.mark.parametrize("R", [8, 16])
@pytest.mark.parametrize("nb", [1, 2])
def test_fragment_major_gemv_reproduces_w_times_x(R, nb):
    rng = np.random.default_rng(R * 10 + nb)
    rows, K = 3 * R - 3, 256  # a ragged last group, one 32-wide block per warp
    W = rng.integers(-4, 5, size=(rows, K)).astype(np.float64)  # small integers: exact in any summation order
    X = rng.integers(-4, 5, size=(nb, K)).astype(np.float64)
    Wf = pack_frag(W, R)
    groups = Wf.shape[0]
    Y = np.zeros((nb, groups * R))
    xs = np.stack([stage_x(X[n], natural=(R == 16)) for n in range(nb)])
    for g in range(groups):
        part = np.zeros((8, 16, 2))  # psum[warp][row][n]
        for w in range(8):           # warp w owns k blocks [w * KB/8, (w+1) * KB/8)
            kb = w                   # K = 256: one block per warp
            blk = Wf[g, kb]
            acc = np.zeros((2, 32, 4))  # [lo/hi][lane][fragment]
            a = [None] * 32
            b = [None] * 32
            for half in range(2):
                for lane in range(32):
                    gl, q = lane >> 2, lane & 3
                    xrow = xs[gl if gl < nb else 0]
                    xv = regs(xrow[kb * 32 + q * 8: kb * 32 + q * 8 + 8])
                    if R == 16:  # weights = A operand: the lane's quad of this half is (a0,a1,a2,a3); x natural
                        wq = regs(blk[half * 256 + lane * 8: half * 256 + lane * 8 + 8])
                        a[lane] = wq
                        b[lane] = [xv[2 * half], xv[2 * half + 1]]
                    else:        # weights = B operand: (b0,b1) = pairs (P0,P1) / (P2,P3); the x quad serves both halves
                        wv = regs(blk[lane * 8: lane * 8 + 8])
                        a[lane] = xv
                        b[lane] = [wv[2 * half], wv[2 * half + 1]]
                acc[half] = mma_m16n8k16(a, b)
            for lane in range(32):
                gl, q = lane >> 2, lane & 3
                if R == 16:
                    if q == 0:  # rows 2g (c0,c1) and 2g+1 (c2,c3) x activation rows 0,1
                        s = acc[0][lane] + acc[1][lane]
                        part[w, 2 * gl, 0], part[w, 2 * gl, 1] = s[0], s[1]
                        part[w, 2 * gl + 1, 0], part[w, 2 * gl + 1, 1] = s[2], s[3]
                elif gl < 2:    # lane (g, q): activation row g, rows 2q, 2q+1; lo valid in c0,c1, hi in c2,c3
                    part[w, 2 * q, gl] = acc[0][lane][0] + acc[1][lane][2]
                    part[w, 2 * q + 1, gl] = acc[0][lane][1] + acc[1][lane][3]
        tot = part.sum(axis=0)
        for r in range(R):
            for n in range(nb):
                Y[n, g * R + r] = tot[r, n]
    ref = X @ W.T
    assert np.array_equal(Y[:, :rows], ref)
    assert not Y[:, rows:].any()  # padded rows are zero
