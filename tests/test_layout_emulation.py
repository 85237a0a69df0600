"""CPU emulation of the conv kernel's data path: the packed weight blob (bk_weights_pack, a host function of
the shared library), the feature / activation operand layouts, the tap row shifts and the row bookkeeping of
bk_forward.cu -- executed with numpy instead of tcgen05.  Comparing the result with the reference's logits
proves the index arithmetic before a GPU is involved.  CPU only."""
import ctypes as C

import numpy as np
import torch

from oracle import nets as onets

GROUP, F_ROWS_B, F_ROWS_G = 5, 121, 605
A_MARGIN, A_ROWS, F_MARGIN, F_ROWS = 12, 536, 24, 653   # emulation buffers (the kernel overlaps the margins)
L0_FULL, L_FULL = 13, 18             # 16 KiB stages (four K steps) per layer; 4 KiB of bias rows follow
L0_BYTES, L_BYTES = L0_FULL * 16384 + 4096, L_FULL * 16384 + 4096


def _fold(sd):
    """BatchNorm folding exactly as bokego_b200.batched.PackedNet does it"""
    ws, bs = [], []
    for i in (0, 3, 6, 9, 12, 15, 18):
        w, b = torch.from_numpy(sd[f"conv.{i}.weight"]).double(), torch.from_numpy(sd[f"conv.{i}.bias"]).double()
        s = torch.from_numpy(sd[f"conv.{i + 1}.weight"]).double() / torch.sqrt(
            torch.from_numpy(sd[f"conv.{i + 1}.running_var"]).double() + 1e-5)
        ws.append(w * s[:, None, None, None])
        bs.append((b - torch.from_numpy(sd[f"conv.{i + 1}.running_mean"]).double()) * s
                  + torch.from_numpy(sd[f"conv.{i + 1}.bias"]).double())
    f32 = lambda t: np.ascontiguousarray(t.float().numpy())
    return f32(ws[0]), f32(torch.stack(ws[1:])), f32(torch.stack(bs)), \
        f32(torch.from_numpy(sd["conv.21.weight"]).reshape(128)), f32(torch.from_numpy(sd["conv.21.bias"]).reshape(81))


def _pack(sd):
    from bokego_b200 import _lib
    L = _lib.lib()
    w0, w16, bias, hw, hb = _fold(sd)
    blob = np.zeros(L.bk_weights_blob_bytes(), np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.bk_weights_pack(p(w0), p(w16), p(bias), p(hw), p(hb), None, p(blob)) == 0
    return blob, bias, hw, hb


def _kstep(blob, layer_off, q):
    """B operand of K step q of a layer as the CTA pair sees it -> float32 [16 k][128 co]:
    stage q/4 = [2 N-halves][4 k-steps][2 k-chunks][64 co][8 k]; each CTA reads its half with k-chunk stride 1024 B"""
    out = np.zeros((16, 128), np.float32)
    for h in range(2):
        off = layer_off + (q >> 2) * 16384 + h * 8192 + (q & 3) * 2048
        half = blob[off: off + 2048].view(np.float16).astype(np.float32).reshape(2, 64, 8)   # [k-chunk][co][k8]
        out[:, 64 * h: 64 * h + 64] = half.transpose(0, 2, 1).reshape(16, 64)
    return out


def _bias_step(blob, layer_off, n_full):
    """the 4 KiB of bias rows behind a layer's stages -> float32 [16 k][128 co]"""
    out = np.zeros((16, 128), np.float32)
    for h in range(2):
        off = layer_off + n_full * 16384 + h * 2048
        half = blob[off: off + 2048].view(np.float16).astype(np.float32).reshape(2, 64, 8)
        out[:, 64 * h: 64 * h + 64] = half.transpose(0, 2, 1).reshape(16, 64)
    return out


ONES = np.zeros(16, np.float32)
ONES[:2] = 1.0                       # the all-ones operand: k = 0, 1 of every row


def _emulate(planes_u8, blob, bias, hw, hb):
    """planes uint8 [5,27,81] (one group) -> logits [5,81] following bk_forward.cu step by step"""
    F = np.zeros((4, F_ROWS, 8), np.float32)
    for b in range(planes_u8.shape[0]):
        for p in range(81):
            r = F_MARGIN + F_ROWS_B * b + 22 + 11 * (p // 9) + p % 9
            v = np.zeros(32, np.float32); v[:27] = planes_u8[b, :, p]
            F[:, r, :] = v.reshape(4, 8)
    A = np.zeros((16, A_ROWS, 8), np.float32)
    # ---- layer 0: two passes (tiles 0..3, tile 4), 13 stages of two 5x5 taps ----
    D = np.zeros((640, 128), np.float32)
    for q in range(50):                                          # K step q: tap q/2, channel chunks 0,1 / 2,3
        W = _kstep(blob, 0, q)
        tap = q >> 1
        off = (tap // 5 - 2) * 11 + (tap % 5 - 2)
        c0 = (q & 1) * 2
        rows = F_MARGIN + np.arange(640) + off
        ok = rows < F_ROWS                                       # rows past the buffer feed invalid outputs only
        a = np.zeros((640, 16), np.float32)
        a[ok] = np.concatenate([F[c0, rows[ok]], F[c0 + 1, rows[ok]]], axis=1)
        D += a @ W
    D += np.tile(ONES, (640, 1)) @ _bias_step(blob, 0, L0_FULL)  # bias hi + lo times the ones operand
    for r0 in range(640):
        board, rem = divmod(r0, F_ROWS_B)
        if board >= GROUP or rem < 22:
            continue
        x, y = divmod(rem - 22, 11)
        if y >= 9:
            continue
        dest = 100 * board + 10 + 10 * x + y
        o = np.maximum(D[r0], 0).astype(np.float16).astype(np.float32)
        A[:, A_MARGIN + dest, :] = o.reshape(16, 8)
    # ---- layers 1..6 ----
    for layer in range(1, 7):
        off_b = L0_BYTES + (layer - 1) * L_BYTES
        D = np.zeros((512, 128), np.float32)
        for q in range(72):                                      # K step q: tap q/8, channel chunks 2*(q%8), +1
            W = _kstep(blob, off_b, q)
            tap = q >> 3
            off = (tap // 3 - 1) * 10 + (tap % 3 - 1)
            rows = A_MARGIN + np.arange(512) + off
            c0 = (q & 7) * 2
            a = np.concatenate([A[c0, rows], A[c0 + 1, rows]], axis=1)
            D += a @ W
        D += np.tile(ONES, (512, 1)) @ _bias_step(blob, off_b, L_FULL)
        logits = np.zeros((GROUP, 81), np.float32)
        for r in range(512):
            board, rem = divmod(r, 100)
            if board >= GROUP or rem < 10 or (rem - 10) % 10 >= 9:
                continue
            act = np.maximum(D[r], 0)
            if layer < 6:
                A[:, A_MARGIN + r, :] = act.astype(np.float16).astype(np.float32).reshape(16, 8)
            else:
                sq = 9 * ((rem - 10) // 10) + (rem - 10) % 10
                logits[board, sq] = act @ hw + hb[sq]
    return logits


def test_emulated_kernel_matches_reference_logits(positions, nets_golden, sd17):
    blob, bias, hw, hb = _pack(sd17)
    src = nets_golden["src"]
    for g0 in (0, 5):
        planes = positions["feats"][src[g0:g0 + 5]]
        got = _emulate(planes, blob, bias, hw, hb)
        want = nets_golden["logits17"][g0:g0 + 5]
        assert np.abs(got - want).max() < 5e-2
        assert (got.argmax(1) == want.argmax(1)).all()


def test_partial_group_is_isolated(positions, nets_golden, sd17):
    """boards of a group do not leak into each other: a group holding 2 boards gives the same rows"""
    blob, bias, hw, hb = _pack(sd17)
    src = nets_golden["src"]
    full = _emulate(positions["feats"][src[0:5]], blob, bias, hw, hb)
    part = _emulate(positions["feats"][src[0:2]], blob, bias, hw, hb)
    assert np.array_equal(full[:2], part[:2])


def test_blob_size_and_value_tail(sd_value):
    from bokego_b200 import _lib
    L = _lib.lib()
    assert L.bk_weights_blob_bytes() == 2036992 and L.bk_feats_conv_bytes(4096) == 820 * 38720
    assert L.bk_feats_conv_bytes(1) == 38720 and L.bk_feats_conv_bytes(0) == 0


def test_bias_rows_reproduce_fp32_bias(sd17):
    """the fp16 hi + lo split of the folded bias (added by the tensor core) carries ~22 bits"""
    blob, bias, _, _ = _pack(sd17)
    for layer in range(7):
        off, n_full = (0, L0_FULL) if layer == 0 else (L0_BYTES + (layer - 1) * L_BYTES, L_FULL)
        b = _bias_step(blob, off, n_full)
        assert np.abs(b[:2].sum(0) - bias[layer]).max() <= 2e-6 * max(1.0, np.abs(bias[layer]).max())
        assert not b[2:].any()


def test_training_raster_of_the_persistent_3x3_kernel():
    """the index arithmetic of bk_train_conv3_tc_kernel (csrc/bk_train_tc.cu) in numpy: GEMM rows are the rows of a padded raster --
    position p, square (x, y) at row 100 p + 10 + 10 x + y, ten zero rows above every board and one zero column to its right -- a tile
    is 128 rows staged with an 11-row halo on both sides, a tap (dx, dy) is the row shift sign * (10 dx + dy) of the staged rows, the
    reduction runs channel group by channel group (the split tail gives each group to a work item of its own and adds the four partial
    results in group order), and 81 of every 100 rows go to the dense [P][81][C] output.  Forward convolution (sign = +1) and data
    gradient (sign = -1, weights mirrored and transposed per tap by the caller) against torch's conv2d / its autograd."""
    rng = np.random.default_rng(3)
    P, Cc = 5, 8                                    # 5 positions = 500 raster rows = 4 tiles, the last one partial; 8 channels = 4 groups of 2
    x = rng.standard_normal((P, 81, Cc)).astype(np.float32)

    def staged_row(rr):                             # the staging warps' gather of raster row rr (zero fill outside the boards)
        if rr < 0:
            return np.zeros(Cc, np.float32)
        p, o = rr // 100, rr % 100 - 10
        xx, yy = o // 10, o % 10                    # floor division: o < 0 (the ten zero rows) gives xx < 0
        ok = p < P and o >= 0 and yy < 9
        return x[p, 9 * xx + yy] if ok else np.zeros(Cc, np.float32)

    def kernel(w_taps, sign, split_groups):
        """w_taps[tap][ci][co]; returns the dense [P][81][Cc] output the result warps write"""
        out = np.zeros((P, 81, Cc), np.float64)
        n_tiles = (100 * P + 127) // 128
        for tile in range(n_tiles):
            R0 = 128 * tile
            stage = np.stack([staged_row(R0 - 11 + r) for r in range(150)])          # rows R0 - 11 .. R0 + 138
            acc = np.zeros((128, Cc), np.float64)
            parts = []
            for g in range(4):                      # channel groups; every group runs its nine taps
                part = np.zeros((128, Cc), np.float64)
                for tap in range(9):
                    ti = tap // 3
                    shift = sign * (10 * (ti - 1) + (tap - 3 * ti - 1))
                    a = stage[11 + shift: 11 + shift + 128, 2 * g: 2 * g + 2].astype(np.float64)
                    part += a @ w_taps[tap][2 * g: 2 * g + 2].astype(np.float64)
                parts.append(part)
                acc += part
            if split_groups:                        # the tail kernel: partial results added in group order
                acc = parts[0] + parts[1] + parts[2] + parts[3]
            for row in range(128):
                rr = R0 + row
                p, o = rr // 100, rr % 100 - 10
                xx, yy = o // 10, o % 10
                if p < P and o >= 0 and yy < 9:
                    out[p, 9 * xx + yy] = acc[row]
        return out

    w = (0.3 * rng.standard_normal((Cc, Cc, 3, 3))).astype(np.float32)              # torch layout [co][ci][kx][ky]
    xt = torch.from_numpy(x).reshape(P, 9, 9, Cc).permute(0, 3, 1, 2).double().requires_grad_(True)
    y = torch.nn.functional.conv2d(xt, torch.from_numpy(w).double(), padding=1)
    want = y.permute(0, 2, 3, 1).reshape(P, 81, Cc).detach().numpy()
    w_fwd = np.stack([w[:, :, tap // 3, tap % 3].T for tap in range(9)])              # [tap][ci][co]
    for split in (False, True):
        assert np.abs(kernel(w_fwd, +1, split) - want).max() < 1e-5
    # data gradient: d x = conv of d y with the mirrored taps and transposed weights (bk_train.cu: bk_train_transpose_kernel), sign = -1
    dy = rng.standard_normal((P, 81, Cc)).astype(np.float32)
    y.backward(torch.from_numpy(dy).reshape(P, 9, 9, Cc).permute(0, 3, 1, 2).double())
    want_dx = xt.grad.permute(0, 2, 3, 1).reshape(P, 81, Cc).numpy()
    x = dy                                                                            # the kernel's input is now d y
    w_bwd = np.stack([w[:, :, tap // 3, tap % 3] for tap in range(9)])                # [tap][co as input channel][ci as output channel]
    assert np.abs(kernel(w_bwd, -1, False) - want_dx).max() < 1e-5
