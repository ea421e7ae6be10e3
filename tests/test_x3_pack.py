"""CPU checks of the HMZ_MODE_FP32X3 weight section (hmz_net_x3.cu: every float32 weight as three bf16 parts, stacked
[w0; w1; w2] along N inside SWIZZLE_128B K-atoms, one block per network and 64-unit chunk): the section is decoded
back through the layout the kernel's descriptors imply and (1) must reproduce every float32 weight exactly as
w0 + w1 + w2, (2) evaluated with the kernel's seven-product schedule in float64 must agree with the reference's recorded
network outputs far inside the 1e-5 gate — i.e. the split itself costs nothing (the tensor core's float32 accumulation
is what the GPU test then adds).  hmz_weights_pack is host code: no GPU needed."""
import ctypes as C

import numpy as np

from oracle import port

SLOT, STRIDE = 30720, 61440
FIRST = {0: "dynamic_net", 1: "rwd_net", 2: "value_net", 3: "policy_net",  # pass order of the kernel
         4: "representation_net"}  # blocks 16..19: the root inference
N2 = {0: 64, 1: 48, 2: 48, 3: 16, 4: 64}


def _bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def _to_bf16(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16)


def _split3(x):
    x = np.asarray(x, np.float32)
    parts = []
    for _ in range(3):
        p = _bf16_to_f32(_to_bf16(x))
        parts.append(p)
        x = (x - p).astype(np.float32)
    return parts


def _decode_block(buf, n_rows):
    """-> (main [3][n_rows][64], extra [3][n_rows][16]) float32 from one packed block."""
    rows = 3 * n_rows
    u16 = buf.view(np.uint16)
    main = np.zeros((rows, 64), np.float32)
    extra = np.zeros((rows, 16), np.float32)
    for row in range(rows):
        for c in range(8):
            off = row * 128 + ((c ^ (row & 7)) << 4)
            main[row, 8 * c:8 * c + 8] = _bf16_to_f32(u16[off // 2:off // 2 + 8])
        for c in range(2):
            off = rows * 128 + (row >> 3) * 256 + c * 128 + (row & 7) * 16
            extra[row, 8 * c:8 * c + 8] = _bf16_to_f32(u16[off // 2:off // 2 + 8])
    return main.reshape(3, n_rows, 64), extra.reshape(3, n_rows, 16)


def _pack(lib, sd, n):
    from muzero_hanoi_b200.engine import STATE_DICT_ORDER

    arrays = [np.ascontiguousarray(sd[k], dtype=np.float32) for k in STATE_DICT_ORDER]
    nbytes = int(lib.hmz_weights_packed_bytes(n, 2))
    assert nbytes > 20 * STRIDE
    host = np.zeros(nbytes, np.uint8)
    table = (C.c_void_p * 20)(*[a.ctypes.data for a in arrays])
    assert lib.hmz_weights_pack(table, n, 2, host.ctypes.data) == 0
    return host


def _layers(host):
    """Decoded parts per network: W1 [3][256][64], X1 [3][256][16], W2 [3][n2][256], X2 [3][n2][16]."""
    out = {}
    for net in range(5):
        n2 = N2[net]
        w1 = np.zeros((3, 256, 64), np.float32)
        x1 = np.zeros((3, 256, 16), np.float32)
        w2 = np.zeros((3, n2, 256), np.float32)
        x2 = None
        for c in range(4):
            blk = host[(net * 4 + c) * STRIDE:(net * 4 + c + 1) * STRIDE]
            m, e = _decode_block(blk[:SLOT], 64)
            w1[:, 64 * c:64 * c + 64], x1[:, 64 * c:64 * c + 64] = m, e
            m, e = _decode_block(blk[SLOT:SLOT + 3 * n2 * 160], n2)
            w2[:, :, 64 * c:64 * c + 64] = m
            if c == 0:
                x2 = e
            else:
                assert not e.any()  # the bias rides with chunk 0 only
        out[net] = (w1, x1, w2, x2)
    return out


import pytest


@pytest.mark.parametrize("n", [1, 5, 12])  # 3N = 3 .. 36 observation columns, zero-padded to K = 64
def test_section_reproduces_every_weight_exactly(lib, n):
    sd = port.make_weights(n, 11)
    L = _layers(_pack(lib, sd, n))
    for net, name in FIRST.items():
        w1, x1, w2, x2 = L[net]
        W1, B1, W2, B2 = (sd[f"{name}.{k}"] for k in ("0.weight", "0.bias", "2.weight", "2.bias"))
        k_in = min(64, W1.shape[1])  # representation_net: 3N input columns, zero-padded to K = 64
        assert np.array_equal(w1.astype(np.float64).sum(0)[:, :k_in], W1[:, :k_in].astype(np.float64)) and not w1[:, :, k_in:].any()
        assert np.array_equal(x1.astype(np.float64).sum(0)[:, 6], B1.astype(np.float64))
        if net == 0:
            assert np.array_equal(x1.astype(np.float64).sum(0)[:, :6], W1[:, 64:70].astype(np.float64))
        else:
            assert not x1[:, :, :6].any()
        assert not x1[:, :, 7:].any()
        out = W2.shape[0]
        assert np.array_equal(w2.astype(np.float64).sum(0)[:out], W2.astype(np.float64)) and not w2[:, out:].any()
        assert np.array_equal(x2.astype(np.float64).sum(0)[:out, 6], B2.astype(np.float64))
        # the parts are what the device-side split produces: bf16(w), bf16(w - w0), bf16(w - w0 - w1)
        for got, want in zip(w1, _split3(W1[:, :k_in])):
            assert np.array_equal(got[:, :k_in], want)


def _linear7(xparts, wparts, ax, xw):
    """The kernel's schedule in float64: block 0 = (x0 + x1 + x2) w0, block 1 = (x0 + x1 + x2) w1 + x0 w2, plus the
    extra slice against [onehot(a), 1]."""
    xs = sum(p.astype(np.float64) for p in xparts)
    b0 = xs @ wparts[0].astype(np.float64).T + ax @ xw[0].astype(np.float64).T
    b1 = xs @ wparts[1].astype(np.float64).T + xparts[0].astype(np.float64) @ wparts[2].astype(np.float64).T
    b1 += ax @ (xw[1].astype(np.float64) + xw[2].astype(np.float64)).T
    return (b0.astype(np.float32) + b1.astype(np.float32)).astype(np.float32)


def test_seven_product_schedule_meets_the_gate(lib, golden):
    g = golden("net_io.npz")
    n = 5
    sd = port.make_weights(n, int(g[f"n{n}_weight_seed"]))
    L = _layers(_pack(lib, sd, n))
    h_in, act = g[f"n{n}_h_in"], g[f"n{n}_action"].astype(np.int64)
    ax = np.zeros((len(act), 16))
    ax[np.arange(len(act)), act] = 1.0
    ax[:, 6] = 1.0

    def mlp(net, x):
        w1, x1, w2, x2 = L[net]
        hid = np.maximum(_linear7(_split3(x), w1, ax if net == 0 else ax * (np.arange(16) == 6), x1), 0.0)
        return _linear7(_split3(hid), w2, ax * (np.arange(16) == 6), x2)

    raw = mlp(0, h_in)[:, :64]
    mn, mx = raw.min(1, keepdims=True), raw.max(1, keepdims=True)
    hn = ((raw - mn) / ((mx - mn) + np.float32(1e-8))).astype(np.float32)
    assert np.all(np.abs(hn - g[f"n{n}_h_out"]) <= 1e-5 * np.abs(g[f"n{n}_h_out"]) + 1e-6)
    logits = mlp(3, hn)[:, :6].astype(np.float64)
    p = np.exp(logits - logits.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    assert np.all(np.abs(p - g[f"n{n}_p"]) <= 1e-5 * np.abs(g[f"n{n}_p"]) + 1e-7)
    import torch

    cls = [v for v in vars(port).values() if isinstance(v, type) and hasattr(v, "_support_to_scalar")][0]
    ref_net = cls(sd)
    for net, key, x in ((1, "r", raw), (2, "v", hn)):
        lg = mlp(net, x)[:, :33]
        got = ref_net._support_to_scalar(torch.from_numpy(lg)).squeeze(-1).numpy()
        ref = g[f"n{n}_{key}"]
        assert np.all(np.abs(got - ref) <= 1e-5 * np.abs(ref) + 2.5e-4)
