"""GPU parity on the paths the benchmark actually takes ("What's weak"): device pointers + asynchronous
runs, the whole 53-operator chain, tensors whose byte offsets exceed 2^32, and device bases that are not 16-byte
aligned.  Everything is compared byte for byte with the unmodified reference (oracle/_ref) or the C oracle."""
import numpy as np
import pytest

from oracle import chain_check as CC
from qnnpack_b200 import mobilenet_v2 as M
from tests import cases as CS, util as U

pytestmark = pytest.mark.gpu


def _device_stack(gpu_lib, batch, seed):
    import torch
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    gpu_lib.set_stream(stream.cuda_stream)
    params = M.make_params(seed=seed)
    stack = M.Stack(gpu_lib, params=params)
    cap = stack.max_activation_bytes(batch)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    x_in = torch.randint(0, 256, (batch * 224 * 224 * 3,), dtype=torch.uint8, device=dev, generator=g)
    buf_a = torch.empty(cap, dtype=torch.uint8, device=dev)
    buf_b = torch.empty(cap, dtype=torch.uint8, device=dev)
    stack.setup(batch, buf_a.data_ptr(), buf_b.data_ptr(), first_input=x_in.data_ptr())
    return stack, params, x_in, buf_a, buf_b


@pytest.mark.timeout(600, method="thread")
@pytest.mark.parametrize("batch", [3, 8])
def test_full_chain_device_pointers_async(gpu_lib, batch):
    """All 53 operators chained through two device buffers, asynchronous runs on a user stream — the benchmark's
    path — every layer's output of every image against the reference chain."""
    import torch
    stack, params, x_in, buf_a, buf_b = _device_stack(gpu_lib, batch, seed=batch)
    try:
        res = CC.check_device_stack(stack, params, batch, x_in, buf_a, buf_b, list(range(batch)))
        assert res["mismatches"] == 0, res["failed_layers"]
        assert res["layers"] == 53
    finally:
        stack.delete()
        gpu_lib.set_stream(0)
        torch.cuda.set_stream(torch.cuda.default_stream())


@pytest.mark.timeout(600, method="thread")
def test_layer_beyond_4gib(gpu_lib, oracle_c):
    """b2_expand (1x1, 16 -> 96 @112x112) at batch 3600: the output tensor is 4.33 GB, so row byte offsets exceed 2^32
    and every CTA walks ~190 work items.  Sampled images (first, around the 2^32 crossing, last) vs the oracle."""
    import torch
    dev = torch.device("cuda", 0)
    batch, hw, cin, cout = 3600, 112 * 112, 16, 96
    layer = next(l for l in M.layers() if l.name == "b2_expand")
    kernel, bias, kw = M.layer_params(layer, 7)
    st, op = gpu_lib.create_convolution(kernel, bias, **kw)
    assert st == 0
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    x = torch.randint(0, 256, (batch * hw * cin,), dtype=torch.uint8, device=dev, generator=g)
    y = torch.zeros(batch * hw * cout, dtype=torch.uint8, device=dev)
    assert y.numel() > 2 ** 32
    try:
        assert gpu_lib.setup_convolution(op, batch, 112, 112, x.data_ptr(), cin, y.data_ptr(), cout) == 0
        torch.cuda.synchronize()  # x / y were produced on torch's stream; the library runs on its own
        assert gpu_lib.run(op) == 0
        cross = (2 ** 32) // (hw * cout)  # the image whose output slice contains byte offset 2^32
        for im in (0, cross - 1, cross, cross + 1, batch - 1):
            xi = x[im * hw * cin:(im + 1) * hw * cin].cpu().numpy().reshape(1, 112, 112, cin)
            want = oracle_c.convolution(xi, kernel, bias, **kw)
            got = y[im * hw * cout:(im + 1) * hw * cout].cpu().numpy().reshape(want.shape)
            U.assert_same_bytes(got, want, f"image {im}")
    finally:
        gpu_lib.delete(op)


OFFSET_CASES = [
    CS.conv_case("off_1x1_k16_n96", 2, 9, 9, 1, 16, 96),
    CS.conv_case("off_1x1_k96_n24", 2, 9, 9, 1, 96, 24),
    CS.conv_case("off_stem", 2, 32, 32, 1, 3, 32, ks=(3, 3), stride=(2, 2), pad=(1, 1, 1, 1)),
    CS.conv_case("off_dw_c32", 2, 14, 14, 32, 1, 1, ks=(3, 3), pad=(1, 1, 1, 1)),
    CS.conv_case("off_dw_c32_s2", 2, 14, 14, 32, 1, 1, ks=(3, 3), stride=(2, 2), pad=(1, 1, 1, 1)),
    CS.conv_case("off_3x3_c16", 1, 10, 9, 1, 16, 32, ks=(3, 3), pad=(1, 1, 1, 1)),
]


@pytest.mark.parametrize("in_off,out_off", [(0, 0), (1, 0), (0, 1), (4, 4), (8, 8), (3, 5)])
@pytest.mark.parametrize("case", OFFSET_CASES, ids=lambda c: c["name"])
def test_device_pointers_with_odd_bases(gpu_lib, oracle_c, case, in_off, out_off):
    """Zero-copy device-pointer runs whose input / output bases are offset by a few bytes from the allocation: the
    loader (TMA / cp.async / byte) and store (bulk / 32 / 16 / 8 / 4 / 1 byte) fallbacks are chosen from the real
    alignment, not from cudaMalloc's 256 bytes."""
    import torch
    dev = torch.device("cuda", 0)
    x, k, b, kw = U.conv_setup(case)
    want = U.run_conv(oracle_c, case, x, k, b, kw)
    st, op = gpu_lib.create_convolution(k, b, **kw)
    assert st == 0
    try:
        xd = torch.zeros(x.size + 64, dtype=torch.uint8, device=dev)
        xd[in_off:in_off + x.size] = torch.from_numpy(x.reshape(-1)).to(dev)
        yd = torch.full((want.size + 64,), 0xA5, dtype=torch.uint8, device=dev)
        n, h, w, cs = x.shape
        assert gpu_lib.setup_convolution(op, n, h, w, xd.data_ptr() + in_off, cs, yd.data_ptr() + out_off, want.shape[-1]) == 0
        torch.cuda.synchronize()  # the fills above ran on torch's stream; the library runs on its own
        assert gpu_lib.run_async(op) == 0
        torch.cuda.synchronize()
        got = yd.cpu().numpy()
        U.assert_same_bytes(got[out_off:out_off + want.size].reshape(want.shape), want, case["name"])
        assert (got[:out_off] == 0xA5).all() and (got[out_off + want.size:] == 0xA5).all(), "wrote outside the output"
    finally:
        gpu_lib.delete(op)


@pytest.mark.timeout(600, method="thread")
@pytest.mark.parametrize("batch", [2, 5])
def test_full_network_with_adds_and_pooling(gpu_lib, batch):
    """The real MobileNetV2 graph (64 operators: convolutions, 10 residual adds, global average pool, per-image
    classifier) on device buffers planned by liveness, every node's output of every image against the reference chain."""
    import torch
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    gpu_lib.set_stream(stream.cuda_stream)
    layers = M.network()
    params = [M.layer_params(l, 50 + i)[:2] for i, l in enumerate(layers)]
    net = M.Network(gpu_lib, params=params)
    try:
        cap = net.max_activation_bytes(batch)
        g = torch.Generator(device=dev)
        g.manual_seed(batch)
        x = torch.randint(0, 256, (batch * 224 * 224 * 3,), dtype=torch.uint8, device=dev, generator=g)
        bufs = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(net.nbuf)]
        net.setup(batch, [b.data_ptr() for b in bufs], x.data_ptr())
        res = CC.check_device_network(net, params, batch, x, bufs, list(range(batch)))
        assert res["mismatches"] == 0, res["failed_layers"]
        assert res["layers"] == 64
    finally:
        net.delete()
        gpu_lib.set_stream(0)
        torch.cuda.set_stream(torch.cuda.default_stream())
