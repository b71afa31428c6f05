"""MobileNetV2-int8 convolution stack as a workload for any qnnpack.h implementation.

Layer shapes: reference bench/convolution.cc:453-537 (the distinct layers) unrolled into network order
with the (t, c, n, s) table of the architecture: 52 convolutions + the 1280->1000 classifier = 53
operators, 300.8 MMAC per 224x224 image.  Residual adds and the global average pool are not
convolutions and are outside the q8gemm/q8conv/q8dwconv path.

Quantisation of the synthetic network: input/kernel zero points 127 as in the reference benches
(bench/convolution.cc:64-76, bench/q8gemm.cc:95-103); weights uniform uint8, bias uniform
[-10000, 10000]; the requantisation scale of each layer is 1/(128*sqrt(K)) so that outputs spread over
the uint8 range instead of saturating (the benches' 0.5*0.5/0.5 saturates nearly every output and is
far from any real network's scale).  Output clamp [0, 255].
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Layer:
    name: str
    kind: str          # "conv" (groups=1, kxk), "pw" (1x1), "dw" (3x3 depthwise), "fc"
    h: int             # input height = width
    cin: int
    cout: int
    k: int
    stride: int
    srcs: tuple = ()   # producers of the inputs (node indices, -1 = network input); () = the previous node
    rowwise: bool = True  # "fc" only: rows of the flat previous output (conv stack) instead of one row per image (network)

    @property
    def groups(self):
        return self.cin if self.kind == "dw" else 1

    @property
    def gic(self):
        return 1 if self.kind == "dw" else self.cin

    @property
    def goc(self):
        return 1 if self.kind == "dw" else self.cout

    @property
    def pad(self):
        return self.k // 2

    @property
    def out_h(self):
        if self.kind == "add":
            return self.h
        if self.kind == "gap":
            return 1
        return (self.h + 2 * self.pad - self.k) // self.stride + 1

    @property
    def k_eff(self):
        return self.k * self.k * self.gic

    def ops(self, batch):
        """2*M*N*K_eff, the reference's own counter (bench/convolution.cc:99-104, bench/q8gemm.cc:108)."""
        if self.kind in ("add", "gap"):
            return batch * self.h * self.h * self.cin  # one add per element
        m = batch * self.out_h * self.out_h if self.kind != "fc" else batch
        return 2 * m * self.cout * self.k_eff

    def algorithmic_bytes(self, batch):
        """input read once + weights/bias + output written once; int32 never counts."""
        if self.kind == "fc":
            return batch * self.cin + self.cout * (self.cin + 4) + batch * self.cout
        if self.kind == "add":
            return 3 * batch * self.h * self.h * self.cin
        if self.kind == "gap":
            return batch * self.h * self.h * self.cin + batch * self.cin
        return (batch * self.h * self.h * self.cin + self.cout * (self.k_eff + 4)
                + batch * self.out_h * self.out_h * self.cout)

    @property
    def in_elems_per_image(self):
        return self.cin if self.kind == "fc" else self.h * self.h * self.cin

    @property
    def out_elems_per_image(self):
        return self.cout if self.kind == "fc" else self.out_h * self.out_h * self.cout


def layers() -> list[Layer]:
    seq = [Layer("stem", "conv", 224, 3, 32, 3, 2)]
    h, c = 112, 32
    idx = 1
    for t, cout, n, s in ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2),
                          (6, 320, 1, 1)):
        for i in range(n):
            stride = s if i == 0 else 1
            hidden = c * t
            if t != 1:
                seq.append(Layer(f"b{idx}_expand", "pw", h, c, hidden, 1, 1))
            seq.append(Layer(f"b{idx}_dw", "dw", h, hidden, hidden, 3, stride))
            h = (h + 2 - 3) // stride + 1
            seq.append(Layer(f"b{idx}_project", "pw", h, hidden, cout, 1, 1))
            c = cout
            idx += 1
    seq.append(Layer("last_1x1", "pw", h, c, 1280, 1, 1))
    seq.append(Layer("classifier", "fc", 1, 1280, 1000, 1, 1))
    return seq


def network() -> list[Layer]:
    """The real MobileNetV2 graph: the 52 convolutions plus the 10 residual adds (stride-1 blocks with equal input and
    output channels), the 7x7 global average pool and the per-image classifier — 64 operators of qnnpack.h."""
    seq = [Layer("stem", "conv", 224, 3, 32, 3, 2, srcs=(-1,))]
    h, c = 112, 32
    idx = 1
    for t, cout, n, s in ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2),
                          (6, 320, 1, 1)):
        for i in range(n):
            stride = s if i == 0 else 1
            hidden = c * t
            block_in = len(seq) - 1
            if t != 1:
                seq.append(Layer(f"b{idx}_expand", "pw", h, c, hidden, 1, 1))
            seq.append(Layer(f"b{idx}_dw", "dw", h, hidden, hidden, 3, stride))
            h = (h + 2 - 3) // stride + 1
            seq.append(Layer(f"b{idx}_project", "pw", h, hidden, cout, 1, 1))
            if stride == 1 and c == cout:
                seq.append(Layer(f"b{idx}_add", "add", h, cout, cout, 1, 1, srcs=(block_in, len(seq) - 1)))
            c = cout
            idx += 1
    seq.append(Layer("last_1x1", "pw", h, c, 1280, 1, 1))
    seq.append(Layer("avgpool", "gap", h, 1280, 1280, 1, 1))
    seq.append(Layer("classifier", "fc", 1, 1280, 1000, 1, 1, rowwise=False))
    return seq


def node_srcs(nodes, i):
    return nodes[i].srcs or (i - 1,)


def plan_buffers(nodes):
    """Liveness-based assignment of node outputs to a small pool of activation buffers: a buffer is reused once every
    consumer of the tensor it holds has run.  -> (buffer index per node, number of buffers)."""
    last_use = {}
    for i in range(len(nodes)):
        for s_ in node_srcs(nodes, i):
            last_use[s_] = i
    free, assign, nbuf = [], [], 0
    holder = {}
    for i in range(len(nodes)):
        if free:
            b = free.pop(0)
        else:
            b, nbuf = nbuf, nbuf + 1
        assign.append(b)
        holder[i] = b
        for s_ in set(node_srcs(nodes, i)):  # inputs whose last consumer this was become free for the NEXT node
            if s_ >= 0 and last_use.get(s_) == i:
                free.append(holder[s_])
    return assign, nbuf


def gemm_sweep() -> list[Layer]:
    """BASELINE.json configs[1]: the distinct 1x1-bottleneck GEMM shapes (+ classifier), each once."""
    seen, out = set(), []
    for l in layers():
        if l.kind in ("pw", "fc"):
            key = (l.h, l.cin, l.cout)
            if key not in seen:
                seen.add(key)
                out.append(l)
    return out


def requant_scale(layer: Layer) -> float:
    return float(np.float32(1.0 / (128.0 * math.sqrt(layer.k_eff))))


def layer_kwargs(layer: Layer):
    """create-kwargs (geometry + quantisation) of a layer for qnnpack_b200.api.QnnpackLibrary."""
    q = dict(izp=127, input_scale=1.0, kzp=127, kernel_scale=requant_scale(layer), ozp=127, output_scale=1.0,
             qmin=0, qmax=255)
    if layer.kind == "fc":
        return q
    p = layer.pad
    return dict(pad=(p, p, p, p), ksize=(layer.k, layer.k), stride=(layer.stride, layer.stride),
                dilation=(1, 1), groups=layer.groups, gic=layer.gic, goc=layer.goc, **q)


def layer_params(layer: Layer, seed: int):
    """-> (kernel uint8, bias int32, create-kwargs) for qnnpack_b200.api.QnnpackLibrary."""
    rng = np.random.default_rng(seed)
    if layer.kind in ("add", "gap"):
        return np.zeros(0, np.uint8), np.zeros(0, np.int32), {}
    if layer.kind == "fc":
        kernel = rng.integers(0, 256, (layer.cout, layer.cin), dtype=np.uint8)
    else:
        kernel = rng.integers(0, 256, (layer.groups, layer.goc, layer.k, layer.k, layer.gic), dtype=np.uint8)
    bias = rng.integers(-10000, 10001, (layer.cout,), dtype=np.int32)
    return kernel, bias, layer_kwargs(layer)


def create_node(lib, layer: Layer, kernel, bias):
    """One operator of the stack in any qnnpack.h implementation (product or reference)."""
    if layer.kind == "add":  # (a - 127) + (b - 127), halved back into the uint8 range
        st, op = lib.create("add_nc_q8", layer.cin, 127, np.float32(1.0), 127, np.float32(1.0), 127, np.float32(2.0), 0, 255)
        if st != 0:
            raise RuntimeError(f"create {layer.name} -> status {st}")
        return op
    if layer.kind == "gap":
        st, op = lib.create("global_average_pooling_nwc_q8", layer.cin, 127, np.float32(1.0), 127, np.float32(1.0), 0, 255)
        if st != 0:
            raise RuntimeError(f"create {layer.name} -> status {st}")
        return op
    kw = layer_kwargs(layer)
    if layer.kind == "fc":
        st, op = lib.create_fully_connected(kernel, bias, **kw)
    else:
        st, op = lib.create_convolution(kernel, bias, **kw)
    if st != 0:
        raise RuntimeError(f"create {layer.name} -> status {st}")
    return op


def setup_node(lib, layer: Layer, op, batch, inputs, out):
    """inputs: [buffer] (NumPy array or device address); batch = images (rows for the classifier)."""
    import ctypes as C

    def ptr(b):
        return b if isinstance(b, np.ndarray) else C.c_void_p(int(b))

    if layer.kind == "add":
        st = lib.setup("add_nc_q8", op, batch * layer.h * layer.h, ptr(inputs[0]), layer.cin, ptr(inputs[1]), layer.cin, ptr(out), layer.cin)
    elif layer.kind == "gap":
        st = lib.setup("global_average_pooling_nwc_q8", op, batch, layer.h * layer.h, ptr(inputs[0]), layer.cin, ptr(out), layer.cin)
    elif layer.kind == "fc":
        st = lib.setup_fully_connected(op, batch, inputs[0], layer.cin, out, layer.cout)
    else:
        st = lib.setup_convolution(op, batch, layer.h, layer.h, inputs[0], layer.cin, out, layer.cout)
    if st != 0:
        raise RuntimeError(f"setup {layer.name} -> status {st}")


def make_params(seed: int = 0, zero: bool = False, only=None):
    """[(kernel uint8, bias int32)] for every layer.  zero=True gives correctly shaped zero arrays — what a
    rank other than 0 holds before qnnpack_b200.shard.replicate_params_from_rank0() fills them."""
    out = []
    for i, l in enumerate(layers() if only is None else only):
        kernel, bias, _ = layer_params(l, seed * 1000 + i)
        out.append((np.zeros_like(kernel), np.zeros_like(bias)) if zero else (kernel, bias))
    return out


class Stack:
    """The 53 operators created once through the C ABI; activations ping-pong between two buffers
    that the caller provides (device addresses for the product, NumPy arrays for the reference)."""

    def __init__(self, lib, seed: int = 0, params=None, only=None):
        self.lib = lib
        self.layers = layers() if only is None else only
        self.ops = []
        for i, l in enumerate(self.layers):
            if params is not None:  # e.g. received from rank 0
                kernel, bias = params[i]
            else:
                kernel, bias, _ = layer_params(l, seed * 1000 + i)
            self.ops.append(create_node(lib, l, kernel, bias))

    def max_activation_bytes(self, batch):
        m = 0
        for l in self.layers:
            if l.kind == "fc":
                m = max(m, batch * l.cin, batch * l.cout)
            else:
                m = max(m, batch * l.h * l.h * l.cin, batch * l.out_h * l.out_h * l.cout)
        return m

    def setup(self, batch, buf_a, buf_b, first_input=None):
        """Chains the operators through two ping-pong activation buffers (device addresses as int, or
        uint8 NumPy arrays, each of at least max_activation_bytes(batch)).  With ``first_input`` the first
        layer reads that (never overwritten) buffer and writes buf_a; otherwise it reads buf_a and writes
        buf_b.  Returns 0 if the final output lies in buf_a, 1 if in buf_b."""
        bufs = (buf_a, buf_b)
        shift = 1 if first_input is not None else 0
        where = 0
        for i, (l, op) in enumerate(zip(self.layers, self.ops)):
            src = first_input if (i == 0 and first_input is not None) else bufs[(i - shift) % 2]
            where = (i - shift + 1) % 2
            dst = bufs[where]
            if l.kind == "fc":
                st = self.lib.setup_fully_connected(op, batch, src, l.cin, dst, l.cout)
            else:
                st = self.lib.setup_convolution(op, batch, l.h, l.h, src, l.cin, dst, l.cout)
            if st != 0:
                raise RuntimeError(f"setup {l.name} -> status {st}")
        return where

    def run(self, asynchronous=False, hook=None):
        for i, op in enumerate(self.ops):
            if hook is not None:
                hook(i, 0)
            st = self.lib.run_async(op) if asynchronous else self.lib.run(op)
            if st != 0:
                raise RuntimeError(f"run {self.layers[i].name} -> status {st}")
            if hook is not None:
                hook(i, 1)

    def delete(self):
        for op in self.ops:
            self.lib.delete(op)
        self.ops = []

    def total_ops(self, batch):
        return sum(l.ops(batch) for l in self.layers)

    def total_bytes(self, batch):
        return sum(l.algorithmic_bytes(batch) for l in self.layers)


class Network:
    """The real MobileNetV2 graph (network()) on any qnnpack.h implementation: operators created once, activations in
    a liveness-planned pool of buffers (plan_buffers) that the caller provides."""

    def __init__(self, lib, seed: int = 0, params=None):
        self.lib = lib
        self.layers = network()
        self.assign, self.nbuf = plan_buffers(self.layers)
        self.ops = []
        for i, l in enumerate(self.layers):
            kernel, bias = params[i] if params is not None else layer_params(l, seed * 1000 + i)[:2]
            self.ops.append(create_node(lib, l, kernel, bias))

    def max_activation_bytes(self, batch):
        return max(batch * l.out_elems_per_image for l in self.layers)

    def setup(self, batch, buffers, first_input):
        """buffers: self.nbuf device addresses (or NumPy arrays) of max_activation_bytes(batch) each."""
        for i, (l, op) in enumerate(zip(self.layers, self.ops)):
            ins = [first_input if s_ < 0 else buffers[self.assign[s_]] for s_ in node_srcs(self.layers, i)]
            setup_node(self.lib, l, op, batch, ins, buffers[self.assign[i]])

    def run(self, asynchronous=False, hook=None):
        for i, op in enumerate(self.ops):
            if hook is not None:
                hook(i, 0)
            st = self.lib.run_async(op) if asynchronous else self.lib.run(op)
            if st != 0:
                raise RuntimeError(f"run {self.layers[i].name} -> status {st}")
            if hook is not None:
                hook(i, 1)

    def delete(self):
        for op in self.ops:
            self.lib.delete(op)
        self.ops = []

    def total_ops(self, batch):
        return sum(l.ops(batch) for l in self.layers)

    def total_bytes(self, batch):
        return sum(l.algorithmic_bytes(batch) for l in self.layers)
