"""A minimal lazy stand-in for the slice of the TensorFlow-1 API that the reference's two propagation hooks use
(``prepare_specific_graph_model`` / ``compute_final_node_representations`` / ``gated_regression`` of chem_tensorflow_sparse.py and
chem_tensorflow_dense.py) -- TEST INFRASTRUCTURE, used only by tests/golden/make_golden.py, in the build container.

TensorFlow 1.3 cannot be installed here, but the reference's graph-BUILDING code is plain Python that only calls ``tf.*``.  With
this module registered as ``tensorflow`` that code runs unmodified: every ``tf.*`` call returns a lazy ``Node``; ``evaluate(node,
feed_dict)`` then computes it with NumPy in float64.  What the fixtures made this way pin is the reference's own dataflow -- which
tensor is gathered, multiplied, concatenated, summed, divided, in which order, which residual states feed which layer, how the
attention softmax is assembled -- i.e. everything in those files.  What they cannot pin is the inside of TensorFlow's own ops: the
``GRUCell`` / ``BasicRNNCell`` arithmetic below is restated from the tensorflow==1.3.0 release (rnn_cell_impl.py), exactly like the
oracle does, and ``unsorted_segment_sum`` adds in index order.
"""
from __future__ import annotations

import contextlib
import types

import numpy as np


class Node:
    """A lazily evaluated tensor: ``fn(*evaluated_args)``."""

    def __init__(self, fn, *args, name=None):
        self.fn, self.args, self.name = fn, args, name

    # arithmetic the reference applies to tensors (+=, -=, /= rebind through these)
    def __add__(self, o): return Node(np.add, self, o)
    def __radd__(self, o): return Node(np.add, o, self)
    def __sub__(self, o): return Node(np.subtract, self, o)
    def __rsub__(self, o): return Node(np.subtract, o, self)
    def __mul__(self, o): return Node(np.multiply, self, o)
    def __rmul__(self, o): return Node(np.multiply, o, self)
    def __truediv__(self, o): return Node(np.divide, self, o)
    def __getitem__(self, idx): return Node(lambda x: x[idx], self)


class Placeholder(Node):
    def __init__(self, dtype, shape=None, name=None):
        super().__init__(None, name=name)
        self.dtype = dtype


class Variable(Node):
    """Eagerly initialised (the reference passes NumPy initial values); promoted to float64 so the fixtures are exact."""

    def __init__(self, initial_value, name=None, **_):
        super().__init__(None, name=name)
        self.value = np.array(initial_value, dtype=np.float64)
        VARIABLES.append(self)


VARIABLES = []          # creation order, like tf.GraphKeys.GLOBAL_VARIABLES
float32, int32 = "float32", "int32"
Tensor = Node


def evaluate(x, feed, memo=None):
    memo = {} if memo is None else memo
    if isinstance(x, (list, tuple)):
        return type(x)(evaluate(v, feed, memo) for v in x)
    if not isinstance(x, Node):
        return x
    if id(x) in memo:
        return memo[id(x)]
    if isinstance(x, Placeholder):
        if x not in feed:
            raise KeyError("placeholder %r was not fed" % x.name)
        v = np.asarray(feed[x])
        v = v.astype(np.float64) if x.dtype == float32 else v.astype(np.int64)
    elif isinstance(x, Variable):
        v = x.value
    else:
        v = x.fn(*[evaluate(a, feed, memo) for a in x.args])
    memo[id(x)] = v
    return v


# ------------------------------------------------------------------------------------------------ graph construction API
def placeholder(dtype, shape=None, name=None): return Placeholder(dtype, shape, name)
def reshape(x, shape): return Node(lambda v, s: np.reshape(v, [int(i) for i in s]), x, list(shape))
def transpose(x, perm): return Node(lambda v: np.transpose(v, perm), x)
def concat(values, axis): return Node(lambda *vs: np.concatenate(vs, axis=axis), *values)
def matmul(a, b): return Node(np.matmul, a, b)
def einsum(eq, *xs): return Node(lambda *vs: np.einsum(eq, *vs), *xs)
def exp(x): return Node(np.exp, x)
def abs(x): return Node(np.abs, x)          # noqa: A001 (mirrors tf.abs)
def square(x): return Node(np.square, x)
def expand_dims(x, axis): return Node(lambda v: np.expand_dims(v, axis), x)
def squeeze(x): return Node(np.squeeze, x)
def gather(params, indices): return Node(lambda p, i: p[i], params, indices)
def ones_like(x, dtype=None): return Node(lambda v: np.ones_like(v, dtype=np.int64 if dtype == int32 else np.float64), x)
def zeros_like(x): return Node(np.zeros_like, x)
def shape(x, out_type=None): return Node(lambda v: np.array(v.shape, dtype=np.int64), x)


def reduce_sum(x, axis=None, keep_dims=False, **_):
    return Node(lambda v: np.sum(v, axis=axis, keepdims=keep_dims), x)


def unsorted_segment_sum(data, segment_ids, num_segments):
    def f(d, ids, n):
        out = np.zeros((int(n),) + d.shape[1:], dtype=np.float64)
        np.add.at(out, ids, d)          # index order == the serial order of TF's CPU kernel
        return out
    return Node(f, data, segment_ids, num_segments)


def unsorted_segment_max(data, segment_ids, num_segments):
    def f(d, ids, n):
        out = np.full((int(n),) + d.shape[1:], np.finfo(np.float64).min, dtype=np.float64)   # TF: lowest() for empty segments
        np.maximum.at(out, ids, d)
        return out
    return Node(f, data, segment_ids, num_segments)


@contextlib.contextmanager
def variable_scope(name, **_):
    yield types.SimpleNamespace(reuse_variables=lambda: None)


def get_variable_scope(): return types.SimpleNamespace(reuse_variables=lambda: None)


def _dropout(x, keep_prob=None, **_):
    def f(v, k):
        if float(k) != 1.0:
            raise NotImplementedError("the fixtures are made in evaluation mode (keep_prob = 1)")
        return v
    return Node(f, x, keep_prob)


def _sigmoid(x): return Node(lambda v: 1.0 / (1.0 + np.exp(-v)), x)
def _tanh(x): return Node(np.tanh, x)
def _relu(x): return Node(lambda v: np.maximum(v, 0.0), x)


# ------------------------------------------------------------------------------------------------ TF-1.3 cells (restated)
CELL_RNG = np.random.RandomState(4242)     # cell kernels are created at first evaluation (their input width is only known then)


def _glorot(shape):
    r = np.sqrt(6.0 / (shape[0] + shape[1]))        # tf.glorot_uniform_initializer, the default of _linear's kernel
    return CELL_RNG.uniform(-r, r, size=shape)


class GRUCell:
    """tf.nn.rnn_cell.GRUCell (1.3): [r|u] = sigmoid([x,h].K_gates + b_gates(=1)), c = act([x, r*h].K_cand + b_cand(=0)),
    h' = u*h + (1-u)*c; returns (h', h')."""

    def __init__(self, num_units, activation=None, **_):
        self.n, self.act = int(num_units), activation or _tanh
        self.vars = None

    def _build(self, din):
        n = self.n
        self.vars = {"gate_kernel": _glorot([din + n, 2 * n]), "gate_bias": np.ones(2 * n), "cand_kernel": _glorot([din + n, n]),
                     "cand_bias": np.zeros(n)}

    def __call__(self, inputs, state):
        def gates(x, h):
            if self.vars is None:
                self._build(x.shape[-1])
            ru = 1.0 / (1.0 + np.exp(-(np.concatenate([x, h], -1) @ self.vars["gate_kernel"] + self.vars["gate_bias"])))
            return ru
        ru = Node(gates, inputs, state)
        r, u = ru[:, :self.n], ru[:, self.n:]
        pre = Node(lambda x, rh: np.concatenate([x, rh], -1) @ self.vars["cand_kernel"] + self.vars["cand_bias"], inputs, r * state)
        c = self.act(pre)
        new_h = u * state + (1 - u) * c
        return new_h, new_h


class CudnnCompatibleGRUCell:
    """tf.contrib.cudnn_rnn.CudnnCompatibleGRUCell (TF >= 1.4; what chem_tensorflow_sparse.py:105-108 instantiates), restated from that
    release: gates as GRUCell (bias 1.0); the candidate has two `_linear`s in their own scopes, candidate/input_projection over the inputs
    and candidate/hidden_projection over the state (biases 0), and the reset gate multiplies the hidden projection AFTER the matmul:
    c = tanh(x.K_in + b_in + r*(h.K_hid + b_hid)); h' = u*h + (1-u)*c; returns (h', h')."""

    def __init__(self, num_units, **_):
        self.n, self.act = int(num_units), _tanh
        self.vars = None

    def _build(self, din):
        n = self.n
        self.vars = {"gate_kernel": _glorot([din + n, 2 * n]), "gate_bias": np.ones(2 * n), "cand_input_kernel": _glorot([din, n]),
                     "cand_bias": np.zeros(n), "cand_hidden_kernel": _glorot([n, n]), "cand_hidden_bias": np.zeros(n)}

    def __call__(self, inputs, state):
        def gates(x, h):
            if self.vars is None:
                self._build(x.shape[-1])
            return 1.0 / (1.0 + np.exp(-(np.concatenate([x, h], -1) @ self.vars["gate_kernel"] + self.vars["gate_bias"])))
        ru = Node(gates, inputs, state)
        r, u = ru[:, :self.n], ru[:, self.n:]
        hi = Node(lambda x, _ru: x @ self.vars["cand_input_kernel"] + self.vars["cand_bias"], inputs, ru)      # ru: the variables exist
        hh = r * Node(lambda h, _ru: h @ self.vars["cand_hidden_kernel"] + self.vars["cand_hidden_bias"], state, ru)
        c = self.act(hi + hh)
        new_h = u * state + (1 - u) * c
        return new_h, new_h


class BasicRNNCell:
    """tf.nn.rnn_cell.BasicRNNCell (1.3): h' = act([x,h].K + b(=0)); returns (h', h')."""

    def __init__(self, num_units, activation=None, **_):
        self.n, self.act = int(num_units), activation or _tanh
        self.vars = None

    def __call__(self, inputs, state):
        def lin(x, h):
            if self.vars is None:
                self.vars = {"rnn_kernel": _glorot([x.shape[-1] + self.n, self.n]), "rnn_bias": np.zeros(self.n)}
            return np.concatenate([x, h], -1) @ self.vars["rnn_kernel"] + self.vars["rnn_bias"]
        new_h = self.act(Node(lin, inputs, state))
        return new_h, new_h


class DropoutWrapper:
    """state_keep_prob only (what the reference passes); keep = 1 is the identity.  vars of the wrapped cell stay reachable."""

    def __init__(self, cell, state_keep_prob=1.0, **_):
        self.cell, self.keep = cell, state_keep_prob

    @property
    def vars(self): return self.cell.vars

    def __call__(self, inputs, state):
        out, new_state = self.cell(inputs, state)
        return out, _dropout(new_state, self.keep)


nn = types.SimpleNamespace(embedding_lookup=lambda params, ids: gather(params, ids), dropout=_dropout, sigmoid=_sigmoid, tanh=_tanh,
                           relu=_relu, rnn_cell=types.SimpleNamespace(GRUCell=GRUCell, BasicRNNCell=BasicRNNCell, DropoutWrapper=DropoutWrapper))
contrib = types.SimpleNamespace(rnn=types.SimpleNamespace(GRUCell=GRUCell),
                                cudnn_rnn=types.SimpleNamespace(CudnnCompatibleGRUCell=CudnnCompatibleGRUCell))


def register_submodules(sys_modules, top="tensorflow"):
    """``import tensorflow.contrib.cudnn_rnn as cudnn_rnn`` (chem_tensorflow_sparse.py:107) needs the dotted names in sys.modules."""
    sys_modules[top + ".contrib"] = contrib
    sys_modules[top + ".contrib.cudnn_rnn"] = contrib.cudnn_rnn
