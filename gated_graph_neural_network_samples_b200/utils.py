"""Helpers the graph-model plugins share -- the reference's ``utils.py`` (1-71) without TensorFlow."""
from __future__ import annotations

import queue
import threading

import numpy as np

SMALL_NUMBER = 1e-7  # utils.py:8


def glorot_init(shape):
    """utils.py:11-13: uniform(+-sqrt(6 / (shape[-2] + shape[-1]))) from NumPy's global RNG, float32."""
    r = np.sqrt(6.0 / (shape[-2] + shape[-1]))
    return np.random.uniform(low=-r, high=r, size=shape).astype(np.float32)


class ThreadedIterator:
    """utils.py:16-36: one producer thread fills a bounded queue; ``None`` is the end sentinel.  Unlike the reference, an
    exception in the producer (e.g. a graph that does not fit the node budget) is handed to the consumer and re-raised there
    instead of leaving it blocked on the queue forever."""

    class _Failure:
        def __init__(self, exc):
            self.exc = exc

    def __init__(self, original_iterator, max_queue_size: int = 2):
        self._queue = queue.Queue(maxsize=max_queue_size)
        self._thread = threading.Thread(target=self._worker, args=(original_iterator,), daemon=True)
        self._thread.start()

    def _worker(self, it):
        try:
            for element in it:
                assert element is not None, "By convention, iterator elements must not be None"
                self._queue.put(element, block=True)
        except BaseException as exc:   # noqa: BLE001 -- delivered to the consumer
            self._queue.put(self._Failure(exc), block=True)
            return
        self._queue.put(None, block=True)

    def __iter__(self):
        item = self._queue.get(block=True)
        while item is not None:
            if isinstance(item, self._Failure):
                self._thread.join()
                raise item.exc
            yield item
            item = self._queue.get(block=True)
        self._thread.join()


class _BoundMLP:
    def __init__(self, mlp, keep):
        self.mlp, self.keep = mlp, keep

    def __call__(self, inputs):
        return self.mlp(inputs, self.keep)

    def affine(self):
        """(W, b) of the single affine map when the MLP has no hidden layers (weight dropout applied), else None."""
        import torch
        if len(self.mlp.weights) != 1:
            return None
        W, b = self.mlp.weights[0], self.mlp.biases[0]
        if self.keep < 1.0:
            W = torch.nn.functional.dropout(W, p=1.0 - self.keep, training=True)
        return W, b


class MLP:
    """utils.py:39-71 as torch parameters: ReLU MLP with inverted weight-dropout; the readout uses it with no
    hidden layers (chem_tensorflow.py:153-157), i.e. one affine map."""

    def __init__(self, in_size, out_size, hid_sizes, device):
        import torch
        dims = [in_size] + list(hid_sizes) + [out_size]
        self.weights, self.biases = [], []
        for a, b in zip(dims[:-1], dims[1:]):
            w = (np.sqrt(6.0 / (a + b)) * (2 * np.random.rand(a, b) - 1)).astype(np.float32)  # utils.py:62-63
            self.weights.append(torch.from_numpy(w).to(device).requires_grad_(True))
            self.biases.append(torch.zeros(b, dtype=torch.float32, device=device, requires_grad=True))

    def parameters(self):
        return self.weights + self.biases

    def bind(self, dropout_keep_prob: float = 1.0):
        """The callable handed to ``gated_regression`` (the reference passes ``self.weights['regression_gate_task%i']`` itself,
        whose dropout placeholder is fixed at construction, chem_tensorflow.py:153-160)."""
        return _BoundMLP(self, dropout_keep_prob)

    def __call__(self, inputs, dropout_keep_prob: float = 1.0):
        import torch
        acts = inputs
        hid = inputs
        for W, b in zip(self.weights, self.biases):
            Wd = W if dropout_keep_prob >= 1.0 else torch.nn.functional.dropout(W, p=1.0 - dropout_keep_prob, training=True)
            hid = acts @ Wd + b
            acts = torch.relu(hid)
        return hid  # utils.py:70-71: the last layer's pre-activation
