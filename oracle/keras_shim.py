"""A minimal stand-in for the ``tensorflow`` / ``keras`` API surface that the reference's
``keras_ocr/recognition.py:54-350`` and ``keras_ocr/detection.py:65-103, 290-468`` touch, so that the
reference's OWN source of ``build_model``, ``_transform``, ``_meshgrid``, ``_repeat``, ``CTCDecoder`` and of the
Keras CRAFT (``build_keras_model``, ``build_vgg_backbone``, ``make_vgg_block``, ``upconv``, ``UpsampleLike``,
``load_torch_weights``) can be executed (AST-lifted at run time by ``oracle/validate_against_reference.py``)
without TensorFlow.

TEST INFRASTRUCTURE (see oracle/__init__.py).  What this pins and what it does not:

* pinned to the reference's source: the graph WIRING of the recognizer (layer order, which layer
  feeds which, BatchNormalization after the ReLU, pool positions, Permute + axis flip, the
  localisation net inside the spatial transformer, ``Add`` vs ``Concatenate`` of the LSTM pairs,
  ``go_backwards`` flags, discarded steps), every line of the STN sampler ``_transform`` (executed
  op by op on the numpy ``tf`` namespace below) and the -1 padding of ``CTCDecoder``;
* NOT pinned: the arithmetic of each Keras layer.  The layers below implement the semantics the Keras
  documentation states (Conv2D "same" padding, BatchNormalization epsilon 1e-3 on the last axis,
  MaxPooling2D "valid", Dense on the last axis, LSTM with gates [i, f, c, o], sigmoid recurrent
  activation, one bias vector, ``go_backwards`` returning the reversed sequence, greedy
  ``ctc_decode`` = argmax, merge repeats, drop the last class) with ``torch.nn.functional`` and
  ``torch.nn.LSTM`` -- independent of the hand-written loops in ``oracle/crnn.py``, but TensorFlow
  itself never ran.

Layers are lazy: calling a layer on a symbolic tensor records a node; ``Model(inputs, outputs)`` evaluates
the recorded graph on real arrays.  Everything is channels-last float32, like Keras.
"""
import types

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------- symbolic graph
class Sym:
    """A symbolic tensor: ``fn(*values of parents)`` gives its value."""

    def __init__(self, fn=None, parents=(), name=None):
        self.fn, self.parents, self.name = fn, tuple(parents), name

    def evaluate(self, feed, cache):
        if id(self) in cache:
            return cache[id(self)]
        if id(self) in feed:
            value = feed[id(self)]
        else:
            assert self.fn is not None, "unfed Input"
            value = self.fn(*[p.evaluate(feed, cache) for p in self.parents])
        cache[id(self)] = value
        return value


def _parents(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


class Layer:
    _counters = {}

    def __init__(self, name=None):
        kind = type(self).__name__.lower()
        if name is None:                          # Keras auto-naming: conv2d, conv2d_1, ...
            k = Layer._counters.get(kind, 0)
            Layer._counters[kind] = k + 1
            name = kind if k == 0 else f"{kind}_{k}"
        self.name = name
        self.weights = {}
        MODEL_LAYERS.append(self)

    def __call__(self, x):
        many = isinstance(x, (list, tuple))
        self.output = Sym(lambda *vals: self.forward(list(vals) if many else vals[0]), _parents(x), self.name)
        self.output.layer = self
        return self.output

    def forward(self, x):                          # subclasses written against the Keras API define call()
        return self.call(x)


MODEL_LAYERS = []          # every layer created since the last reset(), in creation order


def reset():
    MODEL_LAYERS.clear()
    Layer._counters.clear()


def _t(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float32)


def _act(name, x):
    if name is None:
        return x
    if name == "relu":
        return torch.relu(x)
    if name == "softmax":
        return torch.softmax(x, -1)
    raise NotImplementedError(name)


def Input(shape=None, name=None, dtype="float32"):
    return Sym(name=name)


class Permute(Layer):
    def __init__(self, dims, name=None):
        super().__init__(name)
        self.dims = dims

    def forward(self, x):
        return x.permute(0, *self.dims).contiguous()


class Lambda(Layer):
    def __init__(self, function, output_shape=None, name=None):
        super().__init__(name)
        self.function = function

    def forward(self, x):                          # the wrapped function sees numpy arrays (the `tf` namespace below is numpy)
        out = self.function([_np(v) for v in x] if isinstance(x, list) else _np(x))
        out = np.ascontiguousarray(_np(out))
        return torch.from_numpy(out.astype(np.int64)) if out.dtype.kind in "iu" else _t(out)


class Conv2D(Layer):
    def __init__(self, filters=None, kernel_size=None, strides=1, padding="valid", dilation_rate=1, activation=None,
                 name=None):
        super().__init__(name)
        k = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
        assert strides in (1, (1, 1)) and k % 2 == 1 and (padding == "same" or k == 1)
        self.k, self.dil, self.activation = k, dilation_rate, activation

    def set_weights(self, ws):                     # keras order: [kernel (HWIO), bias]
        self.weights["kernel"], self.weights["bias"] = ws

    def forward(self, x):                          # NHWC in, NHWC out; kernel HWIO; "same" = symmetric pad for odd k
        w = _t(self.weights["kernel"]).permute(3, 2, 0, 1)
        y = F.conv2d(x.permute(0, 3, 1, 2), w, _t(self.weights["bias"]), padding=self.dil * (self.k // 2),
                     dilation=self.dil)
        return _act(self.activation, y.permute(0, 2, 3, 1).contiguous())


class BatchNormalization(Layer):
    def __init__(self, epsilon=1e-3, momentum=0.99, axis=-1, name=None):     # keras default epsilon
        super().__init__(name)
        assert axis == -1
        self.epsilon = epsilon

    def set_weights(self, ws):                     # keras order: [gamma, beta, moving_mean, moving_variance]
        for key, value in zip(("gamma", "beta", "moving_mean", "moving_variance"), ws):
            self.weights[key] = value

    def forward(self, x):
        w = self.weights
        inv = _t(w["gamma"]) / torch.sqrt(_t(w["moving_variance"]) + self.epsilon)
        return (x - _t(w["moving_mean"])) * inv + _t(w["beta"])


class MaxPooling2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", name=None):
        super().__init__(name)
        pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)      # noqa: E731
        self.pool = pair(pool_size)
        self.strides = self.pool if strides is None else pair(strides)   # keras default: strides = pool_size
        assert padding == "valid" or (self.strides == (1, 1) and self.pool[0] % 2 == 1)
        self.pad = 0 if padding == "valid" else self.pool[0] // 2        # "same" pads with -inf (ignored by max)

    def forward(self, x):
        return F.max_pool2d(x.permute(0, 3, 1, 2), self.pool, self.strides, self.pad).permute(0, 2, 3, 1).contiguous()


class Activation(Layer):
    def __init__(self, activation, name=None):
        super().__init__(name)
        self.activation = activation

    def forward(self, x):
        return _act(self.activation, x)


class Flatten(Layer):
    def forward(self, x):
        return x.reshape(x.shape[0], -1)


class Reshape(Layer):
    def __init__(self, target_shape, name=None):
        super().__init__(name)
        self.target = tuple(target_shape)

    def forward(self, x):
        return x.reshape(x.shape[0], *self.target)


class Dense(Layer):
    def __init__(self, units, activation=None, kernel_initializer=None, name=None):
        super().__init__(name)
        self.activation = activation

    def forward(self, x):
        return _act(self.activation, x @ _t(self.weights["kernel"]) + _t(self.weights["bias"]))


class LSTM(Layer):
    """keras.layers.LSTM with TF2 defaults, evaluated by torch.nn.LSTM (gate order i, f, g, o = Keras' i, f, c, o)."""

    def __init__(self, units, kernel_initializer=None, go_backwards=False, return_sequences=False, name=None):
        super().__init__(name)
        assert return_sequences
        self.units, self.go_backwards = units, go_backwards

    def forward(self, x):
        w = self.weights
        cell = torch.nn.LSTM(x.shape[-1], self.units, batch_first=True)
        with torch.no_grad():
            cell.weight_ih_l0.copy_(_t(w["kernel"]).t())
            cell.weight_hh_l0.copy_(_t(w["recurrent_kernel"]).t())
            cell.bias_ih_l0.copy_(_t(w["bias"]))
            cell.bias_hh_l0.zero_()
            if self.go_backwards:                  # "process the input sequence backwards and return the reversed sequence"
                x = torch.flip(x, [1])
            return cell(x)[0]


class Add(Layer):
    def forward(self, xs):
        return xs[0] + xs[1]


class Concatenate(Layer):
    def forward(self, xs):
        return torch.cat(list(xs), -1)


class Dropout(Layer):
    def __init__(self, rate, name=None):
        super().__init__(name)

    def forward(self, x):                          # inference
        return x


class Model:
    def __init__(self, inputs, outputs):
        self.inputs, self.outputs = inputs, outputs
        self.input, self.output = inputs, outputs
        self.output_shape = (None, None, None)

    @property
    def layers(self):
        """Layers on the paths from the inputs to the outputs (what keras.Model.layers holds)."""
        seen, found, stack = set(), [], _parents(self.outputs)
        while stack:
            node = stack.pop()
            if id(node) in seen:
                continue
            seen.add(id(node))
            if getattr(node, "layer", None) is not None:
                found.append(node.layer)
            stack.extend(node.parents)
        return found[::-1]

    def get_layer(self, name):
        for layer in self.layers:
            if layer.name == name:
                return layer
        raise ValueError(f"No such layer: {name}")

    def __call__(self, x):
        if isinstance(x, Sym):                     # a model used as a layer (the localisation net)
            return Sym(lambda v: self.predict(v), [x])
        return self.predict(x)

    def predict(self, x):
        ins = _parents(self.inputs)
        vals = _parents(x) if isinstance(self.inputs, (list, tuple)) else [x]
        feed = {id(s): (_t(v) if not isinstance(v, torch.Tensor) else v) for s, v in zip(ins, vals)}
        with torch.no_grad():
            return self.outputs.evaluate(feed, {})


def load_weights(weights):
    """Distribute a flat ``{layer.kind: array}`` dict (keras-ocr_b200/weights.py naming) over the layers created
    since reset(): named layers by name, the auto-named localisation net (conv2d, conv2d_1, dense, dense_1) in
    creation order -> stn.conv_a, stn.conv_b, stn.dense_a, stn.dense_b."""
    auto = {"conv2d": "stn.conv_a", "conv2d_1": "stn.conv_b", "dense": "stn.dense_a", "dense_1": "stn.dense_b"}
    used = set()
    for layer in MODEL_LAYERS:
        prefix = auto.get(layer.name, layer.name) + "."
        for key, value in weights.items():
            if key.startswith(prefix):
                layer.weights[key[len(prefix):]] = value
                used.add(key)
    return used


# ----------------------------------------------------------------------------------------- keras.backend / tf ops
def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def _ctc_decode(y_pred, input_length, greedy=True):
    """keras.backend.ctc_decode(greedy=True): per step argmax of log(y + eps), merge repeated labels, drop the blank
    (= last class); returns ([dense (B, max_len) padded with -1], log_probs) like Keras."""
    y = _np(y_pred)
    best = np.argmax(np.log(y + 1e-7), axis=-1)
    blank = y.shape[-1] - 1
    rows = []
    for b in range(y.shape[0]):
        steps = best[b, : int(_np(input_length)[b])]
        keep = steps[np.concatenate([[True], steps[1:] != steps[:-1]])]
        rows.append(keep[keep != blank])
    width = max((len(r) for r in rows), default=0)
    dense = np.full((y.shape[0], width), -1, dtype=np.int64)
    for b, r in enumerate(rows):
        dense[b, : len(r)] = r
    return [dense], None


def _cast(x, dtype):
    return np.asarray(_np(x)).astype(dtype)


def _slice(x, begin, size):
    x = _np(x)
    idx = tuple(slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
    return x[idx]


def _linspace(start, stop, num):
    # tf.linspace: start + i * (stop - start) / (num - 1), evaluated in the dtype of start (float32)
    step = (np.float32(stop) - np.float32(start)) / np.float32(num - 1)
    return (np.float32(start) + np.arange(num, dtype=np.float32) * step).astype(np.float32)


tf = types.SimpleNamespace(
    ones=lambda shape, dtype="float32": np.ones(shape, dtype=dtype),
    zeros=lambda shape, dtype="float32": np.zeros(shape, dtype=dtype),
    ones_like=lambda x: np.ones_like(_np(x)),
    reshape=lambda x, shape: _np(x).reshape([int(s) for s in np.asarray(shape).reshape(-1)] if not isinstance(shape, (list, tuple))
                                            else [int(s) for s in shape]),
    matmul=lambda a, b: np.matmul(_np(a), _np(b)),
    linspace=_linspace,
    meshgrid=lambda x, y: np.meshgrid(x, y),              # default indexing "xy", as tf.meshgrid
    concat=lambda xs, axis: np.concatenate([_np(x) for x in xs], axis),
    shape=lambda x: np.array(_np(x).shape, dtype=np.int32),
    cast=_cast,
    expand_dims=lambda x, axis: np.expand_dims(_np(x), axis),
    tile=lambda x, multiples: np.tile(_np(x), [int(m) for m in np.asarray(multiples).reshape(-1)]),
    stack=lambda xs: np.array([int(x) for x in xs]),
    slice=_slice,
    floor=lambda x: np.floor(_np(x)),
    clip_by_value=lambda x, lo, hi: np.clip(_np(x), lo, hi),
    range=lambda n: np.arange(int(n), dtype=np.int32),
    gather=lambda params, indices: _np(params)[_np(indices)],
    add_n=lambda xs: sum(xs[1:], xs[0]),
    pad=lambda x, paddings, constant_values=0: np.pad(_np(x), [[int(a), int(b)] for a, b in paddings],
                                                      constant_values=constant_values),
)

def _resize_bilinear(source, size, half_pixel_centers=False):
    """tf.compat.v1.image.resize_bilinear(half_pixel_centers=True) on NHWC = torch bilinear, align_corners=False."""
    assert half_pixel_centers
    x = source if isinstance(source, torch.Tensor) else _t(source)
    y = F.interpolate(x.permute(0, 3, 1, 2), size=(int(size[0]), int(size[1])), mode="bilinear", align_corners=False)
    return y.permute(0, 2, 3, 1).contiguous()


tf.compat = types.SimpleNamespace(v1=types.SimpleNamespace(image=types.SimpleNamespace(resize_bilinear=_resize_bilinear)))
backend = types.SimpleNamespace(shape=tf.shape, cast=_cast, ctc_decode=_ctc_decode, ctc_batch_cost=lambda **kw: None,
                                image_data_format=lambda: "channels_last")
layers = types.SimpleNamespace(Input=Input, Permute=Permute, Lambda=Lambda, Conv2D=Conv2D, Activation=Activation,
                               BatchNormalization=BatchNormalization, MaxPooling2D=MaxPooling2D, Flatten=Flatten,
                               Reshape=Reshape, Dense=Dense, LSTM=LSTM, Add=Add, Concatenate=Concatenate, Dropout=Dropout,
                               Layer=Layer)
keras = types.SimpleNamespace(layers=layers, models=types.SimpleNamespace(Model=Model), backend=backend)
tf.keras = keras
