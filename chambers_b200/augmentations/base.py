"""Host-side plumbing shared by every augmentation layer: a small Keras-``Layer``-shaped base
class (``__call__(inputs, training=None)``, ``get_config`` / ``from_config``,
``compute_output_shape``), the ``Chambers>Name`` serialisation registry
(``register_keras_serializable(package="Chambers")`` in the reference,
e.g. image_augmentations.py:62), seeding, and the one function that hands a policy to the CUDA
library.

No TensorFlow is needed; tensors are torch CUDA uint8 NHWC (device path, stream-ordered) or
numpy / CPU torch uint8 arrays (host path through ``chb_policy_apply_host``).
"""

import ctypes
import itertools
import threading

import numpy as np

from .. import _lib

try:  # torch is the device-memory / stream provider, not a compute path
    import torch
except Exception:  # pragma: no cover
    torch = None

# ------------------------------------------------------------------------------------ registry
_REGISTRY = {}
PACKAGE = "Chambers"


def register(cls):
    """Twin of tf.keras.utils.register_keras_serializable(package="Chambers")."""
    _REGISTRY[PACKAGE + ">" + cls.__name__] = cls
    return cls


def serialize(layer):
    """tf.keras.layers.serialize twin: {"class_name": "Chambers>X", "config": {...}}."""
    return {"class_name": PACKAGE + ">" + type(layer).__name__, "config": layer.get_config()}


def deserialize(spec):
    name = spec["class_name"]
    cls = _REGISTRY.get(name) or _REGISTRY.get(PACKAGE + ">" + name)
    if cls is None:
        raise ValueError("Unknown layer: " + str(name))
    return cls.from_config(dict(spec["config"]))


# ------------------------------------------------------------------------- seeding and phase
_state = threading.local()
# seed None = set_random_seed was never called: like the reference's unseeded TF generator, every
# process then draws its own stream (os.urandom, resolved once, at the first layer that needs it).
_global = {"seed": None, "layer_counter": itertools.count(), "learning_phase": False}
_seed_lock = threading.Lock()


def set_random_seed(seed):
    """Twin of chambers.utils.set_random_seed (utils/generic.py:43-51) for this path: fixes the
    Philox key of every layer CREATED afterwards (the seed is captured in ``Layer.__init__``).  A
    layer's key is a function of (seed, creation index since the last call of this function), so a
    program that reseeds with the same value and rebuilds its layers gets the same streams again --
    the behaviour of ``tf.random.set_seed``, whose op counter restarts as well.  The flip side is the
    same as in TensorFlow: a layer kept alive across a reseed with the SAME value shares its key with
    the layer created at the same position afterwards; rebuild the layers after reseeding.  Ranks of
    a data-parallel job must pass ``image_index_base`` (``chambers_b200.sharding.shard_kwargs``) so
    that equal local indices on different ranks draw different schedules."""
    with _seed_lock:
        _global["seed"] = int(seed) & ((1 << 64) - 1)
        _global["layer_counter"] = itertools.count()


def _resolved_seed():
    with _seed_lock:
        if _global["seed"] is None:
            import os
            _global["seed"] = int.from_bytes(os.urandom(8), "little")
        return _global["seed"]


def set_learning_phase(value):
    """Stand-in for tf.keras.backend.learning_phase(), consulted when training is None
    (augmentation_schemes.py:153-154, :205-206).  Defaults to False (inference)."""
    _global["learning_phase"] = bool(value)


def learning_phase():
    return _global["learning_phase"]


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & ((1 << 64) - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & ((1 << 64) - 1)
    return z ^ (z >> 31)


_name_counters = {}


def _auto_name(cls_name):
    snake = "".join(("_" + ch.lower()) if ch.isupper() and i else ch.lower() for i, ch in enumerate(cls_name))
    n = _name_counters.get(snake, 0)
    _name_counters[snake] = n + 1
    return snake if n == 0 else "%s_%d" % (snake, n)


class Layer:
    """The slice of tf.keras.layers.Layer the reference's augmentation layers rely on."""

    def __init__(self, name=None, **kwargs):
        if kwargs:
            allowed = {"trainable", "dtype", "dynamic"}
            unknown = set(kwargs) - allowed
            if unknown:
                raise TypeError("Keyword argument not understood: " + ", ".join(sorted(unknown)))
        self.name = name if name is not None else _auto_name(type(self).__name__)
        self._layer_id = next(_global["layer_counter"])
        self._base_seed = _resolved_seed()
        self._calls = 0
        self._calls_lock = threading.Lock()
        self.last_schedule = None

    # -- Keras surface
    def __call__(self, inputs, *args, **kwargs):
        _check_ndim4(inputs, type(self).__name__)
        return self.call(inputs, *args, **kwargs)

    def call(self, inputs, **kwargs):  # pragma: no cover
        raise NotImplementedError

    def compute_output_shape(self, input_shape):
        return input_shape

    def get_config(self):
        return {"name": self.name}

    @classmethod
    def from_config(cls, config):
        return cls(**config)

    # -- RNG stream of this layer
    def _stream(self, seed=None, call_counter=None):
        if seed is None:
            seed = _splitmix64(self._base_seed ^ ((self._layer_id + 1) * 0xD1B54A32D192ED03 & ((1 << 64) - 1)))
        if call_counter is None:
            with self._calls_lock:  # a layer may be shared by tf.data-style worker threads
                call_counter = self._calls
                self._calls += 1
        return int(seed) & ((1 << 64) - 1), int(call_counter) & 0xFFFFFFFF


def _check_ndim4(x, who):
    shape = getattr(x, "shape", None)
    if shape is None or len(shape) != 4:
        raise ValueError("Input 0 of layer %s is incompatible with the layer: expected ndim=4, found shape %r"
                         % (who, tuple(shape) if shape is not None else None))


def check_uint8(x, who):
    """InputSpec(ndim=4, dtype=tf.uint8) of the policies (augmentation_schemes.py:150, :202)."""
    dt = str(getattr(x, "dtype", "")).replace("torch.", "")
    if dt != "uint8":
        raise ValueError("Input 0 of layer %s is incompatible with the layer: expected dtype=uint8, found dtype=%s"
                         % (who, dt))


def _run_policy_normalized(inputs, transforms, n_draws, elementwise, seed, call_counter, batch_total,
                           image_index_base, replay, record, built, mode):
    if mode not in _lib.NORM_MODES:
        raise ValueError("Unknown mode " + str(mode))
    if torch is None:
        raise _lib.ChambersAugError(4, "torch is required as the device-memory provider")
    is_torch = isinstance(inputs, torch.Tensor)
    x = inputs if is_torch else torch.from_numpy(np.ascontiguousarray(inputs))
    on_device = x.is_cuda
    if not on_device:
        if not torch.cuda.is_available():
            _lib.context(0)  # raises with the real reason
        x = x.cuda(non_blocking=True)
    x = x.contiguous()
    B, H, W, C = (int(d) for d in x.shape)
    if batch_total is None:
        batch_total = B
    pol, _keep, K = built if built is not None else build_policy(transforms, n_draws, elementwise)
    sched_shape = (B, int(n_draws), K, _lib.CHB_SCHED_FIELDS)
    dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    d_replay = d_record = None
    if replay is not None:
        d_replay = torch.as_tensor(np.ascontiguousarray(replay.cpu() if isinstance(replay, torch.Tensor) else replay,
                                                        dtype=np.int32)).reshape(sched_shape).to(x.device)
    if record:
        d_record = torch.zeros(sched_shape, dtype=torch.int32, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        ctx = _lib.context(dev)
        rc = lib.chb_policy_apply_normalized(
            ctx, x.data_ptr(), out.data_ptr(), B, H, W, C, ctypes.byref(pol), _lib.NORM_MODES[mode], int(batch_total),
            int(image_index_base), seed, call_counter,
            d_replay.data_ptr() if d_replay is not None else None,
            d_record.data_ptr() if d_record is not None else None, torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(ctx, rc)
    sched = d_record.cpu().numpy() if record else None
    if on_device:
        return out, sched
    res = out.cpu()
    return (res if is_torch else res.numpy()), sched


# ----------------------------------------------------------------------- policy -> C structs
def build_policy(transforms, n_draws, elementwise):
    """``transforms``: list of lists of (op_layer, probability-or-None).  Returns (ChbPolicy,
    keepalive, K)."""
    n_table = len(transforms)
    if n_table < 1:
        raise ValueError("RandomChoice needs at least one transform")
    arr = (_lib.ChbTransform * n_table)()
    K = 1
    for t, sub in enumerate(transforms):
        if len(sub) > _lib.CHB_MAX_SUBOPS:
            raise ValueError("a transform may hold at most %d ops" % _lib.CHB_MAX_SUBOPS)
        arr[t].n_ops = len(sub)
        K = max(K, len(sub))
        for j, (layer, prob) in enumerate(sub):
            layer._fill_op(arr[t].ops[j])
            arr[t].ops[j].probability = -1.0 if prob is None else float(prob)
    if n_draws * K > _lib.CHB_MAX_CHAIN:
        raise ValueError("n_transforms * ops-per-transform = %d exceeds the fused chain limit %d"
                         % (n_draws * K, _lib.CHB_MAX_CHAIN))
    pol = _lib.ChbPolicy()
    pol.n_table = n_table
    pol.n_draws = int(n_draws)
    pol.elementwise = 1 if elementwise else 0
    pol.table = ctypes.cast(arr, ctypes.POINTER(_lib.ChbTransform))
    return pol, arr, K


def run_policy(inputs, transforms, n_draws, elementwise, seed, call_counter, batch_total=None,
               image_index_base=0, replay=None, record=False, out=None, built=None, normalize=None):
    """Apply RandomChoice(transforms, n_draws, elementwise) to ``inputs`` on the GPU.

    Returns (output, schedule-or-None).  ``inputs``: torch CUDA uint8 NHWC (stream-ordered device
    path) or numpy / CPU torch uint8 (host path, synchronous).

    ``normalize`` ("tf" | "torch" | "caffe", an extension of the reference API): also apply
    ``ImageNetNormalization(mode)`` -- the layer every chambers backbone puts right behind the policy --
    and return float32; on the image-resident engine it is the write epilogue of the policy's last pass
    (``chb_policy_apply_normalized``)."""
    check_uint8(inputs, "RandomChoice")
    if normalize is not None:
        return _run_policy_normalized(inputs, transforms, n_draws, elementwise, seed, call_counter, batch_total,
                                      image_index_base, replay, record, built, normalize)
    B, H, W, C = (int(d) for d in inputs.shape)
    if batch_total is None:
        batch_total = B
    pol, _keep, K = built if built is not None else build_policy(transforms, n_draws, elementwise)
    sched_shape = (B, int(n_draws), K, _lib.CHB_SCHED_FIELDS)
    lib = _lib.load()

    is_torch = torch is not None and isinstance(inputs, torch.Tensor)
    if is_torch and inputs.is_cuda:
        dev = inputs.device.index if inputs.device.index is not None else torch.cuda.current_device()
        x = inputs.contiguous()
        if out is None:
            out = torch.empty_like(x)
        elif not (out.is_cuda and out.is_contiguous() and out.shape == x.shape and out.dtype == torch.uint8):
            raise ValueError("out must be a contiguous CUDA uint8 tensor of the input's shape")
        d_replay = d_record = None
        if replay is not None:
            if isinstance(replay, torch.Tensor) and replay.is_cuda:  # already resident: no host round trip
                d_replay = replay.to(device=x.device, dtype=torch.int32).reshape(sched_shape).contiguous()
            else:
                d_replay = torch.as_tensor(np.ascontiguousarray(replay, dtype=np.int32)).reshape(sched_shape).to(x.device)
        if record:
            d_record = torch.zeros(sched_shape, dtype=torch.int32, device=x.device)
        with torch.cuda.device(dev):
            # (the context is created inside the guard: chb_init makes `dev` current only for its own
            # duration, and torch's notion of the current device stays the caller's)
            ctx = _lib.context(dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.chb_policy_apply(
                ctx, x.data_ptr(), out.data_ptr(), B, H, W, C, ctypes.byref(pol), int(batch_total),
                int(image_index_base), seed, call_counter,
                d_replay.data_ptr() if d_replay is not None else None,
                d_record.data_ptr() if d_record is not None else None, stream)
        _lib.check(ctx, rc)
        return out, (d_record.cpu().numpy() if record else None)

    # host path: numpy array or CPU torch tensor
    if torch is None or not torch.cuda.is_available():
        # chb_init reports the real reason (no device / wrong arch); no CPU fallback exists.
        ctx = _lib.context(0)
    else:
        ctx = _lib.context(torch.cuda.current_device())
    if is_torch:
        x = inputs.contiguous()
        res = torch.empty_like(x, pin_memory=x.is_pinned()) if out is None else out
        in_ptr, out_ptr = x.data_ptr(), res.data_ptr()
    else:
        x = np.ascontiguousarray(inputs)
        res = np.empty_like(x) if out is None else out
        in_ptr, out_ptr = x.ctypes.data, res.ctypes.data
    h_replay = None if replay is None else np.ascontiguousarray(replay, dtype=np.int32).reshape(sched_shape)
    h_record = np.zeros(sched_shape, dtype=np.int32) if record else None
    rc = lib.chb_policy_apply_host(
        ctx, in_ptr, out_ptr, B, H, W, C, ctypes.byref(pol), int(batch_total), int(image_index_base),
        seed, call_counter,
        h_replay.ctypes.data if h_replay is not None else None,
        h_record.ctypes.data if h_record is not None else None)
    _lib.check(ctx, rc)
    return res, h_record
