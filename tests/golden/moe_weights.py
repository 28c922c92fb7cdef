"""Deterministic values for the variables of the reference's Mixture-of-Experts graph, keyed by the variable's scoped
TensorFlow name.  TEST INFRASTRUCTURE: the "checkpoint" that both the fixture generator (the reference's network text run
on tests/golden/tf1_emulation_nn.py) and the tests (which load the same values into the product's network by name)
restore.  Storing them would take 600 MB for the 7-expert 8^3 network; a name-keyed generator takes none."""
import zlib

import numpy as np


def variable_value(name, shape):
    """float32 value of variable `name` (e.g. 'inception2Expert_3_conv3/bn/gamma')."""
    shape = tuple(int(s) for s in shape)
    rng = np.random.Generator(np.random.PCG64(zlib.crc32(name.encode()) + 20181217))
    u = rng.random(shape, dtype=np.float32)                      # [0, 1)
    leaf = name.rsplit("/", 1)[-1]
    if leaf == "weights":                                        # keeps activations O(1) through conv + ReLU
        fan_in = int(np.prod(shape[:-1]))
        return ((2.0 * u - 1.0) * np.float32(np.sqrt(6.0 / fan_in))).astype(np.float32)
    if leaf in ("gamma", "moving_variance"):
        return (np.float32(0.5) + u).astype(np.float32)
    if leaf in ("biases", "beta", "moving_mean"):
        return (np.float32(0.2) * u - np.float32(0.1)).astype(np.float32)
    raise KeyError("no rule for variable %r" % name)
