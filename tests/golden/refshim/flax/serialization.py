"""Not provided: the checkpoint format is pinned separately (tests/test_checkpoint.py)."""


def to_bytes(target):
    raise NotImplementedError("refshim: flax.serialization is outside the shim")


def from_bytes(target, data):
    raise NotImplementedError("refshim: flax.serialization is outside the shim")
