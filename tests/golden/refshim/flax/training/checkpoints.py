def save_checkpoint(*a, **k):
    raise NotImplementedError("refshim: checkpoints are outside the shim")


def restore_checkpoint(*a, **k):
    raise NotImplementedError("refshim: checkpoints are outside the shim")
