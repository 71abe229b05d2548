"""Stand-in for nerfstudio.utils.misc (test infrastructure, see tests/stubs/README.md)."""


def torch_compile(*args, **kwargs):
    """nerfstudio's guarded torch.compile: the reference imports it (model.py:14) but never calls it."""
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]
    return lambda fn: fn
