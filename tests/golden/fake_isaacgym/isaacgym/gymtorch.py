"""wrap/unwrap are identities: the fake gym hands out torch tensors directly."""


def wrap_tensor(t):
    return t


def unwrap_tensor(t):
    return t
