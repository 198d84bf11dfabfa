import torch


def is_gradtracking(x) -> bool:
    return torch._C._functorch.is_gradtrackingtensor(x) if hasattr(torch._C, "_functorch") else False


def is_batched(x) -> bool:
    return torch._C._functorch.is_batchedtensor(x) if hasattr(torch._C, "_functorch") else False


def is_functorch_tensor(x) -> bool:
    return is_gradtracking(x) or is_batched(x)
