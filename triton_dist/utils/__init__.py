from .symm import SymmetricHeap, tensor_from_ptr  # noqa: F401
