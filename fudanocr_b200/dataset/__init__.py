from .dataset import alignCollate_real, alignCollate_syn, resizeNormalize, resize_normalize_batch  # noqa: F401
