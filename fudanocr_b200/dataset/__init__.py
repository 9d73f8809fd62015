from .dataset import (alignCollate_real, alignCollate_syn, finish_collate, hostCollate_real, hostCollate_syn,  # noqa: F401
                      pack_crops, resizeNormalize, resize_normalize_batch)
