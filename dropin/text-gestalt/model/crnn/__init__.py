from .crnn import CRNN  # model/crnn/__init__.py:1
