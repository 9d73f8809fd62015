from .crnn import CRNN  # same re-export as scene-text-telescope/model/crnn/__init__.py:1
