"""test shim: `import Levenshtein` in the recognisers' util.py"""
from editdistance import eval as distance  # noqa: F401
