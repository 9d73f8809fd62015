"""test shim: model/srcnn.py imports matplotlib.pyplot at module level and never plots on the hot path"""
