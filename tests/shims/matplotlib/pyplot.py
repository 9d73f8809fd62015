"""test shim: see matplotlib/__init__.py"""
