"""Import hook behind ``dropin/<subproject>/sitecustomize.py``: resolves exactly the overlaid module names
(``model.tbsrn``, ``model.crnn``, ``loss.text_focus_loss``, ``model.transformer`` ...) to the re-export files of an overlay
directory and leaves every other import to the reference checkout, so the reference's scripts run unedited
(scene-text-telescope/interfaces/base.py:20-23, stroke-level-decomposition/train.py:12, image-ids-CTR/train.py:5).

    PYTHONPATH=dropin/scene-text-telescope python main.py ...            # from inside the reference subproject
    python -m fudanocr_b200.dropin <path/to/subproject/main.py> [args]   # same thing as a launcher
"""
from __future__ import annotations

import importlib.abc
import importlib.util
import os
import runpy
import sys
from typing import Dict, Optional

DROPIN_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin")


def _scan(overlay_dir: str) -> Dict[str, str]:
    """module name -> file for every .py under the overlay (packages map to their __init__.py)"""
    table: Dict[str, str] = {}
    for base, dirs, files in os.walk(overlay_dir):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(base, overlay_dir)
        parts = [] if rel == "." else rel.split(os.sep)
        for f in files:
            if not f.endswith(".py") or (not parts and f == "sitecustomize.py"):
                continue
            if f == "__init__.py":
                if parts:
                    table[".".join(parts)] = os.path.join(base, f)
            else:
                table[".".join(parts + [f[:-3]])] = os.path.join(base, f)
    return table


class OverlayFinder(importlib.abc.MetaPathFinder):
    """first on sys.meta_path: answers only for the overlaid names; parent packages (`model`, `loss`, `utils`) are whatever
    the reference provides (namespace or regular packages) - only the leaf modules are replaced"""

    def __init__(self, overlay_dir: str):
        self.overlay_dir = os.path.abspath(overlay_dir)
        self.table = _scan(self.overlay_dir)

    def find_spec(self, fullname: str, path=None, target=None):
        file = self.table.get(fullname)
        if file is None:
            return None
        if os.path.basename(file) == "__init__.py":
            return importlib.util.spec_from_file_location(fullname, file, submodule_search_locations=[os.path.dirname(file)])
        return importlib.util.spec_from_file_location(fullname, file)


def install(overlay_dir: str) -> OverlayFinder:
    """idempotent: one finder per overlay directory, ahead of the path-based finder"""
    overlay_dir = os.path.abspath(overlay_dir)
    for f in sys.meta_path:
        if isinstance(f, OverlayFinder) and f.overlay_dir == overlay_dir:
            return f
    finder = OverlayFinder(overlay_dir)
    sys.meta_path.insert(0, finder)
    return finder


def overlay_for(subproject_dir: str) -> Optional[str]:
    name = os.path.basename(os.path.abspath(subproject_dir))
    cand = os.path.join(DROPIN_ROOT, name)
    return cand if os.path.isdir(cand) else None


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m fudanocr_b200.dropin <path/to/subproject/script.py> [script args]")
    script = os.path.abspath(argv[0])
    sub = os.path.dirname(script)
    overlay = overlay_for(sub)
    if overlay is None:
        raise SystemExit(f"no overlay for {os.path.basename(sub)!r} under {DROPIN_ROOT}")
    install(overlay)
    os.chdir(sub)                      # the reference resolves ./config, ./dataset/mydata, ./data relative to the cwd
    sys.path.insert(0, sub)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
