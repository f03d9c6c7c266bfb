"""Locate the reference package for the drop-in tests (test infrastructure).

Search order: an installed ``yastn``; ``$YASTN_REF``; ``baseline/_ref`` (pip --target install of the reference made
by tools/install_reference.sh; git-ignored, travels to the GPU box); ``/root/reference`` (authoring container only,
never used by ``-m gpu`` tests).  Returns the imported module or None.
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "tests", "stubs")


def load_yastn(allow_reference_checkout=True):
    try:
        importlib.import_module("opt_einsum")
    except ImportError:
        if STUBS not in sys.path:
            sys.path.insert(0, STUBS)
    try:
        return importlib.import_module("yastn")
    except ImportError:
        pass
    cands = [os.environ.get("YASTN_REF"), os.path.join(ROOT, "baseline", "_ref")]
    if allow_reference_checkout:
        cands.append("/root/reference")
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "yastn")):
            sys.path.insert(0, c)
            try:
                return importlib.import_module("yastn")
            except ImportError:
                sys.path.remove(c)
    return None
