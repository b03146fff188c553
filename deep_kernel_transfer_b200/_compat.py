"""Helpers for the repo-root shims that shadow the reference's module names (``io_utils``, ``data.datamgr`` ...): when
this repo sits IN FRONT of a checkout of the reference on ``sys.path`` the shims serve the DKT path themselves and pass
everything else (argument parsing, the other few-shot methods, the non-episodic loaders -- out of scope here) through
to the reference's own files further down the path."""
import importlib.util
import os
import sys


def find_shadowed(relpath, shim_file):
    """First ``<entry>/<relpath>`` on sys.path that is not the shim itself, or None."""
    me = os.path.realpath(shim_file)
    for entry in sys.path:
        cand = os.path.join(entry or os.getcwd(), relpath)
        if os.path.isfile(cand) and os.path.realpath(cand) != me:
            return cand
    return None


def load_shadowed(relpath, shim_file, alias):
    """Import the module this shim shadows under ``alias`` (its own imports resolve normally, i.e. shims first)."""
    path = find_shadowed(relpath, shim_file)
    if path is None:
        return None
    if alias in sys.modules:
        return sys.modules[alias]
    spec = importlib.util.spec_from_file_location(alias, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        del sys.modules[alias]
        raise
    return mod
