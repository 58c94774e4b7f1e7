"""Import the reference's own ``tools.urdf_parsing`` without copying or modifying it.

The north-star requires that ``assets/trex.urdf`` is loaded through the reference
repo's own parser (``tools/urdf_parsing.py:223-239``).  That parser cannot be
imported as-is on Python >= 3.11: ``tools/geometry.py:53`` uses a dataclass
instance as a field default, which raises ``ValueError``.
This shim reads ``tools/geometry.py`` from the reference checkout, rewrites that
one line *in memory* (``default_factory=Transform``), registers the result as
``<pkg>.geometry`` and then imports the *unmodified* ``urdf_parsing.py`` from disk.

Nothing under the reference tree is written to, and no reference source is
stored in this repository.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

_PKG = "_trex_reference_tools"

_GEOMETRY_BAD = "origin: Transform = Transform()"
_GEOMETRY_FIX = "origin: Transform = dataclasses.field(default_factory=Transform)"


class ReferenceToolsNotFound(ImportError):
    pass


def find_tools_dir(urdf_path: str | None = None) -> str:
    """Locate the reference ``tools/`` directory.

    Search order: ``$TREX_GYM_REFERENCE/tools``, ``<urdf dir>/../tools`` (the
    layout of the reference checkout: ``assets/trex.urdf`` next to ``tools/``),
    ``/root/reference/tools``.
    """
    cands = []
    env = os.environ.get("TREX_GYM_REFERENCE")
    if env:
        cands.append(os.path.join(env, "tools"))
    if urdf_path:
        cands.append(os.path.join(os.path.dirname(os.path.abspath(urdf_path)), "..", "tools"))
    cands.append("/root/reference/tools")
    for c in cands:
        if os.path.isfile(os.path.join(c, "urdf_parsing.py")) and os.path.isfile(
            os.path.join(c, "geometry.py")
        ):
            return os.path.normpath(c)
    raise ReferenceToolsNotFound(
        "reference tools/urdf_parsing.py not found (looked in: %s); set "
        "TREX_GYM_REFERENCE to the trex-gym checkout" % ", ".join(cands)
    )


def load_urdf_parsing(tools_dir: str):
    """Return the reference ``urdf_parsing`` module loaded from ``tools_dir``."""
    key = _PKG + "::" + tools_dir
    cached = sys.modules.get(key)
    if cached is not None:
        return cached

    pkg = types.ModuleType(_PKG)
    pkg.__path__ = [tools_dir]
    sys.modules[_PKG] = pkg

    # geometry.py: patched in memory only (tools/geometry.py:53).
    with open(os.path.join(tools_dir, "geometry.py"), "r") as f:
        src = f.read()
    if _GEOMETRY_BAD in src:
        src = src.replace(_GEOMETRY_BAD, _GEOMETRY_FIX)
    geo = types.ModuleType(_PKG + ".geometry")
    geo.__file__ = os.path.join(tools_dir, "geometry.py")
    geo.__package__ = _PKG
    sys.modules[_PKG + ".geometry"] = geo
    exec(compile(src, geo.__file__, "exec"), geo.__dict__)
    pkg.geometry = geo

    spec = importlib.util.spec_from_file_location(
        _PKG + ".urdf_parsing", os.path.join(tools_dir, "urdf_parsing.py")
    )
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_PKG + ".urdf_parsing"] = mod
    spec.loader.exec_module(mod)
    pkg.urdf_parsing = mod
    sys.modules[key] = mod
    return mod
