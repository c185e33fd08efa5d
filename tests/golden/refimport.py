"""Import the UNMODIFIED reference (pfh/fitnoise, Python 2) under Python 3, in memory.

Test infrastructure only.  Used by ``tests/golden/make_golden.py`` (run in the build
container, where ``/root/reference`` is mounted) to generate the golden fixtures that
pin the oracle.  Nothing here is imported by the product (``fitnoise_b200``), by
``-m gpu`` tests, by ``smoke()`` or by ``bench.py``: ``/root/reference`` does not exist
on the GPU box.

The reference is Python-2 source.  Its numpy path (``use_theano=False``) runs under
Python 3 after two purely syntactic rewrites applied to the source TEXT at import time:

  * ``xrange`` -> ``range``
  * ``print a, b`` statements -> ``print(a, b)``  (6 sites in fitnoise/fitnoise.py)

No reference source is written to disk or copied into this repository; the four modules
are read from ``/root/reference/fitnoise`` and exec'd into synthetic module objects.
``fittransform`` builds its objective from Theano symbols and cannot be RUN without Theano, but the
module imports, and its numpy-compatible methods (``_unpack_transform``, ``_apply_transform``,
``_configured``) work on plain arrays: ``load_transform()`` gives access to those.
"""
import os
import re
import sys
import types

REFERENCE_ROOT = os.environ.get("FITNOISE_REFERENCE", "/root/reference")
_PKG = "_fitnoise_reference"


def _py3(text):
    text = re.sub(r"\bxrange\b", "range", text)
    out = []
    for line in text.split("\n"):
        mo = re.match(r"^(\s*)print (.*?)(,?)\s*$", line)
        if mo:
            indent, body, trailing = mo.groups()
            end = ", end=' '" if trailing else ""
            line = "%sprint(%s%s)" % (indent, body, end)
        out.append(line)
    return "\n".join(out)


def available():
    return os.path.exists(os.path.join(REFERENCE_ROOT, "fitnoise", "fitnoise.py"))


def load():
    """Return the reference package as a module object with the public names of
    ``fitnoise/__init__.py`` (minus ``transform``)."""
    if _PKG in sys.modules:
        return sys.modules[_PKG]
    if not available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    pkg = types.ModuleType(_PKG)
    pkg.__path__ = []
    sys.modules[_PKG] = pkg
    for name in ("env", "distributions", "p_adjust", "fitnoise"):
        path = os.path.join(REFERENCE_ROOT, "fitnoise", name + ".py")
        with open(path) as f:
            source = _py3(f.read())
        mod = types.ModuleType(_PKG + "." + name)
        mod.__package__ = _PKG
        mod.__file__ = path
        sys.modules[mod.__name__] = mod
        setattr(pkg, name, mod)
        exec(compile(source, path, "exec"), mod.__dict__)
    for key, value in pkg.fitnoise.__dict__.items():
        if not key.startswith("_"):
            setattr(pkg, key, value)
    pkg.fdr = pkg.p_adjust.fdr
    pkg.as_jsonic = pkg.env.as_jsonic
    return pkg


def load_transform():
    """The reference's ``fittransform`` module (its ``fit`` needs Theano; the polynomial transform
    itself, its parameter layout, initial values and bounds are plain numpy)."""
    pkg = load()
    name = _PKG + ".fittransform"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(REFERENCE_ROOT, "fitnoise", "fittransform.py")
    with open(path) as f:
        source = _py3(f.read())
    mod = types.ModuleType(name)
    mod.__package__ = _PKG
    mod.__file__ = path
    sys.modules[name] = mod
    setattr(pkg, "fittransform", mod)
    exec(compile(source, path, "exec"), mod.__dict__)
    return mod
