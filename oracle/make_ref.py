"""Recipe for ``oracle/_ref``: the reference itself, made importable under Python 3.

TEST / BENCHMARK INFRASTRUCTURE, NOT PRODUCT CODE.  ``__graft_entry__.build()`` runs this in the
build container (where ``/root/reference`` is mounted); the output directory ``oracle/_ref/`` is
git-ignored, so no reference source enters this repository's history, but it travels to the GPU box
with the working tree, where ``bench.py --impl reference`` (and the ``cpu_baseline`` leg) time the
reference's OWN numpy scorer on the host cores.

The reference is Python-2 source; the four modules of its numpy path are made runnable by two
purely syntactic rewrites of the source text (the same ones ``tests/golden/refimport.py`` applies in
memory to generate the golden fixtures):

  * ``xrange`` -> ``range``
  * ``print a, b`` statements -> ``print(a, b)``

and a four-line ``__init__`` that leaves out ``fittransform`` (Theano-only).  Nothing else changes:
the arithmetic that runs is the reference's.
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("FITNOISE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref", "fitnoise_ref")
MODULES = ("env", "distributions", "p_adjust", "fitnoise")


def py3(text):
    text = re.sub(r"\bxrange\b", "range", text)
    out = []
    for line in text.split("\n"):
        mo = re.match(r"^(\s*)print (.*?)(,?)\s*$", line)
        if mo:
            indent, body, trailing = mo.groups()
            line = "%sprint(%s%s)" % (indent, body, ", end=' '" if trailing else "")
        out.append(line)
    return "\n".join(out)


def available():
    return os.path.exists(os.path.join(REFERENCE_ROOT, "fitnoise", "fitnoise.py"))


def make():
    """Write oracle/_ref/fitnoise_ref/; returns the directory, or None when the reference is not mounted."""
    if not available():
        return None
    os.makedirs(OUT, exist_ok=True)
    for name in MODULES:
        with open(os.path.join(REFERENCE_ROOT, "fitnoise", name + ".py")) as f:
            source = py3(f.read())
        with open(os.path.join(OUT, name + ".py"), "w") as f:
            f.write(source)
    with open(os.path.join(OUT, "__init__.py"), "w") as f:
        f.write("from .env import as_jsonic\nfrom .p_adjust import fdr\nfrom .fitnoise import *\n")
    return OUT


def load():
    """Import the generated package (``oracle/_ref`` must exist); None when it does not."""
    if not os.path.exists(os.path.join(OUT, "fitnoise.py")):
        return None
    parent = os.path.dirname(OUT)
    if parent not in sys.path:
        sys.path.insert(0, parent)
    import fitnoise_ref
    return fitnoise_ref


if __name__ == "__main__":
    print(make())
