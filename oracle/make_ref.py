#!/usr/bin/env python
"""TEST INFRASTRUCTURE - builds ``oracle/_ref/spfem``: a mechanically Python-3
patched working copy of the (Python 2.7) reference package.

The reference (``/root/reference/spfem``) cannot be imported by the only
interpreter in this image (CPython 3.12), so -- exactly like compiling a C
reference from its own sources -- this recipe *transpiles* it: the sources are
read where they lie, a fixed list of purely syntactic rewrites is applied and
the result is written ONLY into ``oracle/_ref/`` (git-ignored, never committed;
it still travels to the GPU box with the snapshot like a built ``.so``).

No semantic edits are made.  The rewrite list:
  1. ``from mesh import *``            -> ``from .mesh import *``  (__init__.py)
  2. ``print x``                       -> ``print(x)``
  3. ``inspect.getargspec``            -> ``inspect.getfullargspec``
  4. ``np.float_``                     -> ``np.float64``
  5. ``leggauss(np.ceil(..))``         -> ``leggauss(int(np.ceil(..)))``
  6. boolean ``(-I1)``                 -> ``(~I1)``                 (mesh.py)
  7. ``.iteritems()`` -> ``.items()``; ``map(sp.hstack, ..)`` -> ``list(map(..))``
  8. matplotlib / mpl_toolkits stubs (not installed here; plotting only)
  9. ``range(n) + k`` -> ``np.arange(n) + k``       (MeshQuad._single_refine)

Usage:  python oracle/make_ref.py [--src /root/reference] [--dst oracle/_ref]
"""
import argparse
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

MPL_STUB = '''"""Stub: matplotlib is not installed in this image (plotting is out of scope)."""
class _Anything(object):
    def __getattr__(self, name):
        return _Anything()
    def __call__(self, *a, **k):
        return _Anything()
import sys as _sys, types as _types
def _install():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.tri",
                 "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if name not in _sys.modules:
            m = _types.ModuleType(name)
            m.__getattr__ = lambda attr, _n=name: _Anything()
            m.__path__ = []
            _sys.modules[name] = m
    _sys.modules["mpl_toolkits.mplot3d"].Axes3D = _Anything()
try:
    import matplotlib.pyplot  # noqa: F401
    import matplotlib.tri  # noqa: F401
    from mpl_toolkits.mplot3d import Axes3D  # noqa: F401
except Exception:
    _install()
'''

_print_re = re.compile(r'^(\s*)print\s+(?!\()(.*?)\s*$')


def _fix_prints(lines):
    out = []
    i = 0
    while i < len(lines):
        ln = lines[i]
        m = _print_re.match(ln.rstrip("\n"))
        if m and not ln.lstrip().startswith("#"):
            body = m.group(2)
            # explicit line continuation: swallow the following lines
            while body.endswith("\\"):
                i += 1
                body = body[:-1].rstrip() + " " + lines[i].strip()
            out.append("%sprint(%s)\n" % (m.group(1), body))
        else:
            out.append(ln)
        i += 1
    return out


def patch_source(name, text):
    lines = _fix_prints(text.splitlines(True))
    text = "".join(lines)
    text = text.replace("inspect.getargspec", "inspect.getfullargspec")
    text = text.replace("np.float_)", "np.float64)")
    text = text.replace(".iteritems()", ".items()")
    text = re.sub(r"leggauss\(np\.ceil\((.*?)\)\)\n",
                  r"leggauss(int(np.ceil(\1)))\n", text)
    if name == "__init__.py":
        text = re.sub(r"^from (\w+) import \*", r"from .\1 import *", text,
                      flags=re.M)
        text = "from . import _mplstub  # noqa: F401 (py3 patch 8)\n" + text
    if name == "mesh.py":
        text = text.replace("(-I1)", "(~I1)").replace("(-I2)", "(~I2)")
        text = text.replace("(-I3)", "(~I3)")
        text = text.replace("mid = range(self.t.shape[1]) + np.max(t2f) + 1",
                            "mid = np.arange(self.t.shape[1]) + np.max(t2f) + 1")
        text = text.replace("np.issubdtype(self.t[0, 0], int)",
                            "np.issubdtype(self.t.dtype, np.integer)")
    if name == "element.py":
        # Python 2 integer division (SURVEY.md 8(c) patch 7): ElementTriPp's
        # interior DOF count and AbstractElementTriPp's
        text = text.replace("(p-1)*(p-2)/2", "(p-1)*(p-2)//2")
    if name == "utils.py":
        text = text.replace("sp.vstack(map(sp.hstack, block))",
                            "sp.vstack(list(map(sp.hstack, block)))")
    if name in ("mesh.py", "element.py", "utils.py", "test_module.py",
                "test_asm.py", "test_mesh.py", "test_mapping.py",
                "test_perf.py", "weakform.py", "geometry.py"):
        text = "import spfem._mplstub  # noqa: F401 (py3 patch 8)\n" + text \
            if not text.startswith("# -*-") else \
            text.split("\n", 1)[0] + "\nimport spfem._mplstub  # noqa: F401 (py3 patch 8)\n" + text.split("\n", 1)[1]
    return text


def build(src="/root/reference", dst=None, quiet=False):
    dst = dst or os.path.join(HERE, "_ref")
    srcpkg = os.path.join(src, "spfem")
    if not os.path.isdir(srcpkg):
        return None
    outpkg = os.path.join(dst, "spfem")
    os.makedirs(outpkg, exist_ok=True)
    for fn in sorted(os.listdir(srcpkg)):
        if not fn.endswith(".py"):
            continue
        with open(os.path.join(srcpkg, fn), "r", encoding="utf-8") as fh:
            text = fh.read()
        with open(os.path.join(outpkg, fn), "w", encoding="utf-8") as fh:
            fh.write(patch_source(fn, text))
    with open(os.path.join(outpkg, "_mplstub.py"), "w") as fh:
        fh.write(MPL_STUB)
    with open(os.path.join(dst, "README"), "w") as fh:
        fh.write("Generated by oracle/make_ref.py from %s - do not commit.\n" % src)
    if not quiet:
        print("oracle/_ref built at", outpkg)
    return outpkg


def import_ref(dst=None):
    """Import the patched reference as top-level package ``spfem`` (or None)."""
    dst = dst or os.path.join(HERE, "_ref")
    if not os.path.isdir(os.path.join(dst, "spfem")):
        return None
    if dst not in sys.path:
        sys.path.insert(0, dst)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import spfem  # noqa
        import spfem.mesh, spfem.assembly, spfem.element, spfem.mapping  # noqa
        import spfem.quadrature  # noqa
    return spfem


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dst", default=None)
    a = ap.parse_args()
    if build(a.src, a.dst) is None:
        print("reference sources not found at", a.src)
        sys.exit(1)
