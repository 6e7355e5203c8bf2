"""
Import the *real* reference (AdriaJ/pyxu, mounted read-only at /root/reference) in the build
container so that golden vectors can be produced from its own code.

The image lacks two of the reference's hard dependencies (dask, sparse).  Neither is touched by the
NumPy code path we exercise, so they are replaced by inert stub modules *for this process only*.
Nothing here is used at test time on the GPU box: only the committed .npz fixtures travel.
"""
import importlib.metadata as _ilm
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def load(src=REFERENCE_SRC):
    """src: directory that holds the `pyxu` package (the mounted reference, or its staged copy under oracle/_ref)."""
    _orig = _ilm.version
    _ilm.version = lambda name: "0+reference" if name == "pyxu" else _orig(name)

    def mk(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    if "dask" not in sys.modules:
        dask, da, dac = mk("dask"), mk("dask.array"), mk("dask.array.core")
        mk("dask.distributed"), mk("dask.graph_manipulation")

        class _DaskArray:  # never instantiated
            pass

        dac.Array = da.Array = _DaskArray
        da.core, dask.array = dac, da
        da.linalg = types.SimpleNamespace()
        dask.compute = lambda *a, **k: a
        dask.persist = lambda *a, **k: a
    if "sparse" not in sys.modules:
        sp = mk("sparse")

        class _SparseArray:  # never instantiated
            pass

        sp.SparseArray = _SparseArray

    if src not in sys.path:
        sys.path.insert(0, src)
    import pyxu  # noqa: F401
    import pyxu.abc as pxa
    import pyxu.operator as pxo
    import pyxu.opt.solver as pxs
    import pyxu.opt.stop as pxst

    return types.SimpleNamespace(abc=pxa, operator=pxo, solver=pxs, stop=pxst)
