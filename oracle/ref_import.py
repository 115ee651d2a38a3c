"""Import the UNMODIFIED reference package from /root/reference under a stub importer.
TEST INFRASTRUCTURE ONLY; usable only in the build container (``/root/reference`` does not exist on
the GPU box), i.e. by ``oracle/make_golden.py`` and by CPU tests that skip when it is absent.

The reference imports dolfin, hippylib, mpi4py, matplotlib, pylab and ufl at module import time
(hippyflow/collectives/collective.py:15-17, hippyflow/modeling/PODProjector.py:14-30); none is
installed here.  A ``sys.meta_path`` finder fabricates permissive empty modules for those roots so
the pure NumPy/SciPy code of the reference runs verbatim.  ``hippylib`` is
stubbed with the NumPy restatement in ``oracle/hippylib_np.py`` for the names the path uses.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
_STUB_ROOTS = ("dolfin", "hippylib", "mpi4py", "matplotlib", "pylab", "ufl", "petsc4py", "slepc4py", "pympler")


class _StubModule(types.ModuleType):
    __all__ = []
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name[0].isupper():
            obj = type(name, (), {})          # usable as a base class at import time
        else:
            obj = _StubModule(self.__name__ + "." + name)
        setattr(self, name, obj)
        return obj

    def __call__(self, *a, **k):
        return None


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "hippyflow"))


def import_reference():
    """Return the reference's ``hippyflow`` module (v0.2.0), imported verbatim."""
    if not available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    if "hippyflow" in sys.modules and getattr(sys.modules["hippyflow"], "__file__", "").startswith(REFERENCE_ROOT):
        return sys.modules["hippyflow"]
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _StubFinder())
    import hippylib as hp  # the stub
    from . import hippylib_np as hnp
    hp.ParameterList = dict
    hp.STATE, hp.PARAMETER, hp.ADJOINT = 0, 1, 2
    for name in ("MultiVector", "LowRankOperator", "Solver2Operator", "MatMvMult", "MatMvTranspmult",
                 "MvDSmatMult", "doublePass", "doublePassG"):
        setattr(hp, name, getattr(hnp, name))
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import hippyflow
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return hippyflow
