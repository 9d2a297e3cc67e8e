"""Import shims that let the *unmodified* reference (sailfish, a Python-2-era
code base) import under Python 3.12 / numpy 2 / sympy 1.14.

Only used when the reference tree is available (plugin mode and golden-vector
generation); nothing in the product data path depends on it.

What is shimmed and why:
  * ``imp``                     -- removed in py3.12; used by sailfish/controller.py:21
  * ``np.bool/np.float/np.int`` -- removed in numpy 1.24+; sailfish/subdomain.py:433
  * ``sympy.printing.ccode``    -- renamed ``sympy.printing.c``; sailfish/sym_codegen.py:11
  * ``skimage.morphology``      -- imported at module level in sailfish/util.py:312
  * ``mako.*``                  -- not installed; only ``mako.runtime.Undefined``
                                   is needed at import time (sailfish/sym_force.py:9)
"""
import importlib
import importlib.util
import sys
import types


def _install_imp():
    if 'imp' in sys.modules:
        return
    imp = types.ModuleType('imp')

    def find_module(name, path=None):
        spec = importlib.util.find_spec(name)
        if spec is None:
            raise ImportError(name)
        return None, spec.origin, ('', '', 0)

    def load_source(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    imp.find_module = find_module
    imp.load_source = load_source
    sys.modules['imp'] = imp


def _install_numpy_aliases():
    import numpy as np
    for name, typ in (('bool', bool), ('float', float), ('int', int),
                      ('object', object)):
        if name not in np.__dict__:
            setattr(np, name, typ)
    if not hasattr(np, 'core'):
        np.core = types.SimpleNamespace()
    try:
        np.core.records
    except Exception:  # numpy 2 emits a deprecation shim, keep it quiet
        import numpy.rec
        np.core.records = numpy.rec


def _install_sympy_ccode():
    try:
        importlib.import_module('sympy.printing.ccode')
    except ImportError:
        import sympy.printing.c as c
        sys.modules['sympy.printing.ccode'] = c


def _install_skimage():
    try:
        importlib.import_module('skimage.morphology')
    except ImportError:
        sk = types.ModuleType('skimage')
        sk.__path__ = []
        morph = types.ModuleType('skimage.morphology')
        sk.morphology = morph
        sys.modules['skimage'] = sk
        sys.modules['skimage.morphology'] = morph


def _install_mako():
    try:
        importlib.import_module('mako.runtime')
        return
    except ImportError:
        pass
    mako = types.ModuleType('mako')
    mako.__path__ = []
    runtime = types.ModuleType('mako.runtime')

    class Undefined(object):
        def __bool__(self):
            return False
    runtime.Undefined = Undefined
    runtime.UNDEFINED = Undefined()
    exceptions = types.ModuleType('mako.exceptions')
    lookup = types.ModuleType('mako.lookup')
    template = types.ModuleType('mako.template')

    class TemplateLookup(object):
        def __init__(self, *a, **kw):
            pass

        def get_template(self, name):
            raise RuntimeError('mako is not installed; the b200 backend does '
                               'not render templates (codegen is bypassed)')

    class Template(object):
        def __init__(self, *a, **kw):
            raise RuntimeError('mako is not installed')
    lookup.TemplateLookup = TemplateLookup
    template.Template = Template
    mako.runtime, mako.exceptions = runtime, exceptions
    mako.lookup, mako.template = lookup, template
    for m in (mako, runtime, exceptions, lookup, template):
        sys.modules[m.__name__] = m


def install(reference_root=None):
    """Installs all shims; optionally puts the reference tree on sys.path."""
    _install_imp()
    _install_numpy_aliases()
    _install_sympy_ccode()
    _install_skimage()
    _install_mako()
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
