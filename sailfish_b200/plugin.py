"""Plugs ``backend_b200`` into the UNMODIFIED reference (sailfish-team/sailfish).

    from sailfish_b200 import plugin
    plugin.install('/path/to/sailfish')          # before the first `import sailfish`
    # ... then run any reference script with  --backends=b200

or from the shell:  python -m sailfish_b200.run /path/to/sailfish/examples/ldc_3d.py --backends=b200 ...

What install() does:
  1. installs the import shims that let the Python-2-era reference import on py3.12
     (sailfish_b200/compat.py);
  2. registers this package's backend as the module ``sailfish.backend_b200`` -- all
     ``sailfish.util.get_backends`` needs (util.py:52-59: ``__import__('sailfish',
     fromlist=['backend_b200'])`` then ``sys.modules['sailfish.backend_b200'].backend``);
  3. replaces ``BlockCodeGenerator.get_code`` (codegen.py:104-180), which would render the Mako
     templates into CUDA source, by a function that returns the *context* dictionary the templates
     would have received (codegen.py:195-204); ``B200Backend.build`` turns it into a plan.
With ``default=True`` the backend is also tried first whatever ``--backends`` says (the reference's
default is 'cuda,opencl', and its option parser only asks *those* backends for their flags,
controller.py:452-455): scripts and regtests that never mention a backend then run on it unchanged.
Nothing in the reference tree is modified.
"""
import sys
import types


def install(reference_root=None, default=False):
    from sailfish_b200 import compat
    compat.install(reference_root)
    import sailfish  # noqa: F401  (the reference)
    from sailfish import codegen
    from sailfish_b200 import backend_b200

    mod = types.ModuleType('sailfish.backend_b200')
    mod.backend = backend_b200.B200Backend
    mod.__doc__ = backend_b200.__doc__
    sys.modules['sailfish.backend_b200'] = mod
    setattr(sys.modules['sailfish'], 'backend_b200', mod)

    def get_code(self, subdomain_runner, target_type):
        ctx = self._build_context(subdomain_runner)
        return ctx

    codegen.BlockCodeGenerator.get_code = get_code

    if default:
        from sailfish import util
        if not getattr(util.get_backends, '_b200_first', False):
            stock = util.get_backends

            def get_backends(backends=('cuda', 'opencl')):
                names = ['b200'] + [b for b in backends if b != 'b200']
                return stock(names)

            get_backends._b200_first = True
            util.get_backends = get_backends
    return mod
