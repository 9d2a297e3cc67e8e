"""Runs one of the reference's own regression tests (regtest/subdomains/*.py), UNCHANGED, on the b200
backend:  python tests/reference_regtest_runner.py <reference root> <module | path/in/reference.py> <AB|AA> [max_iters]

The reference's multi-process machinery (LBMachineMaster, one process per subdomain, ZeroMQ connectors,
collect/distribute kernels, sparse index exchange for x faces, edges and corners) runs as is;
sailfish_b200.plugin selects the backend.  TEST INFRASTRUCTURE: the build container has no GPU, so the
device behind the C ABI is tests/fake_device.py (host memory + the CPU oracle), inherited by the forked
subdomain processes.  Prints one JSON line: {"<test id>": "ok" | "fail" | "error", ...}."""
import importlib.util
import json
import os
import sys
import unittest

os.environ.setdefault('OMP_NUM_THREADS', '1')        # one process per subdomain already
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ref_root, name, access_pattern = sys.argv[1:4]
    max_iters = int(sys.argv[4]) if len(sys.argv) > 4 else None
    from sailfish_b200 import plugin
    plugin.install(ref_root, default=True)
    sys.path.insert(0, ref_root)
    import fake_device
    from oracle import oracle as orc
    orc.build()
    fake_device.install_fake()

    path = os.path.join(ref_root, name) if name.endswith('.py') else os.path.join(ref_root, 'regtest', 'subdomains', name + '.py')
    spec = importlib.util.spec_from_file_location('regtest_mod', path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules['regtest_mod'] = mod
    spec.loader.exec_module(mod)
    # what the module's __main__ block takes from the command line (regtest/subdomains/util.py:8-24)
    mod.block_size = 64
    mod.access_pattern = access_pattern
    mod.node_addressing = 'direct'
    for attr in ('MAX_ITERS', 'MAX_ITERS_2D', 'MAX_ITERS_3D'):
        if max_iters is not None and hasattr(mod, attr):
            setattr(mod, attr, max_iters)

    class Collect(unittest.TextTestResult):
        outcomes = {}

        def addSuccess(self, test):
            super().addSuccess(test)
            self.outcomes[test.id()] = 'ok'

        def addFailure(self, test, err):
            super().addFailure(test, err)
            self.outcomes[str(test.id() if hasattr(test, 'id') else test)] = 'fail'

        def addError(self, test, err):
            super().addError(test, err)
            self.outcomes[str(test.id() if hasattr(test, 'id') else test)] = 'error'

    suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
    res = unittest.TextTestRunner(stream=sys.stderr, verbosity=1, resultclass=Collect).run(suite)
    print(json.dumps(res.outcomes))


if __name__ == '__main__':
    main()
