"""Runs one example script of the reference, UNCHANGED, for a few steps on the b200 backend
(tests/run_examples.sh of the reference does the same on its own backends):
    python tests/reference_example_runner.py <reference root> <examples/....py> <output base> [flags...]
TEST INFRASTRUCTURE: device = tests/fake_device.py (host memory + CPU oracle).  The reference runs the
subdomain in a child process and swallows its errors, so the verdict is what exists afterwards: 'VERDICT ran'
(an output file of the last step was written) or 'VERDICT no output' -- the reason is then on stderr
(UnsupportedConfiguration / B200Error raised while the plan was built, or an unrelated import error)."""
import glob
import os
import runpy
import sys

os.environ.setdefault('OMP_NUM_THREADS', '2')
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ref_root, script, out = sys.argv[1:4]
    flags = sys.argv[4:]
    from sailfish_b200 import plugin
    plugin.install(ref_root, default=True)
    sys.path.insert(0, ref_root)
    sys.path.insert(0, os.path.join(ref_root, os.path.dirname(script)))
    import fake_device
    from oracle import oracle as orc
    orc.build()
    fake_device.install_fake()
    os.chdir(ref_root)
    sys.argv = [script, '--max_iters=4', '--every=4', '--seed=1234', '--quiet', '--output=' + out] + flags
    verdict = 'no output'
    try:
        runpy.run_path(os.path.join(ref_root, script), run_name='__main__')
    except SystemExit:
        pass
    except BaseException as e:      # noqa: B902 -- reported, the caller classifies
        sys.stderr.write('%s: %s\n' % (type(e).__name__, e))
    if glob.glob(out + '.*4.npz'):
        verdict = 'ran'
    print('VERDICT ' + verdict)


if __name__ == '__main__':
    main()
