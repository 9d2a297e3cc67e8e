"""python -m sailfish_b200.run <reference script.py> [reference options] --backends=b200

Runs an unmodified reference script (e.g. sailfish/examples/ldc_3d.py) on the B200 backend.
The reference tree is located through $SAILFISH_ROOT or the script's own location."""
import os
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        sys.stderr.write(__doc__)
        sys.exit(2)
    script = os.path.abspath(sys.argv[1])
    root = os.environ.get('SAILFISH_ROOT')
    if root is None:
        d = os.path.dirname(script)
        while d != '/' and not os.path.isdir(os.path.join(d, 'sailfish')):
            d = os.path.dirname(d)
        root = d
    from sailfish_b200 import plugin
    plugin.install(root)
    sys.argv = [script] + sys.argv[2:]
    if not any(a.startswith('--backends') for a in sys.argv):
        sys.argv.append('--backends=b200')
    runpy.run_path(script, run_name='__main__')


if __name__ == '__main__':
    main()
