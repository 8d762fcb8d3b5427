"""Run a tuning script once per library build (OFB_LIB_OVERRIDE): python tools/ab_libs.py script.py [args] -- name=lib.so ..."""
import os, subprocess, sys
i = sys.argv.index('--')
cmd, libs = sys.argv[1:i], [a.split('=', 1) for a in sys.argv[i + 1:]]
for name, lib in [('tree', '')] + libs:
    env = dict(os.environ)
    if lib:
        env['OFB_LIB_OVERRIDE'] = os.path.abspath(lib)
    out = subprocess.run([sys.executable] + cmd, env=env, capture_output=True, text=True)
    print('==', name)
    print(out.stdout.strip() if out.stdout.strip() else out.stderr[-600:], flush=True)
