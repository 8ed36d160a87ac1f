#!/bin/bash
# Installs the UNMODIFIED reference (themattinthehatt/behavenet) into baseline/_ref (git-ignored, travels
# with gpurun) for bench.py --impl reference and cpu_baseline.kind == "reference".
#
# The reference's setup.py lists packages=['behavenet', 'tests'] only -- it is meant for `pip install -e`
# (docs/source/installation.rst) -- so a plain install drops every subpackage.  The install therefore runs
# from a copy under /tmp whose setup.py enumerates the subpackages with find_packages(); no module of the
# package is touched.  Dependencies are not resolved (--no-deps): torch / numpy / scipy / sklearn of this
# image are used; commentjson, h5py and test_tube are absent and are not imported on the timed path
# (bench.py stubs commentjson, which ae_model_architecture_generator imports at module top).
set -e
cd "$(dirname "$0")/.."
SRC=${1:-/root/reference}
rm -rf /tmp/behavenet_ref_src baseline/_ref
cp -r "$SRC" /tmp/behavenet_ref_src
python - <<'PY'
import re
p = '/tmp/behavenet_ref_src/setup.py'
s = open(p).read()
s = s.replace("from distutils.core import setup", "from setuptools import setup, find_packages")
s = s.replace("packages=['behavenet', 'tests']", "packages=find_packages(include=['behavenet', 'behavenet.*'])")
open(p, 'w').write(s)
PY
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target baseline/_ref /tmp/behavenet_ref_src
python - <<'PY'
import filecmp, os, sys
bad = []
for root, _, files in os.walk('baseline/_ref/behavenet'):
    for f in files:
        if f.endswith('.py'):
            a = os.path.join(root, f)
            b = os.path.join('/tmp/behavenet_ref_src', os.path.relpath(a, 'baseline/_ref'))
            if not filecmp.cmp(a, b, shallow=False):
                bad.append(a)
assert not bad, bad
print('baseline/_ref: %d python files, identical to the source tree' %
      sum(f.endswith('.py') for _, _, fs in os.walk('baseline/_ref/behavenet') for f in fs))
PY
