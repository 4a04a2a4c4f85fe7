#!/bin/bash
# Installs the UNMODIFIED reference (roboterax/humanoid-gym) into baseline/_ref for bench.py's reference arm.
#   bash baseline/install_reference.sh          (build container only: needs /root/reference)
# baseline/_ref is git-ignored (never in history) but NOT gpurun-ignored, so it travels to the GPU box.
#
# Deviations from a bare `pip install --target baseline/_ref /root/reference`, and why:
#   * install from a copy under /tmp: /root/reference is read-only and setuptools writes build/ + egg-info there;
#   * --no-deps: install_requires names isaacgym (Preview 4, not installable), mujoco==2.3.6, numpy==1.23.5 ...;
#   * the reference keeps humanoid/envs/base, humanoid/envs/custom and humanoid/scripts as namespace directories
#     (no __init__.py); find_packages() skips those, so the wheel would lack the env classes.  The copy gets EMPTY
#     __init__.py files in those three directories (no reference line is changed) so that they are packaged.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC=/tmp/hg_refcopy
rm -rf "$SRC" "$HERE/_ref"
cp -r /root/reference "$SRC"
for d in humanoid/envs/base humanoid/envs/custom humanoid/scripts; do
  [ -f "$SRC/$d/__init__.py" ] || : > "$SRC/$d/__init__.py"
done
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$SRC" 2>&1 | tail -2
# sanity: every reference .py under humanoid/ is present and byte-identical
( cd /root/reference && find humanoid -name '*.py' | sort ) | while read f; do
  cmp -s "/root/reference/$f" "$HERE/_ref/$f" || { echo "MISMATCH $f"; exit 1; }
done
echo "baseline/_ref ok: $(find "$HERE/_ref/humanoid" -name '*.py' | wc -l) files"
