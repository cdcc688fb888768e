"""Memory-safety evidence without compute-sanitizer (closed on this GPU pool): liborbx_boundscheck.so is the product
source compiled with -DORBX_BOUNDS_CHECK, which puts a device-side trap on every shared-memory tile / queue / score-map /
bitmap index of the FAST kernels (strip path, per-cell path, minThFAST retries) and on the descriptor's pattern samples.
A violated bound fails the launch (cudaErrorLaunchFailure).  The five BASELINE configs and the adversarial frames run under
it in a child process (the library is chosen at import time by ORBX_LIB) and must give the product library's bytes."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from orbslam2_with_quadrics_b200 import frames as fr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import hashlib, json, sys
import numpy as np
sys.path.insert(0, %(root)r)
from orbslam2_with_quadrics_b200 import ORBextractor, _capi
from orbslam2_with_quadrics_b200 import frames as fr
out = {"lib": _capi.LIB_PATH}
def digest(res):
    h = hashlib.sha256()
    for kp, desc in res:
        h.update(kp.tobytes()); h.update(desc.tobytes())
    return h.hexdigest()
for name, (w, h, nf, sf, nl, it, mt, nimg) in fr.CONFIGS.items():
    ex = ORBextractor(nf, sf, nl, it, mt, max_batch=3)
    imgs = [fr.cluttered_scene(w, h, 4000 + i) for i in range(2)] + [fr.checker_frame(w, h, 5)]
    out[name] = digest(ex.extract_batch(imgs))
    if name in ("mono_tum", "stereo_kitti"):
        ex2 = ORBextractor(nf, sf, nl, it, mt, max_batch=2, candidate_divisor=1)     # noise: dense strips -> per-cell path
        out[name + "/noise+flat"] = digest(ex2.extract_batch([fr.noise_frame(w, h, 9), fr.flat_frame(w, h)]))
        ex2.close()
    ex.close()
print(json.dumps(out))
"""


def _run(lib):
    env = dict(os.environ)
    if lib:
        env["ORBX_LIB"] = lib
    else:
        env.pop("ORBX_LIB", None)
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_all_configs_run_clean_under_device_side_bounds_checks():
    from orbslam2_with_quadrics_b200 import build
    lib = build.build_library(bounds_check=True)
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    assert sass.count("BPT.TRAP") > 20, "the bounds-check build carries no traps"
    checked = _run(lib)
    plain = _run(None)
    assert checked.pop("lib").endswith("liborbx_boundscheck.so") and plain.pop("lib").endswith("liborbx.so")
    assert checked == plain and len(checked) == 7
