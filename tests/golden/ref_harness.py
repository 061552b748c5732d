"""Test-side alias of baseline/reference_arm.py (the harness that imports the UNMODIFIED reference).

Kept so that the golden scripts and tests keep their `import ref_harness` spelling; the harness
itself (shims, seeded weights, opt builder, reference SRModel factory) lives in
baseline/reference_arm.py, next to the staged reference tree it drives (baseline/_ref/codes).
The golden fixtures (make_golden.py, make_golden_init.py) were generated with the defaults below:
gpu_ids None (CPU), fp32, batch 1 metadata, torch home /tmp/_golden_torch_home.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from baseline.reference_arm import (build_opt, make_vgg19_checkpoint, ref_root, reference_available,  # noqa: E402,F401
                                    seeded_state)
from baseline.reference_arm import install_shims as _install_shims  # noqa: E402,F401
from baseline import reference_arm as _arm  # noqa: E402

REF_ROOT = ref_root()


def create_reference_model(torch_home="/tmp/_golden_torch_home", **kw):
    return _arm.create_reference_model(torch_home=torch_home, **kw)
