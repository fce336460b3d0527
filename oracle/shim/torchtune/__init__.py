"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

A plain-PyTorch restatement of the slice of ``torchtune==0.4.0`` that the
reference's ``sesameai/models.py`` calls (reference ``requirements.txt:7``;
call sites ``sesameai/models.py:5,7,10-39,126-127,153,158,170,173,187-188``).
torchtune itself is not vendored in the reference and is not installed in this
image, so its published semantics are restated here.

Putting ``oracle/shim`` on ``sys.path`` makes ``import torchtune`` resolve to
this package, which lets the *unmodified* reference ``sesameai/models.py`` be
imported in the build container to validate the oracle and to generate the
golden vectors under ``tests/golden/``.

Parity status: the bf16 rounding points follow torchtune 0.4.0 as restated from
its published source (not re-verifiable offline); the fp32 mathematics is
pinned against the independent ``transformers`` CSM port (tests/test_oracle_pin.py).
"""
from . import modules  # noqa: F401
from . import models  # noqa: F401

__version__ = "0.4.0+oracle-shim"
