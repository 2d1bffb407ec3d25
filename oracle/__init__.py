"""CPU oracle for the geometric-distillation hot path of kaist-cvml/3d-vlm-gd.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import anything from this package, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The product
path (``3d-vlm-gd_b200/gd3``) never imports it and fails loudly when the CUDA
library is missing.

Every function is a plain-PyTorch (CPU, fp32 or fp64) restatement of one
reference function or loss body and cites the reference ``file:line`` it follows
(paths relative to the reference checkout).

Parity pinning: the reference ships no tests, golden vectors or known-answer
fixtures for this path, so the oracle is pinned against
outputs of the *live reference functions* executed in the build container:
``oracle/gen_golden.py`` imports ``utils.losses``, ``utils.functions``,
``utils.model`` and ``mast3r.fast_nn`` from the reference checkout, runs them on
seeded synthetic inputs and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` then checks every oracle function against those
files (and, when the reference checkout is present, against the live functions
again).  The third-party arithmetic under the path is PyTorch's own
(bmm / softmax / grid_sample / cdist / max / LayerNorm / GELU; the reference pins
torch==2.1.2, this image has torch 2.11) -- version drift is limited to fp32
reduction order.
"""

from . import functions, losses, bodies, fast_nn, synth  # noqa: F401
