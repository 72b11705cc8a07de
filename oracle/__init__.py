"""CPU oracle for the two-tower hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
``two_tower_models_b200`` never imports it and has no CPU fallback.

See ``oracle/two_tower_oracle.py`` for the restatement and its pinning status.
"""

from .two_tower_oracle import *  # noqa: F401,F403
