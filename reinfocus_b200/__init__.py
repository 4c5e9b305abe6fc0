"""reinfocus_b200: B200-native (sm_100a) implementation of the reinfocus per-step hot path.

Drop-in module layout (reference module -> this package):
    reinfocus.graphics.render  -> reinfocus_b200.graphics.render  (FastRenderer, render)
    reinfocus.graphics.random  -> reinfocus_b200.graphics.random  (make_random_states)
    reinfocus.vision           -> reinfocus_b200.vision           (focus_value, focus_values)
    reinfocus.environments.*   -> reinfocus_b200.environments.*   (Environment, VectorEnvironment
                                                                   and the strategy objects)
The compute lives in libreinfocus_b200.so (hand-written CUDA behind a C-ABI, see
include/reinfocus_b200.h); there is no CPU fallback.
"""

__version__ = "0.1.0"
