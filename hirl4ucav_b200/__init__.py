"""hirl4ucav_b200 -- B200-native (sm_100a) implementation of the HIRL hot path.

Only the path named by BASELINE.json's north_star lives here: the fused agent update
(replay gather, TD target, actor/critic forward+backward, BC blend, Adam, Polyak) and the
batched Harfang observation/reward/lock-flag math.  Everything numerical runs in
hand-written CUDA behind the C ABI declared in ``include/hirl_b200.h``; there is no CPU
fallback -- using an op without the built library raises.
"""
__version__ = "0.1.0"
