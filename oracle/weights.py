"""Synthetic weights / inputs used by the oracle-side tests: re-export of aptai_b200.synth (data generation only,
no reference arithmetic)."""
from aptai_b200.synth import *  # noqa: F401,F403
from aptai_b200.synth import backbone_state_dict, linear_params, waveforms, phoneme_sequences  # noqa: F401
