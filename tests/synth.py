"""Test-side alias of oracle/synth.py (numpy mirror of the read generator and
ragged batch builders)."""
from oracle.synth import *  # noqa: F401,F403
from oracle.synth import ragged_batch, synth_reads, uniform_offsets  # noqa: F401
