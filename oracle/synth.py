"""Re-export of the synthetic weight/input generator (moved to audiolcm_b200/synth.py so that the
product-side bench never imports from oracle/)."""
from audiolcm_b200.synth import *  # noqa: F401,F403
from audiolcm_b200.synth import AttrDict  # noqa: F401
