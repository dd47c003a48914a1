"""The BASELINE.json workloads live in the package (cylindertag_b200/workloads.py); re-exported for the tests."""
from cylindertag_b200.workloads import *  # noqa: F401,F403
from cylindertag_b200.workloads import CODEBOOKS, CONFIG3_FRAMES, CONFIG4_FRAMES, CONFIG5_DISTINCT, codebook, config3_frame, config4_frame, config4_markers, config5_frame, render_many  # noqa: F401
