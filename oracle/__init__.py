"""CPU oracle (test infrastructure only) -- see oracle/grb_oracle.c header."""
from .oracle import *  # noqa: F401,F403
