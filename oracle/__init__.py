"""CPU oracle for the ripped interior-point path -- TEST INFRASTRUCTURE ONLY (see ipm_oracle.py)."""
