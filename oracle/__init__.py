"""CPU oracle for the Sound Bubble hot path — test infrastructure only (see tfgridnet_oracle.py)."""
