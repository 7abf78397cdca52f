"""Parity oracle for the COLA particle-mesh force path.  TEST INFRASTRUCTURE ONLY (see pm_oracle.py)."""
