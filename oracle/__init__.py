"""ORACLE package: test infrastructure only (see torch_oracle.py / warp_oracle.c headers)."""
