"""Test infrastructure only (see oracle/Makefile): nothing under miniasm_b200/ imports this package."""
