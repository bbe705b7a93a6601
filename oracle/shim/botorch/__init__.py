"""Minimal stand-in for the BoTorch surface the reference's L1 files import (test infrastructure)."""
